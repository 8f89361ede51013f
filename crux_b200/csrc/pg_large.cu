// pg_large.cu -- one large world (uniform-grid broadphase).  PLACEHOLDER entry points: the
// pipeline lands in the next commit; until then every call reports PG_ERR_UNSUPPORTED loudly.
#include "pg_common.cuh"

extern "C" {
PgLarge *pg_large_create(int, int, int) { pg_set_error("pg_large_*: not built yet"); return nullptr; }
void pg_large_destroy(PgLarge *) {}
int pg_large_set(PgLarge *, const PgBody *, int, int, const float *, const float *, const float *, float, float) { return PG_ERR_UNSUPPORTED; }
int pg_large_upload_state(PgLarge *, const float *, const float *, const uint8_t *) { return PG_ERR_UNSUPPORTED; }
int pg_large_download_state(PgLarge *, float *, float *, uint8_t *) { return PG_ERR_UNSUPPORTED; }
int pg_large_step(PgLarge *, float, int) { return PG_ERR_UNSUPPORTED; }
int pg_large_sync(PgLarge *) { return PG_ERR_UNSUPPORTED; }
long long pg_large_pairs(PgLarge *, float, int32_t *, long long) { return PG_ERR_UNSUPPORTED; }
long long pg_large_events(PgLarge *, PgEvent *, long long, int *) { return PG_ERR_UNSUPPORTED; }
int pg_large_stats(PgLarge *, PgStats *, void *) { return PG_ERR_UNSUPPORTED; }
long long pg_large_launches(const PgLarge *) { return 0; }
float pg_large_last_step_ms(PgLarge *) { return -1.0f; }
int pg_large_kernel_count(void) { return 0; }
const char *pg_large_kernel_name(int) { return ""; }
float pg_large_kernel_ms(PgLarge *, int) { return -1.0f; }
int pg_large_exactness(PgLarge *, int *, int *, long long *, long long *) { return PG_ERR_UNSUPPORTED; }
}
