/* oracle/dropin_stubs.c -- TEST INFRASTRUCTURE ONLY.  The two engine symbols libcrux_physics.so expects from
 * the game (src/event.c) so that the reference's Unity test binary links against the drop-in library. */
struct GameEventOpaque { char bytes[64]; };
void game_event_queue_enqueue(struct GameEventOpaque ev) { (void)ev; }
int get_event_type(int a, int b) { (void)a; (void)b; return 0; }
