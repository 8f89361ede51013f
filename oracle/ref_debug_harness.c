/*
 * oracle/ref_debug_harness.c -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * Recording stand-in for OpenGL and the shader module around the UNMODIFIED reference debug renderer
 * (/root/reference/src/physics/debug_renderer.c, compiled where it lies by oracle/Makefile into
 * oracle/_ref/libcrux_ref_debug.so together with the physics sources and ref_harness.c).  Nothing here computes
 * geometry: the GL entry points (glad's function pointers, third_party/glad/glad.h) and shader_create / shader_use /
 * shader_set_mat4 (include/shader.h) are replaced by functions that write down what the reference hands them --
 * the vertex and index arrays of glBufferData (init: debug_renderer.c:339-356, :434-455, :618-639), the 8 corners of
 * glBufferSubData (:689), the model matrices of shader_set_mat4 (:660, :700, :709) and the element counts of
 * glDrawElements -- in call order.  tests/make_golden.py turns the log into tests/golden/debug_geometry.npz, the
 * fixture the CUDA debug-geometry kernel and the host mesh helpers (crux_b200/csrc/pg_debug.cu) are compared with.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "glad.h"
#include "scene.h"
#include "shader.h"
#include "physics/world.h"
#include "physics/debug_renderer.h"

void *ref_world_pw(void *h);

enum { CAP_VERTEX_DATA = 1, CAP_INDEX_DATA = 2, CAP_VERTEX_SUBDATA = 3, CAP_MODEL_MAT4 = 4, CAP_DRAW = 5 };

typedef struct {
  int what;          /* CAP_* */
  int count;         /* floats / indices / element count of the draw call */
  int tag;           /* CAP_MODEL_MAT4: shader id (1 wireframe, 2 translucent); CAP_DRAW: GL mode */
  size_t offset;     /* into g_pool, in 4-byte words */
} Capture;

static Capture *g_cap = NULL;
static int g_ncap = 0, g_capcap = 0;
static unsigned int *g_pool = NULL;
static size_t g_npool = 0, g_poolcap = 0;
static GLuint g_next_name = 1;

static void cap_reset(void) { g_ncap = 0; g_npool = 0; }
static void cap_push(int what, int count, int tag, const void *words, int n_words) {
  if (g_ncap == g_capcap) { g_capcap = g_capcap ? 2 * g_capcap : 256; g_cap = (Capture *)realloc(g_cap, sizeof(Capture) * (size_t)g_capcap); }
  if (g_npool + (size_t)n_words > g_poolcap) {
    while (g_npool + (size_t)n_words > g_poolcap) g_poolcap = g_poolcap ? 2 * g_poolcap : 1u << 16;
    g_pool = (unsigned int *)realloc(g_pool, 4 * g_poolcap);
  }
  Capture c = {what, count, tag, g_npool};
  if (n_words) memcpy(g_pool + g_npool, words, 4 * (size_t)n_words);
  g_npool += (size_t)n_words;
  g_cap[g_ncap++] = c;
}

/* ---- OpenGL, as far as debug_renderer.c calls it ---- */
static void APIENTRY s_BindVertexArray(GLuint a) { (void)a; }
static void APIENTRY s_BindBuffer(GLenum t, GLuint b) { (void)t; (void)b; }
static void APIENTRY s_GenVertexArrays(GLsizei n, GLuint *a) { for (GLsizei k = 0; k < n; k++) a[k] = g_next_name++; }
static void APIENTRY s_GenBuffers(GLsizei n, GLuint *a) { for (GLsizei k = 0; k < n; k++) a[k] = g_next_name++; }
static void APIENTRY s_BufferData(GLenum target, GLsizeiptr size, const void *data, GLenum usage) {
  (void)usage;
  cap_push(target == GL_ELEMENT_ARRAY_BUFFER ? CAP_INDEX_DATA : CAP_VERTEX_DATA, (int)(size / 4), 0, data, (int)(size / 4));
}
static void APIENTRY s_BufferSubData(GLenum target, GLintptr offset, GLsizeiptr size, const void *data) {
  (void)target; (void)offset;
  cap_push(CAP_VERTEX_SUBDATA, (int)(size / 4), 0, data, (int)(size / 4));
}
static void APIENTRY s_EnableVertexAttribArray(GLuint i) { (void)i; }
static void APIENTRY s_VertexAttribPointer(GLuint i, GLint s, GLenum t, GLboolean nrm, GLsizei st, const void *p) {
  (void)i; (void)s; (void)t; (void)nrm; (void)st; (void)p;
}
static void APIENTRY s_LineWidth(GLfloat w) { (void)w; }
static void APIENTRY s_DrawElements(GLenum mode, GLsizei count, GLenum type, const void *indices) {
  (void)type; (void)indices;
  cap_push(CAP_DRAW, (int)count, (int)mode, NULL, 0);
}
static GLuint APIENTRY s_GetUniformBlockIndex(GLuint p, const GLchar *n) { (void)p; (void)n; return 0; }
static void APIENTRY s_UniformBlockBinding(GLuint p, GLuint i, GLuint b) { (void)p; (void)i; (void)b; }

PFNGLBINDVERTEXARRAYPROC glad_glBindVertexArray = s_BindVertexArray;
PFNGLBINDBUFFERPROC glad_glBindBuffer = s_BindBuffer;
PFNGLGENVERTEXARRAYSPROC glad_glGenVertexArrays = s_GenVertexArrays;
PFNGLGENBUFFERSPROC glad_glGenBuffers = s_GenBuffers;
PFNGLBUFFERDATAPROC glad_glBufferData = s_BufferData;
PFNGLBUFFERSUBDATAPROC glad_glBufferSubData = s_BufferSubData;
PFNGLENABLEVERTEXATTRIBARRAYPROC glad_glEnableVertexAttribArray = s_EnableVertexAttribArray;
PFNGLVERTEXATTRIBPOINTERPROC glad_glVertexAttribPointer = s_VertexAttribPointer;
PFNGLLINEWIDTHPROC glad_glLineWidth = s_LineWidth;
PFNGLDRAWELEMENTSPROC glad_glDrawElements = s_DrawElements;
PFNGLGETUNIFORMBLOCKINDEXPROC glad_glGetUniformBlockIndex = s_GetUniformBlockIndex;
PFNGLUNIFORMBLOCKBINDINGPROC glad_glUniformBlockBinding = s_UniformBlockBinding;

/* ---- shader module (include/shader.h) ---- */
static Shader g_shaders[2] = {{1}, {2}};
static int g_shader_next = 0;
Shader *shader_create(const char *vs, const char *fs) { (void)vs; (void)fs; return &g_shaders[g_shader_next++ & 1]; }
void shader_use(const Shader *s) { (void)s; }
void shader_set_mat4(const Shader *s, const char *name, mat4 value) {
  (void)name;
  cap_push(CAP_MODEL_MAT4, 16, (int)s->ID, value, 16);
}

/* ---- exported ---- */
int ref_debug_init(void *world) {            /* physics_debug_renderer_init: one mesh upload per drawable body */
  cap_reset();
  g_shader_next = 0;
  physics_debug_renderer_init((struct PhysicsWorld *)ref_world_pw(world));
  return g_ncap;
}
int ref_debug_render(void *world) {          /* physics_debug_render: one frame */
  struct RenderContext ctx;
  memset(&ctx, 0, sizeof(ctx));
  cap_reset();
  physics_debug_render((struct PhysicsWorld *)ref_world_pw(world), &ctx);
  return g_ncap;
}
int ref_debug_capture(int k, int *what, int *count, int *tag, void *out_words, int cap_words) {
  if (k < 0 || k >= g_ncap) return -1;
  const Capture *c = &g_cap[k];
  *what = c->what; *count = c->count; *tag = c->tag;
  const int n = (c->what == CAP_DRAW) ? 0 : c->count;
  if (n > cap_words) return -2;
  if (n) memcpy(out_words, g_pool + c->offset, 4 * (size_t)n);
  return n;
}
