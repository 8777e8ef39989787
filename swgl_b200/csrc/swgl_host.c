/*
 * swgl_host.c -- the GL-style state layer of libswgl_b200.so (plain C).
 *
 * Mirrors the reference's cold API (swgl.c:2870-3147, 3151-3222, 3713-3928): one implicit
 * global context, small-integer object ids (shaders from 0; programs, vertex arrays, buffers
 * and textures from 1), silent returns on invalid state, caller data copied at upload.  The
 * hot path -- glClear and glDrawArrays / glDrawElements -- is handed to the CUDA device layer
 * through the C ABI of swgl_dev.h; nothing in this file computes a vertex or a pixel.
 *
 * Not thread-safe and not re-entrant, exactly like the reference.
 */
#include "swgl_b200.h"
#include "swgl_dev.h"
#include "swgl_glsl.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* swgl.c:15-16 */
#define RMIN(x, y) (((x) < (y)) ? (x) : (y))
#define RMAX(x, y) (((x) > (y)) ? (x) : (y))

/* ---------------------------------------------------------------------------------------- */
typedef struct
{
	swgldev_ptr data;
	uint32_t size;          /* bytes; 0 = never specified (swgl.c:3140) */
	uint32_t capacity;
	uint32_t max_index;     /* element data only: largest u32 in the buffer */
	void* origin;           /* a VAO's own struct: the named buffer it is a snapshot of */
} gl_buffer;

typedef struct
{
	uint32_t stride;
	uint8_t normalized;
	uint32_t type;
	int32_t size;
	uint32_t index;
	int32_t offset;
} gl_attrib;

typedef struct
{
	gl_attrib* attribs;
	int n_attribs, cap_attribs;
	gl_buffer vertex;       /* the VAO's own Buffer struct (swgl.c:3053-3055) */
	gl_buffer element;
} gl_vao;

typedef struct
{
	uint32_t type;
	char* source;
	swgl_shader* compiled;  /* NULL until glCompileShader */
} gl_shader;

typedef struct { swgl_shader* sh; int var; } gl_varref;

typedef struct
{
	int linked;
	swgl_shader* vs;
	swgl_shader* fs;
	gl_varref uniforms[2 * SWGL_MAX_VARS];
	int n_uniforms;
	gl_varref layouts[2 * SWGL_MAX_VARS];
	int n_layouts;
	struct { int fs_in, vs_out; } pairs[SWGL_MAX_VARS];
	int n_pairs;
} gl_program;

typedef struct
{
	swgldev_ptr data;
	int32_t width, height, fpp, is_float;
	uint32_t wrap_s, wrap_t;
	swgldev_ptr mips;            /* glGenerateMipmap chain (swgldev_build_mipmaps), 0 = none */
	int32_t n_mipmaps;
} gl_texture;

#define VEC(T) struct { T** items; int n, cap; }

static struct
{
	swgldev_ctx* dev;
	int device_ordinal;          /* -1: default */
	int device_count;            /* devices the next glInit drives (swglSetDeviceCount), 0/1 = one */
	uint32_t width, height;

	VEC(gl_shader) shaders;
	VEC(gl_program) programs;
	VEC(gl_vao) vaos;
	VEC(gl_buffer) buffers;
	VEC(gl_texture) textures;

	gl_program* active_program;
	gl_vao* active_vao;
	gl_buffer* array_buffer;     /* GlobalArrayBuffer (swgl.c:3100) */
	gl_buffer* element_buffer;
	gl_texture* active_texture;
	gl_texture* texture_units[SWGL_MAX_TEX_UNITS];
	int active_unit;

	int32_t vx, vy;              /* these survive glInit, as the reference's globals do */
	uint32_t vw, vh;
	float clear_r, clear_g, clear_b, clear_a;

	char error[256];
} G = { .device_ordinal = -1 };

static void set_error(const char* msg)
{
	snprintf(G.error, sizeof(G.error), "%s", msg);
}

static void* vec_push(void*** items, int* n, int* cap, size_t elem)
{
	if (*n == *cap)
	{
		int ncap = *cap ? *cap * 2 : 16;
		void** ni = (void**)realloc(*items, sizeof(void*) * (size_t)ncap);
		if (!ni) return NULL;
		*items = ni;
		*cap = ncap;
	}
	void* obj = calloc(1, elem);
	if (!obj) return NULL;
	(*items)[(*n)++] = obj;
	return obj;
}
#define VEC_PUSH(v, T) ((T*)vec_push((void***)&(v).items, &(v).n, &(v).cap, sizeof(T)))
#define VEC_GET(v, i) (((i) >= 0 && (i) < (v).n) ? (v).items[(i)] : NULL)

static void free_tables(void)
{
	for (int i = 0; i < G.shaders.n; i++) { free(G.shaders.items[i]->source); /* compiled IR may be shared with programs: kept */ free(G.shaders.items[i]); }
	for (int i = 0; i < G.programs.n; i++) free(G.programs.items[i]);
	for (int i = 0; i < G.vaos.n; i++) { free(G.vaos.items[i]->attribs); free(G.vaos.items[i]); }
	for (int i = 0; i < G.buffers.n; i++) free(G.buffers.items[i]);
	for (int i = 0; i < G.textures.n; i++) free(G.textures.items[i]);
	free(G.shaders.items); free(G.programs.items); free(G.vaos.items); free(G.buffers.items); free(G.textures.items);
	memset(&G.shaders, 0, sizeof(G.shaders));
	memset(&G.programs, 0, sizeof(G.programs));
	memset(&G.vaos, 0, sizeof(G.vaos));
	memset(&G.buffers, 0, sizeof(G.buffers));
	memset(&G.textures, 0, sizeof(G.textures));
	G.active_program = NULL; G.active_vao = NULL; G.array_buffer = NULL; G.element_buffer = NULL;
	G.active_texture = NULL; G.active_unit = 0;
	memset(G.texture_units, 0, sizeof(G.texture_units));
}

/* ---------------------------------------------------------------------------------------- */
/* glInit (swgl.c:3713-3736): fresh object tables, colour + depth attachments, no clear, no
 * viewport.  The attachments are device memory here. */
void glInit(GLsizei width, GLsizei height)
{
	if (G.dev) { swgldev_destroy(G.dev); G.dev = NULL; }
	free_tables();
	G.width = width; G.height = height;
	int ordinal = G.device_ordinal;
	if (ordinal < 0)
	{
		const char* lr = getenv("LOCAL_RANK");
		ordinal = lr ? atoi(lr) : -1;
	}
	G.dev = G.device_count > 1 ? swgldev_create_group(ordinal < 0 ? 0 : ordinal, G.device_count, width, height)
	                           : swgldev_create(ordinal, width, height);
	if (!G.dev)
	{
		/* no CPU fallback: every later hot-path call is a no-op and the error is sticky */
		set_error("glInit: CUDA device context could not be created (no GPU, or out of memory)");
		fprintf(stderr, "swgl_b200: %s\n", G.error);
	}
}

uint32_t* glGetFramePtr(void)
{
	if (!G.dev) return NULL;
	return swgldev_map_color(G.dev);
}

float* swglGetDepthPtr(void)
{
	if (!G.dev) return NULL;
	return swgldev_map_depth(G.dev);
}

void swglFinish(void) { if (G.dev) swgldev_sync(G.dev); }

/* the step after the path */
uint64_t swglFrameSubmit(void) { return G.dev ? swgldev_frame_submit(G.dev) : 0; }
const uint32_t* swglFrameWait(uint64_t ticket) { return G.dev ? swgldev_frame_wait(G.dev, ticket) : NULL; }
void* swglHostAlloc(uint64_t bytes, int write_combined) { return swgldev_host_alloc(bytes, write_combined); }
void swglHostFree(void* p) { if (p) swgldev_host_free(p); }
int swglReadPixelsRGBA8(void* dst) { return (G.dev && dst) ? swgldev_read_rgba8(G.dev, dst) : -1; }

int swglWritePPM(const char* path)
{
	const uint32_t* px = glGetFramePtr();
	if (!px || !path) return -1;
	FILE* f = fopen(path, "wb");
	if (!f) { set_error("swglWritePPM: cannot open the file"); return -1; }
	fprintf(f, "P6\n%u %u\n255\n", (unsigned)G.width, (unsigned)G.height);
	uint8_t* row = (uint8_t*)malloc((size_t)G.width * 3u + 3u);
	int rc = row ? 0 : -1;
	for (uint32_t y = 0; row && y < (uint32_t)G.height; y++)
	{
		const uint32_t* src = px + (size_t)y * (size_t)G.width;
		for (uint32_t x = 0; x < (uint32_t)G.width; x++)
		{
			row[3u * x] = (uint8_t)(src[x] >> 24); row[3u * x + 1u] = (uint8_t)(src[x] >> 16); row[3u * x + 2u] = (uint8_t)(src[x] >> 8);
		}
		if (fwrite(row, 3u, (size_t)G.width, f) != (size_t)G.width) { rc = -1; break; }
	}
	free(row);
	if (fclose(f) != 0) rc = -1;
	if (rc) set_error("swglWritePPM: write failed");
	return rc;
}

uint64_t swglHashWords(const void* words, uint64_t n_words)
{
	const uint32_t* w = (const uint32_t*)words;
	uint64_t h = 1469598103934665603ull;
	for (uint64_t i = 0; i < n_words; i++) h = (h ^ (uint64_t)w[i]) * 1099511628211ull;
	return h;
}

const char* swglGetLastError(void)
{
	static char out[512];
	out[0] = 0;
	if (G.error[0]) { snprintf(out, sizeof(out), "%s", G.error); G.error[0] = 0; return out; }
	if (G.dev)
	{
		const char* e = swgldev_last_error(G.dev);
		if (e && e[0]) { snprintf(out, sizeof(out), "%s", e); return out; }
	}
	return out;
}

void swglGetStats(swglStats* out)
{
	if (!out) return;
	memset(out, 0, sizeof(*out));
	if (!G.dev) return;
	swgldev_stats s;
	swgldev_get_stats(G.dev, &s);
	out->draws = s.draws; out->triangles_in = s.triangles_in; out->prims_out = s.prims_out;
	out->tested = s.tested; out->shaded = s.shaded; out->tile_pairs = s.tile_pairs; out->bands = s.bands;
}

void swglSetDevice(int ordinal) { G.device_ordinal = ordinal; }
void swglSetDeviceCount(int count) { G.device_count = count; }
void* swglGetStream(void) { return G.dev ? swgldev_stream(G.dev) : NULL; }
uint64_t swglGetColorDevicePtr(void) { return G.dev ? swgldev_color_devptr(G.dev) : 0; }
uint64_t swglGetDepthDevicePtr(void) { return G.dev ? swgldev_depth_devptr(G.dev) : 0; }
void swglFillFramebuffer(uint32_t color_word, float depth) { if (G.dev) swgldev_fill(G.dev, color_word, depth); }
void swglSetStripe(GLuint rank, GLuint n_ranks, GLuint band_tile_rows) { if (G.dev) swgldev_set_stripe(G.dev, rank, n_ranks, band_tile_rows); }
void swglSetPeerColorTarget(uint64_t p) { if (G.dev) swgldev_set_peer_color(G.dev, p); }
int swglSetSharedFrameMirror(void* host_ptr, uint64_t bytes) { return G.dev ? swgldev_set_shared_mirror(G.dev, host_ptr, bytes) : -1; }
int swglIpcExportColor(void* handle64) { return G.dev ? swgldev_ipc_export_color(G.dev, handle64) : -1; }
uint64_t swglIpcOpen(const void* handle64) { return G.dev ? swgldev_ipc_open(G.dev, handle64) : 0; }
void swglIpcClose(uint64_t p) { if (G.dev) swgldev_ipc_close(G.dev, p); }
void swglSetOption(const char* name, int64_t value) { if (G.dev) swgldev_set_option(G.dev, name, value); }
int64_t swglGetOption(const char* name) { return swgldev_get_option(G.dev, name); }   /* without a device: -1, except the process-wide jit_* counters */

/* ---------------------------------------------------------------------------------------- */
/* shaders (swgl.c:2870-2898) */
GLuint glCreateShader(GLenum type)
{
	gl_shader* s = VEC_PUSH(G.shaders, gl_shader);
	if (!s) return 0;
	s->type = (uint32_t)type;
	return (GLuint)(G.shaders.n - 1); /* ids start at 0 */
}

void glShaderSource(GLuint shader, const GLchar* string)
{
	gl_shader* s = VEC_GET(G.shaders, (int)shader);
	if (!s || !string) return;
	free(s->source);
	size_t n = strlen(string);
	s->source = (char*)malloc(n + 1);
	if (s->source) memcpy(s->source, string, n + 1); /* copied: the caller may free it */
}

void glCompileShader(GLuint shader)
{
	gl_shader* s = VEC_GET(G.shaders, (int)shader);
	if (!s || !s->source) return;
	s->compiled = swgl_glsl_compile(s->source); /* a previous IR stays alive for attached programs */
	if (s->compiled && !s->compiled->ok)
	{
		char msg[256];
		snprintf(msg, sizeof(msg), "glCompileShader(%u): %s", shader, s->compiled->error);
		set_error(msg);
	}
}

void glDeleteShader(GLuint shader) { (void)shader; /* no-op in the reference (swgl.c:2895) */ }

int swglGetShaderCompiled(GLuint shader)
{
	gl_shader* s = VEC_GET(G.shaders, (int)shader);
	return s && s->compiled && s->compiled->ok;
}

size_t swglDebugShaderIR(GLuint shader, char* buf, size_t buf_len)
{
	gl_shader* s = VEC_GET(G.shaders, (int)shader);
	if (!s || !s->compiled) { if (buf && buf_len) buf[0] = 0; return 0; }
	return swgl_glsl_dump(s->compiled, buf, buf_len);
}

/* programs (swgl.c:2916-3014) */
GLuint glCreateProgram(void)
{
	gl_program* p = VEC_PUSH(G.programs, gl_program);
	if (!p) return 0;
	return (GLuint)G.programs.n; /* ids start at 1 */
}

void glAttachShader(GLuint program, GLuint shader)
{
	gl_program* p = VEC_GET(G.programs, (int)program - 1);
	gl_shader* s = VEC_GET(G.shaders, (int)shader);
	if (!p || !s) return;
	/* the reference copies the compiled data by value at attach time; its variables are
	 * pointers, so programs sharing a shader share its uniform storage */
	if (s->type == GL_VERTEX_SHADER) p->vs = s->compiled;
	if (s->type == GL_FRAGMENT_SHADER) p->fs = s->compiled;
}

void glLinkProgram(GLuint program)
{
	gl_program* p = VEC_GET(G.programs, (int)program - 1);
	if (!p || !p->vs || !p->fs) return;
	p->n_uniforms = 0; p->n_layouts = 0;
	/* uniforms and layouts: vertex shader globals first, then fragment (swgl.c:2958-2981) */
	swgl_shader* stages[2] = { p->vs, p->fs };
	for (int st = 0; st < 2; st++)
	{
		swgl_shader* sh = stages[st];
		for (int v = 0; v < sh->n_vars; v++)
		{
			if (sh->vars[v].is_local) continue;
			if (sh->vars[v].is_uniform) { p->uniforms[p->n_uniforms].sh = sh; p->uniforms[p->n_uniforms++].var = v; }
			if (sh->vars[v].is_layout) { p->layouts[p->n_layouts].sh = sh; p->layouts[p->n_layouts++].var = v; }
		}
	}
	/* FS `in` <-> first VS `out` of the same name (swgl.c:2983-3005); pairs accumulate on relink */
	for (int f = 0; f < p->fs->n_vars; f++)
	{
		if (p->fs->vars[f].is_local || !p->fs->vars[f].is_in) continue;
		for (int v = 0; v < p->vs->n_vars; v++)
		{
			if (p->vs->vars[v].is_local || !p->vs->vars[v].is_out) continue;
			if (!strcmp(p->fs->vars[f].name, p->vs->vars[v].name))
			{
				if (p->n_pairs < SWGL_MAX_VARS) { p->pairs[p->n_pairs].fs_in = f; p->pairs[p->n_pairs++].vs_out = v; }
				break;
			}
		}
	}
	p->linked = 1;
}

void glUseProgram(GLuint program)
{
	G.active_program = program == 0 ? NULL : VEC_GET(G.programs, (int)program - 1);
}

/* ---------------------------------------------------------------------------------------- */
/* vertex arrays and buffers (swgl.c:3043-3147) */
GLuint glGenVertexArrays(GLsizei n, GLuint* arrays)
{
	(void)n; /* one object per call */
	gl_vao* v = VEC_PUSH(G.vaos, gl_vao);
	if (v && arrays) *arrays = (GLuint)G.vaos.n;
	return 0;
}

void glBindVertexArray(GLuint array)
{
	G.active_vao = array == 0 ? NULL : VEC_GET(G.vaos, (int)array - 1);
}

void glVertexAttribPointer(GLuint index, GLint size, GLenum type, GLboolean normalized, GLsizei stride, const void* pointer)
{
	gl_vao* v = G.active_vao;
	if (!v) return;
	if (v->n_attribs == v->cap_attribs)
	{
		int ncap = v->cap_attribs ? v->cap_attribs * 2 : 8;
		gl_attrib* na = (gl_attrib*)realloc(v->attribs, sizeof(gl_attrib) * (size_t)ncap);
		if (!na) return;
		v->attribs = na; v->cap_attribs = ncap;
	}
	gl_attrib* a = &v->attribs[v->n_attribs++]; /* appended, never replaced (swgl.c:3080) */
	a->index = index; a->normalized = normalized; a->offset = (int32_t)(intptr_t)pointer;
	a->size = size; a->stride = stride; a->type = (uint32_t)type;
}

void glEnableVertexAttribArray(GLuint index) { (void)index; }

GLuint glGenBuffers(GLsizei n, GLuint* buffers)
{
	(void)n;
	gl_buffer* b = VEC_PUSH(G.buffers, gl_buffer);
	if (b && buffers) *buffers = (GLuint)G.buffers.n;
	return 0;
}

void glBindBuffer(GLenum type, GLuint buffer)
{
	gl_buffer** slot;
	gl_buffer* vao_own;
	if (type == GL_ARRAY_BUFFER) { slot = &G.array_buffer; vao_own = G.active_vao ? &G.active_vao->vertex : NULL; }
	else if (type == GL_ELEMENT_ARRAY_BUFFER) { slot = &G.element_buffer; vao_own = G.active_vao ? &G.active_vao->element : NULL; }
	else return;
	if (buffer == 0) { *slot = NULL; return; }
	gl_buffer* target = VEC_GET(G.buffers, (int)buffer - 1);
	if (!target) return;
	if (vao_own)
	{
		/* with a vertex array bound the VAO's own Buffer struct becomes the bound buffer and
		 * takes a snapshot of the named buffer's fields (swgl.c:3116-3122) */
		*vao_own = *target;
		vao_own->origin = target;
		*slot = vao_own;
	}
	else *slot = target;
}

static void scan_indices(gl_buffer* b, const void* data, uint32_t size)
{
	(void)data;
	b->max_index = swgldev_max_index(G.dev, b->data, size);   /* reduction on the device copy */
}

void glBufferData(GLenum target, GLsizei size, const void* data, GLenum usage)
{
	(void)usage;
	gl_buffer* b = target == GL_ARRAY_BUFFER ? G.array_buffer
	             : target == GL_ELEMENT_ARRAY_BUFFER ? G.element_buffer : NULL;
	if (!b || !G.dev || !data) return;
	if (b->size != 0) return; /* re-specification is ignored (swgl.c:3140) */
	if (size == 0) return;
	b->data = swgldev_alloc(G.dev, size);
	if (!b->data) return;
	b->size = size; b->capacity = size;
	swgldev_upload(G.dev, b->data, data, size);
	if (target == GL_ELEMENT_ARRAY_BUFFER) scan_indices(b, data, size);
}

void swglBufferRespecify(GLenum target, GLsizei size, const void* data)
{
	gl_buffer* b = target == GL_ARRAY_BUFFER ? G.array_buffer
	             : target == GL_ELEMENT_ARRAY_BUFFER ? G.element_buffer : NULL;
	if (!b || !G.dev || !data || size == 0) return;
	if (size > b->capacity)
	{
		/* The storage moves.  glBindBuffer snapshots a named buffer into the vertex array's own struct
		 * (swgl.c:3116-3122), so other vertex arrays -- and the named buffer -- may hold the old address: every
		 * struct that shares it follows the move (or is emptied when the allocation fails). */
		const swgldev_ptr old = b->data;
		if (old) swgldev_free(G.dev, old);
		b->data = swgldev_alloc(G.dev, size);
		b->capacity = b->data ? size : 0;
		if (!b->data) b->size = 0;
		if (old)
		{
			for (int i = 0; i < G.vaos.n; i++)
			{
				gl_vao* v = G.vaos.items[i];
				gl_buffer* both[2] = { &v->vertex, &v->element };
				for (int k = 0; k < 2; k++)
					if (both[k] != b && both[k]->data == old) { both[k]->data = b->data; both[k]->capacity = b->capacity; if (!b->data) both[k]->size = 0; }
			}
			for (int i = 0; i < G.buffers.n; i++)
			{
				gl_buffer* o = G.buffers.items[i];
				if (o != b && o->data == old) { o->data = b->data; o->capacity = b->capacity; if (!b->data) o->size = 0; }
			}
		}
		if (!b->data) return;
	}
	b->size = size;
	if (target == GL_ELEMENT_ARRAY_BUFFER) swgldev_upload_indices(G.dev, b->data, data, size, &b->max_index);
	else swgldev_upload_overlapped(G.dev, b->data, data, size);
	if (b->origin && ((gl_buffer*)b->origin)->size != 0)
	{
		/* the named buffer owns this storage too (specified before it was bound into the vertex
		 * array): keep it current, so that binding the name again does not resurrect stale fields */
		gl_buffer* o = (gl_buffer*)b->origin;
		o->data = b->data; o->size = b->size; o->capacity = b->capacity; o->max_index = b->max_index;
	}
}

static gl_buffer* bound_buffer(GLenum target)
{
	return target == GL_ARRAY_BUFFER ? G.array_buffer : target == GL_ELEMENT_ARRAY_BUFFER ? G.element_buffer : NULL;
}

static void sync_origin(gl_buffer* b)
{
	if (b->origin && ((gl_buffer*)b->origin)->size != 0)
	{
		gl_buffer* o = (gl_buffer*)b->origin;
		o->data = b->data; o->size = b->size; o->capacity = b->capacity; o->max_index = b->max_index;
	}
}

void swglBufferSubData(GLenum target, uint64_t offset, GLsizei size, const void* data)
{
	gl_buffer* b = bound_buffer(target);
	if (!b || !G.dev || !data || size == 0 || !b->data) return;
	if (offset + (uint64_t)size > (uint64_t)b->size) { set_error("swglBufferSubData: range outside the buffer"); return; }
	swgldev_upload_range(G.dev, b->data, offset, data, size);
	/* element data: the largest index is refreshed by swglBufferDeviceWritten once the whole buffer is in place */
}

uint64_t swglGetBufferDevicePtr(GLenum target)
{
	gl_buffer* b = bound_buffer(target);
	return (b && G.dev) ? (uint64_t)b->data : 0;
}

void swglBufferDeviceWritten(GLenum target)
{
	gl_buffer* b = bound_buffer(target);
	if (!b || !G.dev || !b->data || b->size == 0) return;
	if (target == GL_ELEMENT_ARRAY_BUFFER) b->max_index = swgldev_max_index_after_stream(G.dev, b->data, b->size);
	sync_origin(b);
}

/* ---------------------------------------------------------------------------------------- */
/* textures (swgl.c:2038-2173) */
void glGenTextures(GLsizei n, GLuint* textures)
{
	(void)n;
	gl_texture* t = VEC_PUSH(G.textures, gl_texture);
	if (!t) return;
	t->fpp = 3; t->wrap_s = GL_REPEAT; t->wrap_t = GL_REPEAT;
	if (textures) *textures = (GLuint)G.textures.n;
}

void glBindTexture(GLenum target, GLuint texture)
{
	if (target != GL_TEXTURE_2D) return;
	if (texture == 0) { G.active_texture = NULL; return; } /* the unit keeps its texture (swgl.c:2056-2059) */
	gl_texture* t = VEC_GET(G.textures, (int)texture - 1);
	if (!t) return;
	G.active_texture = t;
	if (G.active_unit >= 0 && G.active_unit < SWGL_MAX_TEX_UNITS) G.texture_units[G.active_unit] = t;
}

void glActiveTexture(GLenum target) { G.active_unit = (int)target - (int)GL_TEXTURE0; }

void glTexParameteri(GLenum target, GLenum type, GLenum mode)
{
	if (target != GL_TEXTURE_2D || !G.active_texture) return;
	if (type == GL_TEXTURE_WRAP_S) G.active_texture->wrap_s = (uint32_t)mode;
	if (type == GL_TEXTURE_WRAP_T) G.active_texture->wrap_t = (uint32_t)mode;
}

void glTexImage2D(GLenum target, GLint level, GLint internalformat, GLsizei width, GLsizei height, GLint border, GLenum format, GLenum type, const void* data)
{
	(void)level;
	if (!data || border != 0 || target != GL_TEXTURE_2D || !G.active_texture || !G.dev) return;
	gl_texture* t = G.active_texture;
	if (t->data) { swgldev_free(G.dev, t->data); t->data = 0; }
	/* the chain stays: the reference never clears MipMaps (swgl.c:2094-2118) and goes on sampling the old image's
	 * levels, with the new image's floats per texel, until glGenerateMipmap appends new ones behind them */
	if (internalformat != (GLint)format) return; /* the reference returns here with the old data freed */
	if (internalformat == GL_RGBA) t->fpp = 4;
	if (internalformat == GL_RGB) t->fpp = 3;
	if (internalformat == GL_RG) t->fpp = 2;
	if (internalformat == GL_RED) t->fpp = 1;
	if (type != GL_FLOAT && type != GL_UNSIGNED_BYTE) { set_error("glTexImage2D: only GL_FLOAT and GL_UNSIGNED_BYTE texel data"); return; }
	t->width = (int32_t)width; t->height = (int32_t)height;
	t->is_float = type == GL_FLOAT;
	/* Texels stay in their upload format; bytes are converted with /255.0f when sampled, which
	 * yields the float the reference stores at upload (swgl.c:2116) from 4x fewer bytes. */
	uint64_t bytes = (uint64_t)t->fpp * width * height * (t->is_float ? 4u : 1u);
	if (bytes == 0) return;
	t->data = swgldev_alloc(G.dev, bytes);
	if (t->data) swgldev_upload(G.dev, t->data, data, bytes);
}

void glGenerateMipmap(GLenum target)
{
	if (target != GL_TEXTURE_2D || !G.active_texture || !G.active_texture->data || !G.dev) return;
	/* The reference builds a 2x2-box chain here (swgl.c:2129-2171) that is only read when
	 * MipMapLevel > 0 (swgl.c:2527).  MipMapLevel is 40 / DistBetweenPointAndLine(...) and that
	 * distance is multiplied by rsqrt(), whose 8-byte pun of a 4-byte float (swgl.c:3246) folds
	 * the low bit of the adjacent stack word -- rsqrt's own return address -- into the sign:
	 * in the compiled reference (gcc -O2, oracle/_ref) that bit is 1, rsqrt is negative for every
	 * input, MipMapLevel is never positive and the base level is always sampled.  That is the
	 * DEFAULT here too (tests/test_next_rows_gpu.py compares against the compiled reference).
	 * The chain is built all the same (k_mipmap_box); swglSetOption("mip_lod", 1) selects the
	 * defined variant -- rsqrt with a 32-bit pun -- in which the chain is sampled with the
	 * per-triangle level (tests/test_mipmap_gpu.py, against the reference built the same way). */
	gl_texture* t = G.active_texture;
	swgldev_texture base;
	memset(&base, 0, sizeof(base));
	base.data = t->data; base.width = t->width; base.height = t->height; base.fpp = t->fpp; base.is_float = t->is_float;
	/* a chain that is already there is kept and the new levels go behind it (swgl.c:2134-2166 pushes onto the same
	 * vector): the device layer builds old + new as one allocation */
	base.mips = t->mips; base.n_mips = t->n_mipmaps;
	int32_t n = 0;
	const swgldev_ptr chain = swgldev_build_mipmaps(G.dev, &base, &n);
	if (!chain) return;             /* nothing new (image too small, table full) or an error: what was there stays */
	if (t->mips) swgldev_free(G.dev, t->mips);
	t->mips = chain; t->n_mipmaps = n;
}

/* ---------------------------------------------------------------------------------------- */
/* uniforms (swgl.c:3743-3928) */
GLint glGetUniformLocation(GLuint program, const GLchar* name)
{
	gl_program* p = VEC_GET(G.programs, (int)program - 1);
	if (!p || !name) return -1;
	for (int i = 0; i < p->n_uniforms; i++)
		if (!strcmp(p->uniforms[i].sh->vars[p->uniforms[i].var].name, name))
			return (GLint)(((program - 1) << 16) | (uint32_t)i);
	return -1;
}

static uint32_t* uniform_words(GLint location, int* type)
{
	if (location < 0) return NULL; /* -1 is undefined behaviour in the reference; guarded */
	gl_program* p = VEC_GET(G.programs, location >> 16);
	if (!p) return NULL;
	int i = location & 0xFFFF;
	if (i >= p->n_uniforms) return NULL;
	swgl_shader* sh = p->uniforms[i].sh;
	const swgl_var* v = &sh->vars[p->uniforms[i].var];
	*type = v->type;
	return &sh->image[v->word];
}

static void put_floats(GLint location, int want_type, const float* f, int n)
{
	int type;
	uint32_t* w = uniform_words(location, &type);
	if (!w || type != want_type) return; /* AssignToExVal ignores a type mismatch */
	memcpy(w, f, sizeof(float) * (size_t)n);
}

void glUniform1f(GLint location, GLfloat v0) { float f[1] = { v0 }; put_floats(location, SWT_FLOAT, f, 1); }
void glUniform2f(GLint location, GLfloat v0, GLfloat v1) { float f[2] = { v0, v1 }; put_floats(location, SWT_VEC2, f, 2); }
void glUniform3f(GLint location, GLfloat v0, GLfloat v1, GLfloat v2) { float f[3] = { v0, v1, v2 }; put_floats(location, SWT_VEC3, f, 3); }
void glUniform4f(GLint location, GLfloat v0, GLfloat v1, GLfloat v2, GLfloat v3) { float f[4] = { v0, v1, v2, v3 }; put_floats(location, SWT_VEC4, f, 4); }

void glUniform1i(GLint location, GLint v0)
{
	int type;
	uint32_t* w = uniform_words(location, &type);
	if (!w || (type != SWT_INT && type != SWT_SAMPLER2D)) return;
	memcpy(w, &v0, 4);
}

/* `transpose = !transpose`: GL_FALSE stores the transposed array (swgl.c:3846, 3869, 3897) */
static void put_matrix(GLint location, int dim, GLboolean transpose, const GLfloat* value)
{
	int type;
	uint32_t* w = uniform_words(location, &type);
	if (!w || !value) return;
	if (swt_words(type) < dim * dim) return; /* the reference would overflow the variable here */
	float m[16];
	if (transpose) memcpy(m, value, sizeof(float) * (size_t)(dim * dim));
	else
		for (int r = 0; r < dim; r++)
			for (int c = 0; c < dim; c++)
				m[r * dim + c] = value[c * dim + r];
	memcpy(w, m, sizeof(float) * (size_t)(dim * dim));
}

void glUniformMatrix2fv(GLint location, GLsizei count, GLboolean transpose, const GLfloat* value) { (void)count; put_matrix(location, 2, transpose, value); }
void glUniformMatrix3fv(GLint location, GLsizei count, GLboolean transpose, const GLfloat* value) { (void)count; put_matrix(location, 3, transpose, value); }
void glUniformMatrix4fv(GLint location, GLsizei count, GLboolean transpose, const GLfloat* value) { (void)count; put_matrix(location, 4, transpose, value); }

/* ---------------------------------------------------------------------------------------- */
/* clear / viewport (swgl.c:3175-3222) */
void glClearColor(GLfloat red, GLfloat green, GLfloat blue, GLfloat alpha)
{
	G.clear_r = RMIN(RMAX(red, 0.0f), 1.0f);
	G.clear_g = RMIN(RMAX(green, 0.0f), 1.0f);
	G.clear_b = RMIN(RMAX(blue, 0.0f), 1.0f);
	G.clear_a = RMIN(RMAX(alpha, 0.0f), 1.0f);
}

void glViewport(GLint x, GLint y, GLsizei width, GLsizei height)
{
	G.vx = x; G.vy = y; G.vw = width; G.vh = height;
}

void glClear(GLuint flags)
{
	if (!G.dev) return;
	uint32_t word = 0;
	word |= (uint32_t)(G.clear_r * 255) << 24;
	word |= (uint32_t)(G.clear_g * 255) << 16;
	word |= (uint32_t)(G.clear_b * 255) << 8;
	word |= (uint32_t)(G.clear_a * 255);
	/* rows MAX(VY,0) .. MIN(VY+VH, Height): the MIN is unsigned (swgl.c:3193-3195); no Y flip */
	int32_t y0 = RMAX(G.vy, 0), x0 = RMAX(G.vx, 0);
	uint32_t y1u = RMIN((uint32_t)G.vy + G.vh, G.height);
	uint32_t x1u = RMIN((uint32_t)G.vx + G.vw, G.width);
	swgldev_clear(G.dev, flags & 3u, word, x0, y0, (int32_t)x1u, (int32_t)y1u);
}

/* ---------------------------------------------------------------------------------------- */
/* draw (swgl.c:3475-3711): assemble the by-value draw description and hand it to the device */
static int vec_floats(int type)
{
	return type == SWT_FLOAT ? 1 : type == SWT_VEC2 ? 2 : type == SWT_VEC3 ? 3 : type == SWT_VEC4 ? 4 : 0;
}

static int find_fetch_for_word(const swgldev_draw* d, uint32_t word)
{
	int found = -1;
	for (uint32_t i = 0; i < d->n_fetch; i++) if (d->fetch[i].dst_word == word) found = (int)i; /* last wins */
	return found;
}

/* Everything a draw needs, gathered from the GL state into `d` (returns 0 when nothing is to be drawn).
 * compile_only: the caller only wants the shader interface (swglPrecompileProgram): no device, buffer
 * contents or count are required. */
static int assemble_draw(GLenum mode, GLint first, GLsizei count, int indexed, uint32_t index_byte_offset, swgldev_draw* dp, int compile_only)
{
#define d (*dp)
	if (!G.active_vao) return 0;     /* swgl.c:3477-3478 */
	if (!G.active_program) return 0;
	if (!G.dev && !compile_only) return 0;
	if (mode != GL_TRIANGLES && mode != GL_POINTS)
	{
		/* GL_LINES has an enumerator but no implementation in the reference either (swgl.h:65) */
		return 0;
	}
	gl_program* p = G.active_program;
	if (!p->linked || !p->vs || !p->fs || !p->vs->ok || !p->fs->ok)
	{
		set_error("glDraw*: program is not linked or a shader is outside the executable subset; nothing drawn");
		return 0;
	}
	gl_vao* vao = G.active_vao;
	if (!compile_only && (!vao->vertex.data || count == 0)) return 0;

	memset(&d, 0, sizeof(d));
	d.vx = G.vx; d.vy = G.vy; d.vw = G.vw; d.vh = G.vh;
	d.vbo = vao->vertex.data; d.vbo_bytes = vao->vertex.size;
	d.first = first; d.count = count;
	if (indexed)
	{
		if (!vao->element.data) return 0;
		if (index_byte_offset & 3u) { set_error("glDrawElements: index offset must be a multiple of 4"); return 0; }
		d.ibo = vao->element.data; d.ibo_bytes = vao->element.size;
		d.first = (int32_t)(index_byte_offset / 4u);
		d.n_vertices = vao->element.max_index + 1u;
	}

	swgl_shader* vs = p->vs;
	swgl_shader* fs = p->fs;
	d.vs_code = &vs->code; d.fs_code = &fs->code;
	d.vs_code_id = vs->id; d.fs_code_id = fs->id;
	d.vs_image = vs->image; d.vs_words = vs->code.n_words;
	d.fs_image = fs->image; d.fs_words = fs->code.n_words;
	d.pos_word = vs->vars[0].word; /* gl_Position is always variable 0 */

	/* attribute fetch table: VAO attributes in order x layout variables with that location */
	for (int a = 0; a < vao->n_attribs; a++)
	{
		const gl_attrib* at = &vao->attribs[a];
		if (at->type != GL_FLOAT) continue; /* other types are skipped silently (swgl.c:3624) */
		for (int l = 0; l < p->n_layouts; l++)
		{
			if (p->layouts[l].sh != vs) continue; /* fragment-stage layout variables are never observable */
			const swgl_var* lv = &vs->vars[p->layouts[l].var];
			if (lv->location != (int32_t)at->index) continue;
			if (d.n_fetch >= SWGL_MAX_FETCH) { set_error("glDraw*: too many attribute bindings"); return 0; }
			int n = at->size;
			if (n > swt_words(lv->type)) n = swt_words(lv->type); /* the reference overflows the variable here */
			if (n < 0) n = 0;
			if (at->offset < 0) { set_error("glDraw*: negative attribute offset"); return 0; }
			swgldev_fetch* f = &d.fetch[d.n_fetch++];
			f->src_offset = (uint32_t)at->offset; f->stride = at->stride; f->n_floats = (uint32_t)n; f->dst_word = lv->word;
		}
	}

	/* linked varyings, fragment-`in` order; a type mismatch leaves the `in` unfed (swgl.c:3656) */
	uint32_t slot = 0;
	int vs_fast = vs->simple_copies && vs->pos_kind != 0;
	for (int k = 0; k < p->n_pairs; k++)
	{
		const swgl_var* fi = &fs->vars[p->pairs[k].fs_in];
		const swgl_var* vo = &vs->vars[p->pairs[k].vs_out];
		if (fi->type != vo->type) continue;
		int n = vec_floats(fi->type);
		if (n == 0) continue; /* int varyings interpolate to an unspecified value in the reference */
		if (n == 4) slot = (slot + 3u) & ~3u;   /* vec4 varyings on 16-byte boundaries of the packed record: 128-bit loads */
		if (d.n_varying >= 8 || slot + (uint32_t)n > SWGL_MAX_VARYING_FLOATS) { set_error("glDraw*: too many varyings"); return 0; }
		swgldev_varying* v = &d.varying[d.n_varying++];
		v->vs_word = vo->word; v->fs_word = fi->word; v->n_floats = (uint32_t)n; v->slot = slot;
		slot += (uint32_t)n;
		if (vo->copy_of >= 0 && vs->vars[vo->copy_of].is_layout)
		{
			int fe = find_fetch_for_word(&d, vs->vars[vo->copy_of].word);
			if (fe >= 0) { v->src_offset = d.fetch[fe].src_offset; v->src_stride = d.fetch[fe].stride; v->src_floats = d.fetch[fe].n_floats; }
		}
		else vs_fast = 0;
	}
	for (uint32_t k = 0; k < d.n_varying; k++) if (d.varying[k].n_floats == 4) { slot = (slot + 3u) & ~3u; break; }   /* ... of every record */
	d.varying_floats = slot;

	/* vertex shape */
	d.vs_kind = SWVS_GENERIC;
	if (vs_fast && vs->pos_attr_var >= 0 && vs->vars[vs->pos_attr_var].is_layout && vs->vars[vs->pos_attr_var].type == SWT_VEC4)
	{
		int fe = find_fetch_for_word(&d, vs->vars[vs->pos_attr_var].word);
		if (fe >= 0) { d.pos_src_offset = d.fetch[fe].src_offset; d.pos_src_stride = d.fetch[fe].stride; d.pos_src_floats = d.fetch[fe].n_floats; }
		if (vs->pos_kind == 1) d.vs_kind = SWVS_PASS;
		else if (vs->pos_kind == 2 && vs->pos_mat_var >= 0)
		{
			/* mat4 variable load: the fourth row starts at Data[10] (swgl.c:2231-2235) */
			const float* D = (const float*)&vs->image[vs->vars[vs->pos_mat_var].word];
			memcpy(d.pos_matrix, D, sizeof(float) * 12);
			d.pos_matrix[12] = D[10]; d.pos_matrix[13] = D[13]; d.pos_matrix[14] = D[14]; d.pos_matrix[15] = D[15];
			d.vs_kind = SWVS_MATRIX;
		}
	}

	/* fragment output: the first `out` global, four floats (swgl.c:3412-3426) */
	int out_var = -1;
	for (int v = 0; v < fs->n_vars; v++) if (!fs->vars[v].is_local && fs->vars[v].is_out) { out_var = v; break; }
	if (out_var < 0) { set_error("glDraw*: fragment shader has no `out` variable"); return 0; }
	d.out_word = fs->vars[out_var].word;
	d.out_floats = (uint32_t)RMIN(4, swt_words(fs->vars[out_var].type));

	/* fragment shape */
	d.fs_kind = SWFS_GENERIC;
	if (fs->fs_kind != SWFS_GENERIC && fs->fs_in_var >= 0)
	{
		for (uint32_t k = 0; k < d.n_varying; k++)
			if (d.varying[k].fs_word == fs->vars[fs->fs_in_var].word)
			{
				d.fs_kind = fs->fs_kind;
				d.fs_slot = d.varying[k].slot; d.fs_slot_floats = d.varying[k].n_floats;
			}
		if (d.fs_kind == SWFS_TEXTURE)
		{
			d.fs_swz_u = (uint32_t)fs->fs_swz[0]; d.fs_swz_v = (uint32_t)fs->fs_swz[1];
			int32_t unit;
			memcpy(&unit, &fs->image[fs->vars[fs->fs_sampler_var].word], 4);
			d.fs_tex_unit = unit;
			if (unit < 0 || unit >= SWGL_MAX_TEX_UNITS) d.fs_kind = SWFS_GENERIC;
		}
	}

	for (int u = 0; u < SWGL_MAX_TEX_UNITS; u++)
	{
		const gl_texture* t = G.texture_units[u];
		if (!t || !t->data) continue;
		d.tex[u].data = t->data; d.tex[u].width = t->width; d.tex[u].height = t->height;
		d.tex[u].fpp = t->fpp; d.tex[u].is_float = t->is_float;
		d.tex[u].wrap_s_repeat = t->wrap_s == GL_REPEAT; d.tex[u].wrap_t_repeat = t->wrap_t == GL_REPEAT;
		d.tex[u].mips = t->mips; d.tex[u].n_mips = t->n_mipmaps;
	}

	return 1;
#undef d
}

static void draw_common(GLenum mode, GLint first, GLsizei count, int indexed, uint32_t index_byte_offset)
{
	swgldev_draw d;
	if (!assemble_draw(mode, first, count, indexed, index_byte_offset, &d, 0)) return;
	if (mode == GL_POINTS) swgldev_draw_points(G.dev, &d);
	else swgldev_draw_triangles(G.dev, &d);
}

/* Compile the kernels the bound program + vertex array need now instead of at the first draw (a shader
 * outside the built-in shapes is compiled at run time, swgl_jit.cpp).  Works without a device: the code is
 * then generated and compiled, not loaded. */
int swglPrecompileProgram(void)
{
	swgldev_draw d;
	if (!assemble_draw(GL_TRIANGLES, 0, 3, 0, 0, &d, 1)) return -1;
	char msg[2048];
	msg[0] = 0;
	if (swgldev_precompile(G.dev, &d, msg, sizeof(msg))) { set_error(msg); return -1; }
	return 0;
}

void glDrawArrays(GLenum mode, GLint first, GLsizei count)
{
	draw_common(mode, first, count, 0, 0);
}

void glDrawElements(GLenum mode, GLsizei count, GLenum type, const void* indices)
{
	if (type != GL_UNSIGNED_INT) { set_error("glDrawElements: only GL_UNSIGNED_INT indices"); return; }
	draw_common(mode, 0, count, 1, (uint32_t)(uintptr_t)indices);
}
