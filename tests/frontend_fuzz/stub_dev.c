/* tests/frontend_fuzz/stub_dev.c -- test infrastructure: a stand-in for the CUDA device layer (include/swgl_dev.h) with
 * which the C host layer (swgl_host.c + swgl_glsl.c) is built under AddressSanitizer / UBSan on a box without a GPU
 * (tests/test_host_sanitizers.py).  "Device memory" is malloc'ed host memory; nothing is rendered.  What it does do is
 * READ every byte a draw hands it -- vertex and element buffers, textures and their mip chains, the variable files, the
 * IR code pointers -- so that a pointer the host layer kept past a free, or a size that does not match its allocation,
 * is reported by the sanitizer at the draw that would have used it on the GPU. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "swgl_dev.h"
#include "swgl_ir.h"

struct swgldev_ctx
{
	uint32_t W, H;
	uint32_t* color;
	float* depth;
	char err[64];
	swgldev_stats stats;
	uint64_t ticket;
};

static volatile unsigned long long g_sink;

static void touch(const void* p, uint64_t bytes)
{
	const unsigned char* q = (const unsigned char*)p;
	unsigned long long s = 0;
	for (uint64_t i = 0; i < bytes; i++) s += q[i];
	g_sink += s;
}

swgldev_ctx* swgldev_create(int device, uint32_t width, uint32_t height)
{
	(void)device;
	swgldev_ctx* c = (swgldev_ctx*)calloc(1, sizeof(*c));
	if (!c) return NULL;
	c->W = width; c->H = height;
	c->color = (uint32_t*)calloc((size_t)width * height + 1, 4);
	c->depth = (float*)calloc((size_t)width * height + 1, 4);
	return c;
}
swgldev_ctx* swgldev_create_group(int device, int count, uint32_t width, uint32_t height) { (void)count; return swgldev_create(device, width, height); }
void swgldev_destroy(swgldev_ctx* c) { if (!c) return; free(c->color); free(c->depth); free(c); }
const char* swgldev_last_error(swgldev_ctx* c) { (void)c; return ""; }
void* swgldev_stream(swgldev_ctx* c) { (void)c; return NULL; }
int swgldev_sync(swgldev_ctx* c) { (void)c; return 0; }

/* allocations carry their size in front, so that uploads and draws can be checked against it */
swgldev_ptr swgldev_alloc(swgldev_ctx* c, uint64_t bytes)
{
	(void)c;
	uint64_t* p = (uint64_t*)malloc(16 + (size_t)bytes);
	if (!p) return 0;
	p[0] = bytes; p[1] = 0x5157474cu;
	memset(p + 2, 0, (size_t)bytes);
	return (swgldev_ptr)(uintptr_t)(p + 2);
}
static uint64_t size_of(swgldev_ptr p)
{
	const uint64_t* q = (const uint64_t*)(uintptr_t)p - 2;
	if (q[1] != 0x5157474cu) { fprintf(stderr, "stub device: not the start of an allocation\n"); abort(); }
	return q[0];
}
void swgldev_free(swgldev_ctx* c, swgldev_ptr p) { (void)c; if (p) { size_of(p); free((uint64_t*)(uintptr_t)p - 2); } }
static int put(swgldev_ptr dst, uint64_t offset, const void* src, uint64_t bytes)
{
	if (offset + bytes > size_of(dst)) { fprintf(stderr, "stub device: upload of %llu bytes at %llu into %llu\n", (unsigned long long)bytes, (unsigned long long)offset, (unsigned long long)size_of(dst)); abort(); }
	memcpy((char*)(uintptr_t)dst + offset, src, (size_t)bytes);
	return 0;
}
int swgldev_upload(swgldev_ctx* c, swgldev_ptr dst, const void* src, uint64_t bytes) { (void)c; return put(dst, 0, src, bytes); }
int swgldev_upload_overlapped(swgldev_ctx* c, swgldev_ptr dst, const void* src, uint64_t bytes) { (void)c; return put(dst, 0, src, bytes); }
int swgldev_upload_range(swgldev_ctx* c, swgldev_ptr base, uint64_t offset, const void* src, uint64_t bytes) { (void)c; return put(base, offset, src, bytes); }
uint32_t swgldev_max_index(swgldev_ctx* c, swgldev_ptr indices, uint64_t bytes)
{
	(void)c;
	if (bytes > size_of(indices)) { fprintf(stderr, "stub device: index scan past the buffer\n"); abort(); }
	uint32_t m = 0;
	const uint32_t* ix = (const uint32_t*)(uintptr_t)indices;
	for (uint64_t i = 0; i < bytes / 4; i++) if (ix[i] > m) m = ix[i];
	return m;
}
uint32_t swgldev_max_index_after_stream(swgldev_ctx* c, swgldev_ptr indices, uint64_t bytes) { return swgldev_max_index(c, indices, bytes); }
int swgldev_upload_indices(swgldev_ctx* c, swgldev_ptr dst, const void* src, uint64_t bytes, uint32_t* max_index)
{
	put(dst, 0, src, bytes);
	if (max_index) *max_index = swgldev_max_index(c, dst, bytes);
	return 0;
}

swgldev_ptr swgldev_build_mipmaps(swgldev_ctx* c, const swgldev_texture* base, int32_t* n_levels)
{
	*n_levels = 0;
	if (!base->data) return 0;
	touch((const void*)(uintptr_t)base->data, (uint64_t)base->width * base->height * base->fpp * (base->is_float ? 4u : 1u));
	if (base->mips) touch((const void*)(uintptr_t)base->mips, size_of(base->mips));
	int cw = base->width / 2, ch = base->height / 2, n = base->mips ? base->n_mips : 0;
	const int before = n;
	while (cw + ch > 4 && n < SWGL_MIP_MAX_LEVELS) { n++; cw /= 2; ch /= 2; }
	if (n == before) return 0;
	*n_levels = n;
	return swgldev_alloc(c, 4 * SWGL_MIP_HEADER_WORDS + 64);
}

int swgldev_clear(swgldev_ctx* c, uint32_t flags, uint32_t color_word, int32_t x0, int32_t y0, int32_t x1, int32_t y1)
{
	(void)flags; (void)color_word;
	if (x0 < 0 || y0 < 0 || x1 > (int32_t)c->W || y1 > (int32_t)c->H) { fprintf(stderr, "stub device: clear rectangle outside the framebuffer\n"); abort(); }
	return 0;
}

static int draw(swgldev_ctx* c, const swgldev_draw* d)
{
	if (d->vbo) { if (d->vbo_bytes > size_of(d->vbo)) { fprintf(stderr, "stub device: vbo_bytes beyond the allocation\n"); abort(); } touch((const void*)(uintptr_t)d->vbo, d->vbo_bytes); }
	if (d->ibo) { if (d->ibo_bytes > size_of(d->ibo)) { fprintf(stderr, "stub device: ibo_bytes beyond the allocation\n"); abort(); } touch((const void*)(uintptr_t)d->ibo, d->ibo_bytes); }
	if (d->vs_image) touch(d->vs_image, 4ull * d->vs_words);
	if (d->fs_image) touch(d->fs_image, 4ull * d->fs_words);
	if (d->vs_code) touch(d->vs_code, 8);
	if (d->fs_code) touch(d->fs_code, 8);
	if (d->n_fetch > 16 || d->n_varying > 8) { fprintf(stderr, "stub device: table sizes\n"); abort(); }
	for (int u = 0; u < 8; u++)
	{
		const swgldev_texture* t = &d->tex[u];
		if (!t->data) continue;
		const uint64_t bytes = (uint64_t)t->width * t->height * t->fpp * (t->is_float ? 4u : 1u);
		if (bytes > size_of(t->data)) { fprintf(stderr, "stub device: texture larger than its allocation\n"); abort(); }
		touch((const void*)(uintptr_t)t->data, bytes);
		if (t->mips) touch((const void*)(uintptr_t)t->mips, size_of(t->mips));
	}
	c->stats.draws++;
	c->stats.triangles_in = (d->count + 2) / 3;
	return 0;
}
int swgldev_draw_triangles(swgldev_ctx* c, const swgldev_draw* d) { return draw(c, d); }
int swgldev_draw_points(swgldev_ctx* c, const swgldev_draw* d) { return draw(c, d); }
#ifdef STUB_WITH_JIT
/* the real code generator + NVRTC (swgl_jit.cpp, built with the sanitizers too): generate and compile, load nothing */
#include "swgl_jit.h"
int swgldev_precompile(swgldev_ctx* c, const swgldev_draw* d, char* msg, size_t msg_len)
{
	(void)c;
	if (msg_len) msg[0] = 0;
	const int want_v = d->vs_kind == SWVS_GENERIC, want_r = d->fs_kind == SWFS_GENERIC;
	if (!want_v && !want_r) return 0;
	swgljit_kernels k;
	return swgljit_get(-1, d, want_v, want_r, &k, msg, msg_len);
}
#else
int swgldev_precompile(swgldev_ctx* c, const swgldev_draw* d, char* msg, size_t msg_len) { (void)c; (void)d; if (msg_len) msg[0] = 0; return 0; }
#endif

uint32_t* swgldev_map_color(swgldev_ctx* c) { return c->color; }
float* swgldev_map_depth(swgldev_ctx* c) { return c->depth; }
uint64_t swgldev_frame_submit(swgldev_ctx* c) { return ++c->ticket; }
const uint32_t* swgldev_frame_wait(swgldev_ctx* c, uint64_t ticket) { return (ticket && ticket <= c->ticket) ? c->color : NULL; }
int swgldev_read_rgba8(swgldev_ctx* c, void* dst) { memcpy(dst, c->color, (size_t)c->W * c->H * 4); return 0; }
void* swgldev_host_alloc(uint64_t bytes, int write_combined) { (void)write_combined; return malloc((size_t)bytes); }
void swgldev_host_free(void* p) { free(p); }
swgldev_ptr swgldev_color_devptr(swgldev_ctx* c) { return (swgldev_ptr)(uintptr_t)c->color; }
swgldev_ptr swgldev_depth_devptr(swgldev_ctx* c) { return (swgldev_ptr)(uintptr_t)c->depth; }
void swgldev_fill(swgldev_ctx* c, uint32_t color_word, float depth) { for (size_t i = 0; i < (size_t)c->W * c->H; i++) { c->color[i] = color_word; c->depth[i] = depth; } }
void swgldev_get_stats(swgldev_ctx* c, swgldev_stats* out) { *out = c->stats; }
void swgldev_set_stripe(swgldev_ctx* c, uint32_t rank, uint32_t n_ranks, uint32_t band_tile_rows) { (void)c; (void)rank; (void)n_ranks; (void)band_tile_rows; }
void swgldev_set_peer_color(swgldev_ctx* c, swgldev_ptr peer_color) { (void)c; (void)peer_color; }
int swgldev_set_shared_mirror(swgldev_ctx* c, void* host_ptr, uint64_t bytes) { (void)c; (void)host_ptr; (void)bytes; return 0; }
int swgldev_ipc_export_color(swgldev_ctx* c, void* handle64) { (void)c; memset(handle64, 0, 64); return 0; }
swgldev_ptr swgldev_ipc_open(swgldev_ctx* c, const void* handle64) { (void)c; (void)handle64; return 0; }
void swgldev_ipc_close(swgldev_ctx* c, swgldev_ptr p) { (void)c; (void)p; }
void swgldev_set_option(swgldev_ctx* c, const char* name, int64_t value) { (void)c; (void)name; (void)value; }
int64_t swgldev_get_option(swgldev_ctx* c, const char* name) { (void)c; (void)name; return 0; }
