/*
 * include/swgl_dev.h -- the thin C-ABI device layer under the GL-style entry points.
 *
 * The host side of libswgl_b200.so is plain C (swgl_host.c + swgl_glsl.c): it owns the GL
 * object tables and the shader front-end and reaches CUDA only through the functions below
 * (plain pointers, sizes and POD structs; no C++ or torch types).  Each function names the
 * part of the reference's draw path it replaces.
 *
 * Applications normally use swgl.h; this header is what a maintainer binds when wiring the
 * reference's own C file to the CUDA path call by call.
 */
#ifndef SWGL_DEV_H
#define SWGL_DEV_H

#ifndef __CUDACC_RTC__   /* NVRTC (run-time compiled shader kernels) brings its own fixed-width types */
#include <stdint.h>
#include <stddef.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct swgldev_ctx swgldev_ctx;   /* one device context: framebuffer + scratch + stream */
typedef uint64_t swgldev_ptr;             /* device address */

/* One entry per (VAO attribute, matching layout variable) pair, GL_FLOAT attributes only
 * (reference: attribute fetch, swgl.c:3618-3639). */
typedef struct
{
	uint32_t src_offset;   /* byte offset inside a vertex */
	uint32_t stride;       /* bytes between vertices (0 = every vertex reads the same bytes) */
	uint32_t n_floats;     /* floats copied (already clamped to the variable's size) */
	uint32_t dst_word;     /* destination word in the VS variable file */
} swgldev_fetch;

/* One linked varying (reference: VertexFragInOut pairs, swgl.c:2983-3005, 3648-3666). */
typedef struct
{
	uint32_t vs_word;      /* VS out variable, word offset in the VS variable file */
	uint32_t fs_word;      /* FS in variable, word offset in the FS variable file */
	uint32_t n_floats;     /* 1..4 */
	uint32_t slot;         /* float offset inside the packed per-vertex varying record */
	/* SWVS_PASS / SWVS_MATRIX only: the attribute bytes the VS out is a copy of */
	uint32_t src_offset, src_stride, src_floats;
	uint32_t _pad;
} swgldev_varying;

typedef struct
{
	swgldev_ptr data;      /* texels: uint8 [h][w][fpp] or float [h][w][fpp]  (swgl.c:2107-2118) */
	int32_t  width, height;
	int32_t  fpp;          /* channels per texel, 1..4 (swgl.c:2099-2102) */
	int32_t  is_float;     /* 0: bytes, converted with /255.0f at sample time; 1: floats */
	int32_t  wrap_s_repeat, wrap_t_repeat;
	/* glGenerateMipmap chain (swgl.c:2122-2173), built by swgldev_build_mipmaps: a table of 32 entries x 4
	 * fields, field-major (offset in floats behind the table, width, height, floats per texel at build time),
	 * followed by the levels as floats.  The levels keep their own sizes because the reference's vector
	 * outlives the image it was built from.  0 = no chain. */
	swgldev_ptr mips;
	int32_t  n_mips;
	int32_t  _pad;
} swgldev_texture;

struct swgl_ir_code;       /* swgl_ir.h */

/* Everything a draw needs, by value: the reference reads the same facts from its globals
 * (ActiveProgram, ActiveVertexArray, Viewport*, TextureUnits) inside glDrawArrays. */
typedef struct
{
	/* viewport (swgl.c:3151-3154) */
	int32_t  vx, vy;
	uint32_t vw, vh;
	/* geometry */
	swgldev_ptr vbo;  uint64_t vbo_bytes;
	swgldev_ptr ibo;  uint64_t ibo_bytes;     /* ibo == 0: glDrawArrays */
	int32_t  first;                           /* glDrawArrays `first`, or first index (elements) */
	uint32_t count;                           /* vertices in the stream */
	uint32_t n_vertices;                      /* indexed: vertices [0, n_vertices) are shaded once */
	/* program */
	int32_t  vs_kind, fs_kind;                /* SWVS_*, SWFS_* */
	const struct swgl_ir_code* vs_code;       /* host pointers; uploaded and cached by the layer */
	const struct swgl_ir_code* fs_code;
	uint64_t vs_code_id, fs_code_id;          /* cache keys (unique per compiled shader) */
	const uint32_t* vs_image; uint32_t vs_words;  /* initial VS variable file (uniforms set) */
	const uint32_t* fs_image; uint32_t fs_words;
	uint32_t pos_word;                        /* gl_Position, word offset in the VS file */
	uint32_t out_word; uint32_t out_floats;   /* first FS `out` variable (swgl.c:3412-3426) */
	swgldev_fetch   fetch[16];   uint32_t n_fetch;
	swgldev_varying varying[8];  uint32_t n_varying;
	uint32_t varying_floats;                  /* packed varying record size, floats */
	/* SWVS_PASS / SWVS_MATRIX */
	uint32_t pos_src_offset, pos_src_stride, pos_src_floats;
	float    pos_matrix[16];                  /* effective row-major matrix, load quirk applied */
	/* SWFS_VARYING / SWFS_TEXTURE */
	uint32_t fs_slot;                         /* float offset of the varying the FS consumes */
	uint32_t fs_slot_floats;
	uint32_t fs_swz_u, fs_swz_v;              /* component picks for the texture coordinate */
	int32_t  fs_tex_unit;
	swgldev_texture tex[8];                   /* texture units (swgl.c:2035) */
} swgldev_draw;

typedef struct
{
	uint64_t draws;
	uint64_t triangles_in;     /* input triangles of the last draw */
	uint64_t prims_out;        /* after near clip */
	uint64_t tested;           /* fragments evaluated (Barycentric calls in the reference) */
	uint64_t shaded;           /* fragments that passed the depth test */
	uint64_t tile_pairs;       /* (tile, primitive) bin entries */
	uint64_t bands;            /* band-entry records */
} swgldev_stats;

#ifndef __CUDACC_RTC__   /* host entry points: not part of a run-time compiled kernel's translation unit */
/* context ------------------------------------------------------------------------------ */
/* glInit (swgl.c:3713-3736): device = CUDA ordinal, or -1 for the current / LOCAL_RANK one. */
swgldev_ctx* swgldev_create(int device, uint32_t width, uint32_t height);
/* glInit after swglSetDeviceCount(n): a group of n device contexts (ordinals device .. device+n-1) behind one
 * handle, driven by the calling thread.  Every function below fans out to the members: allocations are replicated,
 * per-frame uploads are split over the members' PCIe links and completed over NVLink, draws are sharded
 * sort-first by tile-row bands, and the frame is assembled in one pinned host mirror all members write into. */
swgldev_ctx* swgldev_create_group(int device, int count, uint32_t width, uint32_t height);
void         swgldev_destroy(swgldev_ctx* c);
const char*  swgldev_last_error(swgldev_ctx* c);      /* "" when none; sticky until read */
void*        swgldev_stream(swgldev_ctx* c);          /* cudaStream_t all work is queued on */
int          swgldev_sync(swgldev_ctx* c);

/* memory: glBufferData / glTexImage2D copies (swgl.c:3142-3144, 2107-2118) ---------------- */
swgldev_ptr  swgldev_alloc(swgldev_ctx* c, uint64_t bytes);
int          swgldev_upload(swgldev_ctx* c, swgldev_ptr dst, const void* src, uint64_t bytes);
/* same contract, but the copy waits only for the draws that read `dst` (an allocation base): the
 * next frame's buffers cross PCIe while the current frame is still being rasterised */
int          swgldev_upload_overlapped(swgldev_ctx* c, swgldev_ptr dst, const void* src, uint64_t bytes);
void         swgldev_free(swgldev_ctx* c, swgldev_ptr p);
/* largest u32 in an uploaded element buffer (device reduction; the extension glDrawElements
 * shades vertices [0, max] once each) */
uint32_t     swgldev_max_index(swgldev_ctx* c, swgldev_ptr indices, uint64_t bytes);
/* swgldev_upload_overlapped() into part of an allocation (`base` + offset); and the largest index of element
 * data that work on the library's stream has written (multi-GPU: each rank uploads a slice, an all-gather over
 * NVLink on the library's stream replicates it) */
int          swgldev_upload_range(swgldev_ctx* c, swgldev_ptr base, uint64_t offset, const void* src, uint64_t bytes);
uint32_t     swgldev_max_index_after_stream(swgldev_ctx* c, swgldev_ptr indices, uint64_t bytes);
/* swgldev_upload_overlapped() of element data and swgldev_max_index() of it, with a single wait
 * (swglBufferRespecify(GL_ELEMENT_ARRAY_BUFFER), the per-frame upload of the end-to-end step). */
int          swgldev_upload_indices(swgldev_ctx* c, swgldev_ptr dst, const void* src, uint64_t bytes, uint32_t* max_index);

/* glGenerateMipmap (swgl.c:2122-2173): the 2x2 box chain of `base` as one allocation (layout: see
 * swgldev_texture.mips).  base->mips / n_mips = the chain the texture already has: its levels come first in the
 * new allocation, unchanged, and the levels of the current image behind them, as in the reference, whose
 * vector is only ever appended to (the caller frees the old allocation).  *n_levels = levels in the new
 * chain; returns 0 when nothing was added (image too small, 32 levels reached) or on error.
 * The chain is only sampled when the "mip_lod" option selects the defined LOD (swgl_b200.h). */
swgldev_ptr  swgldev_build_mipmaps(swgldev_ctx* c, const swgldev_texture* base, int32_t* n_levels);

/* glClear (swgl.c:3183-3214): rectangle is viewport ∩ framebuffer, already resolved. */
int swgldev_clear(swgldev_ctx* c, uint32_t flags, uint32_t color_word,
                  int32_t x0, int32_t y0, int32_t x1, int32_t y1);

/* glDrawArrays(GL_TRIANGLES) / glDrawElements (swgl.c:3609-3710 + DrawTriangle 3314-3473). */
int swgldev_draw_triangles(swgldev_ctx* c, const swgldev_draw* d);

/* Shaders outside the built-in shapes are compiled to kernels at run time (NVRTC, swgl_jit.cpp) on the
 * first draw that uses them; this does it ahead of that draw.  `c` may be NULL (no device): the code is
 * generated and compiled but not loaded.  0 = nothing to compile or compiled; -1 = failed, reason in msg. */
int swgldev_precompile(swgldev_ctx* c, const swgldev_draw* d, char* msg, size_t msg_len);

/* glDrawArrays(GL_POINTS) (swgl.c:3496-3608): one pixel per vertex, x scaled by VH/2 (sic), no
 * near clip, no depth test, no blend, no Y flip; the last point submitted to a pixel wins. */
int swgldev_draw_points(swgldev_ctx* c, const swgldev_draw* d);

/* glGetFramePtr (swgl.c:3738): refreshes and returns the pinned host mirror. */
uint32_t* swgldev_map_color(swgldev_ctx* c);
float*    swgldev_map_depth(swgldev_ctx* c);
/* the step after the path: frame pipelining over two pinned mirrors.  submit stamps the
 * frame rendered so far (ticket > 0) and returns at once; wait blocks until that frame is in its mirror
 * and returns it (valid until the next-but-one submit). */
uint64_t        swgldev_frame_submit(swgldev_ctx* c);
const uint32_t* swgldev_frame_wait(swgldev_ctx* c, uint64_t ticket);
/* colour attachment as bytes R, G, B, A (swizzled on the device) into `dst` (W*H*4 bytes of host memory) */
int       swgldev_read_rgba8(swgldev_ctx* c, void* dst);
/* page-locked host memory for upload sources (cudaHostAlloc; write-combined on request); no context needed */
void*     swgldev_host_alloc(uint64_t bytes, int write_combined);
void      swgldev_host_free(void* p);
swgldev_ptr swgldev_color_devptr(swgldev_ctx* c);
swgldev_ptr swgldev_depth_devptr(swgldev_ctx* c);
void      swgldev_fill(swgldev_ctx* c, uint32_t color_word, float depth); /* whole framebuffer */

void swgldev_get_stats(swgldev_ctx* c, swgldev_stats* out);

/* sort-first stripes: this context rasterises only tile rows r with (r / band_rows) % n == rank */
void swgldev_set_stripe(swgldev_ctx* c, uint32_t rank, uint32_t n_ranks, uint32_t band_tile_rows);
/* peer colour target: finished tiles are also stored to this (peer-mapped) colour buffer */
void swgldev_set_peer_color(swgldev_ctx* c, swgldev_ptr peer_color);
/* shared frame mirror (W*H*4 bytes of host memory every rank has mapped, e.g. POSIX shared memory): each
 * rank registers it and its raster kernels store finished tiles there over the rank's own PCIe link;
 * swgldev_sync copies the rank's bands when a frame could not be written through; NULL disables */
int         swgldev_set_shared_mirror(swgldev_ctx* c, void* host_ptr, uint64_t bytes);
/* CUDA IPC for the peer colour target: export this context's colour attachment (64-byte
 * cudaIpcMemHandle_t), open / close a handle exported by another process */
int         swgldev_ipc_export_color(swgldev_ctx* c, void* handle64);
swgldev_ptr swgldev_ipc_open(swgldev_ctx* c, const void* handle64);
void        swgldev_ipc_close(swgldev_ctx* c, swgldev_ptr p);
/* tuning / test hooks */
void swgldev_set_option(swgldev_ctx* c, const char* name, int64_t value);
int64_t swgldev_get_option(swgldev_ctx* c, const char* name);
#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif

#endif /* SWGL_DEV_H */
