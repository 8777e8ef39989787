/*
 * swgl_dev_types.cuh -- device-side data layout shared by the kernels of swgl_dev.cu.
 *
 * HBM layout of one draw (all scratch is grow-only and reused across draws):
 *
 *   clip[V]            float4   per shaded vertex: ((float)X, (float)Y, z_clip, w_clip) -- divide + viewport snap
 *                               done once per vertex instead of once per triangle corner
 *   clip_xy[V]         float2   clip-space x, y (only read when a triangle crosses the near plane)
 *   vary[V + 2T][NVF]  float    packed varyings; the last 2T records are the vertices the near
 *                               clipper creates (at most two per input triangle)
 *   prims[T] prims2[T] 64 B     screen-space primitive: 3 x (X, Y, z_clip, w_clip) + varying
 *                               record ids + first band entry; primitive id 2t+k keeps submission order
 *                               (k = 1, the second triangle of a near-clipped input, lives in prims2)
 *   bands[...]         16 B     TALL primitives only (more than two tile heights), per tile row: span-walk
 *                               state (x0, x1) on entering the row band + tile columns touched there
 *   tile_count[Nt]     u32      per-tile list length (atomic cursor; the raster CTA re-zeroes it)
 *   pairs[Nt][K]       u32      per-tile primitive lists, fixed capacity K (unordered; sorted by
 *                               primitive id inside the raster CTA)
 *   color/depth[H][W]  u32/f32  the framebuffer, row 0 = top (swgl.c:3156-3166)
 */
#ifndef SWGL_DEV_TYPES_CUH
#define SWGL_DEV_TYPES_CUH

#ifdef __CUDACC_RTC__
/* NVRTC translation unit (swgl_jit.cpp): no host headers */
typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;
typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
#else
#include <stdint.h>
#include <cuda_runtime.h>
#endif

#include "swgl_dev.h"
#include "swgl_ir.h"

#define SWGL_TILE        32
#define SWGL_TILE_SHIFT  5
#define SWGL_RASTER_THREADS 256
#define SWGL_BATCH       256      /* primitives staged per raster batch */
#define SWGL_SORT_CAP    2048     /* tile lists up to this length are sorted in shared memory */
#define SWGL_CTR_SLOTS   64

struct Prim
{
	float4   v[3];      /* (float)X, (float)Y, clip z, clip w -- submission order (swgl.c:3688-3691) */
	uint32_t vid[3];    /* varying record of each vertex */
	uint32_t band;      /* tall primitives: first band entry; 0xffffffff: short, no entries */
};

struct BandEntry                /* tall primitives only, one per tile row they cross */
{
	float    x0, x1;    /* walk state before the first row of the band (swgl.c:3351-3356) */
	uint32_t prim;
	uint32_t cols;      /* first tile column (11 bits) | last tile column << 11 | tile row << 22;
	                       0xffffffff = touches nothing (or a tile row of another rank) */
};

struct Counters
{
	uint32_t band_cursor;       /* band entries requested */
	uint32_t max_list;          /* longest tile list requested (only maintained when a list overflows) */
	uint32_t overflow;          /* bit0: a tile list exceeded K, bit1: band scratch too small; the draw
	                               is dropped and re-issued by the host */
	uint32_t prims_out;
	uint32_t ov_cursor;         /* entries requested in the overflow pool (lists longer than K) */
	uint32_t hot_tiles;         /* tiles whose list is longer than K */
	uint32_t hot_cursor;        /* words handed out of hot_store to the raster warps of those tiles */
	uint32_t _pad;
	unsigned long long pair_total;   /* bin entries consumed by the raster CTAs */
	unsigned long long tested[SWGL_CTR_SLOTS];
	unsigned long long shaded[SWGL_CTR_SLOTS];
};

struct DevTex
{
	const void* data;
	int32_t w, h, fpp, is_float, rep_s, rep_t;
	const uint32_t* mips;       /* glGenerateMipmap chain: SWGL_MIP_HEADER_WORDS of level table, then float levels (swgldev_texture.mips) */
	int32_t n_mips, _pad;
};

struct ClearParams
{
	uint32_t flags;             /* bit0 colour, bit1 depth; 0 = nothing pending */
	uint32_t word;
	int32_t  x0, y0, x1, y1;    /* framebuffer rectangle, rows NOT flipped (swgl.c:3193-3209) */
};

struct DrawParams
{
	/* framebuffer */
	uint32_t* color; float* depth; uint32_t* peer_color;
	uint32_t W, H;
	/* viewport and derived constants */
	int32_t vx, vy; uint32_t vw, vh;
	float hw, hh;               /* (float)(VW/2u), (float)(VH/2u)            swgl.c:3685-3686 */
	float fvx, fvy;             /* (float)VX, (float)VY */
	float xlimit, ylimit;       /* (float)(uint32)(VX+VW), (float)(uint32)(VY+VH) */
	int32_t ytop;               /* VH-1+2*VY: storage row = ytop - y           swgl.c:3386 */
	uint32_t tiles_x, tiles_y;  /* tiles are 32 wide and (1 << th_shift) tall: 32 (CTA kernels) or 8 (warp kernel) */
	uint32_t th_shift;
	uint32_t rank, n_ranks, band_rows;   /* sort-first bands, in units of 32 framebuffer rows */
	uint32_t owned_tile_rows;   /* warp rasteriser: tile rows of this rank (= tiles_y on one GPU) */
	/* geometry */
	const uint8_t* vbo; unsigned long long vbo_bytes;
	const uint32_t* ibo; unsigned long long ibo_count;
	int32_t first; uint32_t count, ntri, n_shade;
	/* scratch */
	float4* clip; float2* clip_xy; float* vary; uint32_t nvf; uint32_t clip_vid_base;
	Prim* prims; Prim* prims2;  /* primitive 2t at prims[t], 2t+1 (second triangle of a near-clipped input) at prims2[t] */
	BandEntry* bands; uint32_t cap_bands;
	uint32_t* tile_count; uint32_t* pairs; uint32_t bin_cap;   /* K: list capacity per tile; entries are (id << 1) | has_record */
	/* A few tiles may be much deeper than the rest (localised overdraw): entries beyond K go to one pool as
	 * (tile, entry) pairs, and the raster warp of such a tile assembles its whole list in hot_store (K inline
	 * entries + its pool entries) before sorting it.  ov_cap = 0: no pool (the CTA cross-check kernels). */
	uint2* ov_pool; uint32_t ov_cap; uint32_t max_hot;
	uint32_t* hot_store; uint32_t hot_cap;
	uint32_t lean_prims;        /* short unclipped primitives have no record (warp rasteriser draws) */
	uint32_t inline_tall;       /* tall primitives are inserted by the set-up kernel itself (no k_bin_tall launch) */
	uint32_t setup_big;         /* host side: the draw is set up by k_setup_big (a warp per triangle) */
	Counters* ctr;
	const float* lut255;        /* byte / 255.0f (swgl.c:2116, 3434-3437), computed once on the device */
	uint32_t* winner;           /* GL_POINTS: per-pixel index+1 of the last point submitted to it (0 = none) */
	/* shaders */
	int32_t vs_kind, fs_kind;
	const swgl_ir_op* vs_ops; uint32_t vs_nops; uint32_t vs_words;
	const swgl_ir_op* fs_ops; uint32_t fs_nops; uint32_t fs_words;
	uint32_t pos_word, out_word, out_floats;
	swgldev_fetch fetch[SWGL_MAX_FETCH]; uint32_t n_fetch;
	swgldev_varying varying[8]; uint32_t n_varying;
	uint32_t pos_src_offset, pos_src_stride, pos_src_floats;
	float pos_matrix[16];
	uint32_t fs_slot, fs_slot_floats, fs_swz_u, fs_swz_v; int32_t fs_tex_unit;
	DevTex tex[SWGL_MAX_TEX_UNITS];
	float* last_level;          /* mip_lod: MipMapLevel as the last DrawTriangle call left it (a global in the reference, swgl.c:3314-3316) */
	/* fused clear */
	ClearParams clear;
	uint32_t count_fragments;
	uint32_t mip_lod;           /* 1: textures with a mip chain are sampled with the per-triangle LOD (defined rsqrt) */
	uint32_t diag;              /* development only: skip parts of kernels to attribute time */
	/* initial variable files (uniform values) for the generic evaluator */
	uint32_t vs_image[SWGL_MAX_VAR_WORDS];
	uint32_t fs_image[SWGL_MAX_VAR_WORDS];
	/* host side only: run-time compiled kernels of this draw (vs_kind SWVS_JIT / fs_kind SWFS_JIT) */
	void* jit_vertex; void* jit_raster;
};

#endif
