/*
 * swgl_dev.cu -- CUDA device layer of libswgl_b200.so (sm_100a), behind the C ABI of
 * include/swgl_dev.h.  It replaces the reference's per-draw work:
 *
 *   k_clear          glClear                       swgl.c:3183-3214
 *   k_vertex         attribute fetch + VS + varying capture, one thread per vertex
 *                                                  swgl.c:3618-3666
 *   k_setup_bin      near clip, divide + viewport snap, triangle set-up, span walk, band
 *                    entries and tile counts, one thread per input triangle
 *                                                  swgl.c:499-697, 3683-3692, 3316-3361, 3466-3471
 *   k_scan_tiles     exclusive scan of the tile counts
 *   k_fill_bins      per-tile primitive lists
 *   k_raster         per-tile: sort the list by primitive id, then per pixel Barycentric,
 *                    perspective correction, depth test, varying interpolation, fragment
 *                    shader, blend, pack; 128-bit write-back
 *                                                  swgl.c:3358-3462
 *
 * Order dependence: the reference's depth test (LEQUAL with 0.0f = empty) and its
 * unconditional blend make the result depend on submission order, so every pixel sees its
 * fragments in ascending primitive id -- lists are sorted, and one thread owns a pixel.
 *
 * No tensor cores: the path is scan/scatter shaped, not a contraction.  Compiled with
 * -fmad=false (see swgl_dev_math.cuh).  There is no CPU fallback anywhere in this file.
 */
#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <vector>

#include "swgl_dev_math.cuh"

/* ========================================================================================
 * context
 * ====================================================================================== */
struct swgldev_ctx
{
	int device;
	cudaStream_t stream;
	uint32_t W, H, tiles_x, tiles_y;
	uint32_t* color; float* depth;
	uint32_t* h_color; float* h_depth;   /* pinned host mirrors (glGetFramePtr) */
	uint32_t* peer_color;
	uint32_t rank, n_ranks, band_rows;

	/* grow-only scratch */
	float4* clip; size_t cap_clip;
	float* vary; size_t cap_vary;        /* floats */
	Prim* prims; size_t cap_prims;
	uint2* prim_band; size_t cap_prim_band;
	BandEntry* bands; size_t cap_bands;
	uint32_t* pairs; size_t cap_pairs;
	uint32_t* tile_count; uint32_t* tile_off;
	Counters* ctr; Counters* h_ctr;      /* device counters, pinned snapshot */
	cudaEvent_t ctr_event;
	int ctr_pending;                     /* a snapshot copy is in flight for last_draw */

	std::map<uint64_t, swgl_ir_op*> code_cache;
	std::vector<void*> allocations;

	ClearParams pending_clear;
	DrawParams last_draw;                /* for re-issue after a scratch overflow */
	int last_draw_valid;
	int last_raster_path;

	/* options */
	int opt_fuse_clear, opt_count_fragments, opt_raster_path, opt_stage_timing;
	uint64_t n_launches;                 /* kernels launched since creation */
	cudaEvent_t stage_ev[8];
	double stage_us[8];                  /* accumulated per-stage device time (stage timing mode) */
	uint64_t stage_draws;

	swgldev_stats stats;
	uint64_t n_draws;
	char error[512];
};

static void set_err(swgldev_ctx* c, const char* what, cudaError_t e)
{
	if (c->error[0]) return; /* keep the first */
	if (e != cudaSuccess) snprintf(c->error, sizeof(c->error), "%s: %s", what, cudaGetErrorString(e));
	else snprintf(c->error, sizeof(c->error), "%s", what);
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(c, #call, e_); return -1; } } while (0)

template <typename T>
static int grow(swgldev_ctx* c, T** p, size_t* cap, size_t need)
{
	if (need <= *cap) return 0;
	size_t ncap = need + need / 4 + 1024;
	T* np = nullptr;
	/* all earlier work may still read the old block: stream-ordered free */
	CK(cudaMallocAsync((void**)&np, ncap * sizeof(T), c->stream));
	if (*p) CK(cudaFreeAsync(*p, c->stream));
	*p = np; *cap = ncap;
	return 0;
}

/* ========================================================================================
 * kernels
 * ====================================================================================== */

/* ---- glClear (swgl.c:3183-3214) ---- */
__global__ void k_clear(uint32_t* __restrict__ color, float* __restrict__ depth, uint32_t W,
                        ClearParams cp)
{
	int y = cp.y0 + (int)blockIdx.y;
	int x = cp.x0 + (int)(blockIdx.x * blockDim.x + threadIdx.x) * 4;
	if (y >= cp.y1) return;
	size_t base = (size_t)y * W;
	if (x + 3 < cp.x1 && (((base + (size_t)x) & 3u) == 0))
	{
		if (cp.flags & 1u) *(uint4*)(color + base + x) = make_uint4(cp.word, cp.word, cp.word, cp.word);
		if (cp.flags & 2u) *(float4*)(depth + base + x) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		return;
	}
	for (int k = 0; k < 4 && x + k < cp.x1; k++)
	{
		if (cp.flags & 1u) color[base + x + k] = cp.word;
		if (cp.flags & 2u) depth[base + x + k] = 0.0f;
	}
}

__global__ void k_fill_fb(uint32_t* __restrict__ color, float* __restrict__ depth, size_t n, uint32_t word, float d)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	size_t stride = (size_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride) { color[i] = word; depth[i] = d; }
}

/* ---- vertex stage (swgl.c:3618-3666) ---- */
template <int VS>
__global__ void __launch_bounds__(256) k_vertex(const __grid_constant__ DrawParams P)
{
	uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v == 0 && blockIdx.x == 0)
	{
		/* per-draw counters: this kernel is the first of the draw */
		P.ctr->band_cursor = 0; P.ctr->pair_total = 0; P.ctr->overflow = 0; P.ctr->prims_out = 0;
	}
	if (v < SWGL_CTR_SLOTS && blockIdx.x == 0) { P.ctr->tested[v] = 0ull; P.ctr->shaded[v] = 0ull; }
	if (v >= P.n_shade) return;
	/* glDrawArrays: stream vertex first + v;  glDrawElements: unique vertex v */
	long long vid = P.ibo ? (long long)v : (long long)P.first + (long long)v;
	float4 pos = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	float* vout = P.vary + (size_t)v * P.nvf;

	if (VS == SWVS_GENERIC)
	{
		uint32_t V[SWGL_MAX_VAR_WORDS];
		for (uint32_t k = 0; k < P.vs_words; k++) V[k] = P.vs_image[k];
		if (vid >= 0)
			for (uint32_t f = 0; f < P.n_fetch; f++)
			{
				float tmp[16];
				uint32_t n = P.fetch[f].n_floats;
				fetch_floats(P, (unsigned long long)vid, P.fetch[f].src_offset, P.fetch[f].stride, n, tmp);
				for (uint32_t k = 0; k < n; k++) V[P.fetch[f].dst_word + k] = __float_as_uint(tmp[k]);
			}
		ir_execute(P.vs_ops, P.vs_nops, V, P);
		pos = make_float4(__uint_as_float(V[P.pos_word]), __uint_as_float(V[P.pos_word + 1]),
		                  __uint_as_float(V[P.pos_word + 2]), __uint_as_float(V[P.pos_word + 3]));
		for (uint32_t k = 0; k < P.n_varying; k++)
			for (uint32_t j = 0; j < P.varying[k].n_floats; j++)
				vout[P.varying[k].slot + j] = __uint_as_float(V[P.varying[k].vs_word + j]);
	}
	else
	{
		float a[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
		if (vid >= 0) fetch_floats(P, (unsigned long long)vid, P.pos_src_offset, P.pos_src_stride, P.pos_src_floats, a);
		if (VS == SWVS_PASS) pos = make_float4(a[0], a[1], a[2], a[3]);
		else
		{   /* MatMulMat4Vec (swgl.c:758-768), left to right, no FMA */
			const float* m = P.pos_matrix;
			pos.x = m[0] * a[0] + m[1] * a[1] + m[2] * a[2] + m[3] * a[3];
			pos.y = m[4] * a[0] + m[5] * a[1] + m[6] * a[2] + m[7] * a[3];
			pos.z = m[8] * a[0] + m[9] * a[1] + m[10] * a[2] + m[11] * a[3];
			pos.w = m[12] * a[0] + m[13] * a[1] + m[14] * a[2] + m[15] * a[3];
		}
		for (uint32_t k = 0; k < P.n_varying; k++)
		{
			float t[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
			if (vid >= 0) fetch_floats(P, (unsigned long long)vid, P.varying[k].src_offset, P.varying[k].src_stride, P.varying[k].src_floats, t);
			for (uint32_t j = 0; j < P.varying[k].n_floats; j++) vout[P.varying[k].slot + j] = t[j];
		}
	}
	P.clip[v] = pos;
}

/* ---- near clip + snap + set-up + span walk + binning counts, one thread per triangle ---- */
__device__ __forceinline__ bool owns_tile_row(const DrawParams& P, uint32_t tr)
{
	return P.n_ranks <= 1 || ((tr / P.band_rows) % P.n_ranks) == P.rank;
}

__device__ __forceinline__ float4 near_intersect(const float4& a, const float4& b, float& t)
{
	/* IntersectNearPlane (swgl.c:455-466) */
	t = (a.z + a.w) / (a.w - b.w + a.z - b.z);
	float4 r;
	r.x = a.x + t * (b.x - a.x);
	r.y = a.y + t * (b.y - a.y);
	r.z = a.z + t * (b.z - a.z);
	r.w = a.w + t * (b.w - a.w);
	return r;
}

__device__ __forceinline__ void lerp_vary(const DrawParams& P, uint32_t dst, uint32_t a, uint32_t b, float t)
{
	/* InterpolateExValue (swgl.c:468-497) on the packed varying record */
	const float* pa = P.vary + (size_t)a * P.nvf;
	const float* pb = P.vary + (size_t)b * P.nvf;
	float* pd = P.vary + (size_t)dst * P.nvf;
	for (uint32_t k = 0; k < P.nvf; k++) { float x = pa[k], y = pb[k]; pd[k] = x + t * (y - x); }
}

__device__ __forceinline__ float4 to_screen(const float4& p, const DrawParams& P)
{
	/* swgl.c:3685-3691: int <- x / w * (VW/2) + (VW/2) + VX, stored back as float */
	int X = cvt_x86((fdiv(p.x, p.w) * P.hw + P.hw) + P.fvx);
	int Y = cvt_x86((fdiv(p.y, p.w) * P.hh + P.hh) + P.fvy);
	return make_float4((float)X, (float)Y, p.z, p.w);
}

__global__ void __launch_bounds__(128) k_setup_bin(const __grid_constant__ DrawParams P)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const bool active = t < P.ntri;      /* no early return: the warp allocates band entries together */

	/* stream positions 3t, 3t+1, 3t+2 (a trailing partial triangle is still drawn, swgl.c:3611) */
	uint32_t sid[3] = { 0u, 0u, 0u };
	float4 p[3];
	for (int j = 0; j < 3; j++)
	{
		uint32_t s = 3u * t + (uint32_t)j;
		p[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		if (!active) continue;
		if (P.ibo)
		{
			unsigned long long at = (unsigned long long)(long long)P.first + s;
			uint32_t idx = (at < P.ibo_count) ? __ldg(P.ibo + at) : 0xffffffffu;
			sid[j] = idx;
			if (idx < P.n_shade) p[j] = P.clip[idx];
		}
		else
		{
			sid[j] = s;
			if (s < P.n_shade) p[j] = P.clip[s];
		}
	}

	/* ClipTriangleAgainstNearPlane (swgl.c:532-696): inside iff z >= -w */
	int in_idx[3], out_idx[3], n_in = 0, n_out = 0;
	for (int j = 0; j < 3; j++)
	{
		if (p[j].z >= -p[j].w) in_idx[n_in++] = j; else out_idx[n_out++] = j;
	}
	if (!active) n_in = 0;

	Prim pr[2];
	int n_prims = 0;
	const uint32_t new0 = P.clip_vid_base + 2u * t, new1 = new0 + 1u;
	if (n_in == 3)
	{
		for (int j = 0; j < 3; j++) { pr[0].v[j] = p[j]; pr[0].vid[j] = sid[j]; }
		n_prims = 1;
	}
	else if (n_in == 1)
	{
		float t0, t1;
		const int a = in_idx[0];
		pr[0].v[0] = p[a]; pr[0].vid[0] = sid[a];
		pr[0].v[1] = near_intersect(p[a], p[out_idx[0]], t0); pr[0].vid[1] = new0;
		pr[0].v[2] = near_intersect(p[a], p[out_idx[1]], t1); pr[0].vid[2] = new1;
		lerp_vary(P, new0, sid[a], sid[out_idx[0]], t0);
		lerp_vary(P, new1, sid[a], sid[out_idx[1]], t1);
		n_prims = 1;
	}
	else if (n_in == 2)
	{
		float t0, t1;
		const int a = in_idx[0], b = in_idx[1], o = out_idx[0];
		pr[0].v[0] = p[a]; pr[0].vid[0] = sid[a];
		pr[0].v[1] = p[b]; pr[0].vid[1] = sid[b];
		pr[0].v[2] = near_intersect(p[a], p[o], t0); pr[0].vid[2] = new0;
		pr[1].v[0] = p[b]; pr[1].vid[0] = sid[b];
		pr[1].v[1] = pr[0].v[2]; pr[1].vid[1] = new0;
		pr[1].v[2] = near_intersect(p[b], p[o], t1); pr[1].vid[2] = new1;
		lerp_vary(P, new0, sid[a], sid[o], t0);
		lerp_vary(P, new1, sid[b], sid[o], t1);
		n_prims = 2;
	}

	/* divide + viewport snap, triangle set-up, band counts for both primitives */
	TriWalk wk[2];
	uint32_t tr_hi[2] = { 0u, 0u }, nbv[2] = { 0u, 0u };
	for (int k = 0; k < 2; k++)
	{
		if (k >= n_prims) continue;
		Prim& q = pr[k];
		for (int j = 0; j < 3; j++) q.v[j] = to_screen(q.v[j], P);
		q.pad = 0;
		if (tri_setup(q.v[0], q.v[1], q.v[2], P, wk[k]))
		{
			/* rows [ys, ye) -> storage rows ytop-ys (bottom-most) .. ytop-ye+1: tile rows hi..lo */
			tr_hi[k] = (uint32_t)(P.ytop - wk[k].ys) >> SWGL_TILE_SHIFT;
			const uint32_t tr_lo = (uint32_t)(P.ytop - (wk[k].ye - 1)) >> SWGL_TILE_SHIFT;
			nbv[k] = tr_hi[k] - tr_lo + 1u;
			if (P.n_ranks > 1)
			{
				/* sort-first: a primitive none of whose tile rows belong to this rank is dropped here */
				bool mine = false;
				for (uint32_t tr = tr_lo; tr <= tr_hi[k] && !mine; tr++) mine = owns_tile_row(P, tr);
				if (!mine) nbv[k] = 0u;
			}
		}
	}

	/* one atomic per warp allocates the band entries of all its primitives */
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t need = nbv[0] + nbv[1];
	uint32_t incl = need;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += y; }
	const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
	uint32_t warp_base = 0;
	if (lane == 31 && warp_total)
	{
		warp_base = atomicAdd(&P.ctr->band_cursor, warp_total);
	}
	warp_base = __shfl_sync(0xffffffffu, warp_base, 31);
	const uint32_t n_live = __popc(__ballot_sync(0xffffffffu, nbv[0] != 0u)) + __popc(__ballot_sync(0xffffffffu, nbv[1] != 0u));
	if (lane == 0 && n_live) atomicAdd(&P.ctr->prims_out, n_live);
	if ((unsigned long long)warp_base + warp_total > (unsigned long long)P.cap_bands)
	{
		if (lane == 0 && warp_total) P.ctr->overflow = 1u;   /* the draw is dropped and re-issued */
		return;
	}
	uint32_t base = warp_base + incl - need;

	for (int k = 0; k < 2; k++)
	{
		if (nbv[k] == 0u) continue;              /* dead slots are never referenced: nothing to write */
		const uint32_t pid = 2u * t + (uint32_t)k;
		const TriWalk& w = wk[k];
		P.prim_band[pid] = make_uint2(base, tr_hi[k]);
		P.prims[pid] = pr[k];
		/* the walk (swgl.c:3356-3361, 3466-3471), recording the state at every band entry */
		float x0 = w.c0x, x1 = w.c0x, s1 = w.s1;
		bool switched = false;
		uint32_t tr = tr_hi[k];
		int band_last_y = P.ytop - (int)(tr << SWGL_TILE_SHIFT);   /* last raster row of this band */
		float ex0 = x0, ex1 = x1;
		int cmin = 0x7fffffff, cmax = -1;
		for (int y = w.ys; y < w.ye; y++)
		{
			int xa, xb;
			row_span(x0, x1, P, xa, xb);
			if (xa < xb) { cmin = min(cmin, xa); cmax = max(cmax, xb - 1); }
			if (!switched && (float)y + 1.0f >= w.c1y) { switched = true; s1 = w.s2; x1 = w.c1x; }
			x0 += w.s0; x1 += s1;
			if (y == band_last_y || y == w.ye - 1)
			{
				BandEntry e;
				e.x0 = ex0; e.x1 = ex1; e.prim = pid; e.cols = 0xffffffffu;
				if (cmax >= 0 && owns_tile_row(P, tr))
				{
					const uint32_t c0 = (uint32_t)cmin >> SWGL_TILE_SHIFT, c1 = (uint32_t)cmax >> SWGL_TILE_SHIFT;
					e.cols = c0 | (c1 << 11) | (tr << 22);
					for (uint32_t cx = c0; cx <= c1; cx++) atomicAdd(&P.tile_count[tr * P.tiles_x + cx], 1u);
				}
				P.bands[base + (tr_hi[k] - tr)] = e;
				tr--; band_last_y += SWGL_TILE;
				ex0 = x0; ex1 = x1; cmin = 0x7fffffff; cmax = -1;
			}
		}
		base += nbv[k];
	}
}

/* ---- exclusive scan of the tile counts (one CTA) ---- */
__global__ void __launch_bounds__(1024) k_scan_tiles(const __grid_constant__ DrawParams P)
{
	__shared__ uint32_t warp_sums[32];
	__shared__ uint32_t carry_s;
	const uint32_t n = P.tiles_x * P.tiles_y;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
	if (tid == 0) carry_s = 0;
	__syncthreads();
	const bool dropped = P.ctr->overflow != 0;
	for (uint32_t base = 0; base < n; base += 1024u)
	{
		uint32_t i = base + tid;
		uint32_t c = (i < n) ? P.tile_count[i] : 0u;
		if (dropped && i < n) P.tile_count[i] = 0u;
		uint32_t x = c;
		for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((int)lane >= o) x += y; }
		if (lane == 31) warp_sums[wid] = x;
		__syncthreads();
		if (wid == 0)
		{
			uint32_t s = warp_sums[lane];
			for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if ((int)lane >= o) s += y; }
			warp_sums[lane] = s;
		}
		__syncthreads();
		const uint32_t carry = carry_s;
		uint32_t excl = carry + (wid ? warp_sums[wid - 1] : 0u) + (x - c);
		if (i < n) P.tile_off[i] = excl;
		__syncthreads();
		if (tid == 1023) carry_s = carry + warp_sums[31];
		__syncthreads();
	}
	if (tid == 0)
	{
		const uint32_t total = carry_s;
		P.tile_off[n] = total;
		P.ctr->pair_total = total;
		if (total > P.cap_pairs) P.ctr->overflow = 1u;
	}
	/* a dropped draw must leave the counts zero for the next one */
	__syncthreads();
	if (!dropped && P.ctr->overflow)
		for (uint32_t i = tid; i < n; i += 1024u) P.tile_count[i] = 0u;
}

/* ---- per-tile primitive lists: one thread per band entry ---- */
__global__ void __launch_bounds__(256) k_fill_bins(const __grid_constant__ DrawParams P)
{
	if (P.ctr->overflow) return;
	const uint32_t total = P.ctr->band_cursor;
	for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x)
	{
		const BandEntry be = P.bands[e];
		if (be.cols == 0xffffffffu) continue;
		const uint32_t c0 = be.cols & 0x7ffu, c1 = (be.cols >> 11) & 0x7ffu, tr = be.cols >> 22;
		for (uint32_t cx = c0; cx <= c1; cx++)
		{
			const uint32_t tile = tr * P.tiles_x + cx;
			/* counting back down to zero re-arms tile_count for the next draw */
			const uint32_t slot = atomicSub(&P.tile_count[tile], 1u) - 1u;
			P.pairs[P.tile_off[tile] + slot] = be.prim;
		}
	}
}

/* ---- fragment shading for the three shader shapes ---- */
struct FragIn
{
	float u, v, w;                /* perspective-corrected weights */
	uint32_t vid0, vid1, vid2;    /* varying records (generic shape) */
	/* fast shapes: the varying the shader consumes, per vertex; component k is at [k * stride]
	 * (stride 1 = straight from the packed records, SWGL_BATCH = staged in shared memory) */
	const float* a; const float* b; const float* c;
	uint32_t stride;
};

/* InterpolateLinearEx (swgl.c:3270-3297): a*u + b*v + c*w, left to right */
__device__ __forceinline__ float lerp3(const FragIn& f, uint32_t k)
{
	return f.a[k * f.stride] * f.u + f.b[k * f.stride] * f.v + f.c[k * f.stride] * f.w;
}

template <int FS>
__device__ __forceinline__ float4 run_fragment(const DrawParams& P, const FragIn& f)
{
	if (FS == SWFS_VARYING)
	{
		if (f.stride == 1 && (((uintptr_t)f.a | (uintptr_t)f.b | (uintptr_t)f.c) & 15u) == 0)
		{
			/* packed vec4 records: three 128-bit loads */
			const float4 a = __ldg((const float4*)f.a), b = __ldg((const float4*)f.b), c = __ldg((const float4*)f.c);
			return make_float4(a.x * f.u + b.x * f.v + c.x * f.w, a.y * f.u + b.y * f.v + c.y * f.w,
			                   a.z * f.u + b.z * f.v + c.z * f.w, a.w * f.u + b.w * f.v + c.w * f.w);
		}
		return make_float4(lerp3(f, 0), lerp3(f, 1), lerp3(f, 2), lerp3(f, 3));
	}
	if (FS == SWFS_TEXTURE)
	{
		const float tu = lerp3(f, P.fs_swz_u), tv = lerp3(f, P.fs_swz_v);
		return sample_nearest(P.tex[P.fs_tex_unit], tu, tv);
	}
	/* generic: interpolate every linked varying into the FS variable file, run the op list */
	uint32_t V[SWGL_MAX_VAR_WORDS];
	for (uint32_t k = 0; k < P.fs_words; k++) V[k] = P.fs_image[k];
	const float* va = P.vary + (size_t)f.vid0 * P.nvf;
	const float* vb = P.vary + (size_t)f.vid1 * P.nvf;
	const float* vc = P.vary + (size_t)f.vid2 * P.nvf;
	for (uint32_t k = 0; k < P.n_varying; k++)
		for (uint32_t j = 0; j < P.varying[k].n_floats; j++)
		{
			const uint32_t s = P.varying[k].slot + j;
			V[P.varying[k].fs_word + j] = __float_as_uint(va[s] * f.u + vb[s] * f.v + vc[s] * f.w);
		}
	ir_execute(P.fs_ops, P.fs_nops, V, P);
	float o[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
	for (uint32_t k = 0; k < P.out_floats; k++) o[k] = __uint_as_float(V[P.out_word + k]);
	return make_float4(o[0], o[1], o[2], o[3]);
}

/* ---- per-tile rasteriser, pixel-owner form ----
 * CTA = one 32x32 tile, 256 threads; thread t owns the 4x1 strip (row t/8, columns 4*(t%8)..+3)
 * and keeps its colour words and depths in registers for the whole list, so the tile is
 * read at most once and written once with 128-bit stores. */
struct RasterShared
{
	uint32_t ids[SWGL_SORT_CAP];
	uint16_t span[SWGL_TILE][SWGL_BATCH];   /* [row][prim]: xa | xb << 8, tile-local columns */
	uint32_t rowmask[SWGL_BATCH];
	float    bc[16][SWGL_BATCH];            /* BaryConst, one field per row */
	float    sv[12][SWGL_BATCH];            /* staged varyings: 3 vertices x 4 floats */
	uint32_t vid[3][SWGL_BATCH];
	unsigned long long red[2][SWGL_RASTER_THREADS / 32];
};

__device__ __forceinline__ void sort_ids_shared(uint32_t* ids, uint32_t n, uint32_t n_pow2)
{
	/* bitonic sort, ascending; padding entries are 0xffffffff */
	for (uint32_t i = n + threadIdx.x; i < n_pow2; i += blockDim.x) ids[i] = 0xffffffffu;
	__syncthreads();
	for (uint32_t k = 2; k <= n_pow2; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1)
		{
			for (uint32_t i = threadIdx.x; i < n_pow2; i += blockDim.x)
			{
				uint32_t ixj = i ^ j;
				if (ixj > i)
				{
					uint32_t a = ids[i], b = ids[ixj];
					bool up = (i & k) == 0;
					if ((a > b) == up) { ids[i] = b; ids[ixj] = a; }
				}
			}
			__syncthreads();
		}
}

__device__ __forceinline__ void sort_ids_global(uint32_t* ids, uint32_t n)
{
	/* fallback for very long lists: in-place bitonic network over global memory (L2-resident),
	 * virtual padding with 0xffffffff beyond n */
	uint32_t n_pow2 = 1; while (n_pow2 < n) n_pow2 <<= 1;
	for (uint32_t k = 2; k <= n_pow2; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1)
		{
			for (uint32_t i = threadIdx.x; i < n_pow2; i += blockDim.x)
			{
				uint32_t ixj = i ^ j;
				if (ixj > i && i < n)
				{
					uint32_t a = ids[i], b = ixj < n ? ids[ixj] : 0xffffffffu;
					bool up = (i & k) == 0;
					if ((a > b) == up && ixj < n) { ids[i] = b; ids[ixj] = a; }
				}
			}
			__syncthreads();
		}
}

template <int FS>
__global__ void __launch_bounds__(SWGL_RASTER_THREADS) k_raster(const __grid_constant__ DrawParams P)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RasterShared& S = *reinterpret_cast<RasterShared*>(smem_raw);

	if (P.ctr->overflow) return;
	const uint32_t tx = blockIdx.x, ty = blockIdx.y;
	if (!owns_tile_row(P, ty)) return;
	const uint32_t tile = ty * P.tiles_x + tx;
	const uint32_t list_off = P.tile_off[tile];
	const uint32_t n_list = P.tile_off[tile + 1] - list_off;

	const uint32_t tid = threadIdx.x;
	const uint32_t q = tid & 7u, r = tid >> 3;             /* strip column group, tile row */
	const int px0 = (int)(tx << SWGL_TILE_SHIFT) + (int)(q << 2);
	const int row = (int)(ty << SWGL_TILE_SHIFT) + (int)r; /* storage row */
	const bool row_ok = row < (int)P.H;
	const float fy = (float)(P.ytop - row);                /* raster y of this storage row */
	const size_t pix = (size_t)row * P.W + (size_t)px0;
	const bool vec_ok = row_ok && (px0 + 3 < (int)P.W) && ((pix & 3u) == 0);

	/* fused clear: pixels inside the pending clear rectangle start from the clear value */
	const ClearParams cp = P.clear;
	const bool in_clear_row = cp.flags && row >= cp.y0 && row < cp.y1;
	bool any_clear = false, all_clear_c = true, all_clear_d = true;
	for (int k = 0; k < 4; k++)
	{
		bool inside = in_clear_row && (px0 + k) >= cp.x0 && (px0 + k) < cp.x1;
		any_clear |= inside;
		all_clear_c &= inside && (cp.flags & 1u);
		all_clear_d &= inside && (cp.flags & 2u);
	}
	if (n_list == 0 && !cp.flags) return;

	uint32_t col[4]; float dep[4];
	if (row_ok)
	{
		if (vec_ok)
		{
			if (!all_clear_c) { uint4 c4 = *(const uint4*)(P.color + pix); col[0] = c4.x; col[1] = c4.y; col[2] = c4.z; col[3] = c4.w; }
			if (!all_clear_d) { float4 d4 = *(const float4*)(P.depth + pix); dep[0] = d4.x; dep[1] = d4.y; dep[2] = d4.z; dep[3] = d4.w; }
		}
		else
			for (int k = 0; k < 4; k++)
				if (px0 + k < (int)P.W) { col[k] = P.color[pix + k]; dep[k] = P.depth[pix + k]; }
		for (int k = 0; k < 4; k++)
		{
			bool inside = in_clear_row && (px0 + k) >= cp.x0 && (px0 + k) < cp.x1;
			if (inside && (cp.flags & 1u)) col[k] = cp.word;
			if (inside && (cp.flags & 2u)) dep[k] = 0.0f;
		}
	}
	bool dirty = any_clear;
	uint32_t n_tested = 0, n_shaded = 0;

	if (n_list > 0)
	{
		/* ascending primitive id = submission order */
		uint32_t* gl_ids = P.pairs + list_off;
		const bool in_shared = n_list <= SWGL_SORT_CAP;
		if (in_shared)
		{
			for (uint32_t i = tid; i < n_list; i += blockDim.x) S.ids[i] = gl_ids[i];
			uint32_t n_pow2 = 1; while (n_pow2 < n_list) n_pow2 <<= 1;
			sort_ids_shared(S.ids, n_list, n_pow2);
		}
		else sort_ids_global(gl_ids, n_list);
		const uint32_t* ids = in_shared ? S.ids : gl_ids;

		const int tile_x0 = (int)(tx << SWGL_TILE_SHIFT);
		const int band_last_y = P.ytop - (int)(ty << SWGL_TILE_SHIFT);  /* raster y of tile row 0 */
		const int band_first_y = band_last_y - (SWGL_TILE - 1);

		for (uint32_t base = 0; base < n_list; base += SWGL_BATCH)
		{
			const uint32_t nb = min((uint32_t)SWGL_BATCH, n_list - base);
			__syncthreads();
			/* phase A: one thread per primitive -- spans of the tile's rows + Barycentric constants */
			if (tid < nb)
			{
				const uint32_t pid = ids[base + tid];
				const Prim pr = P.prims[pid];
				const uint2 pb = P.prim_band[pid];
				const BandEntry be = P.bands[pb.x + (pb.y - ty)];
				TriWalk w;
				tri_setup(pr.v[0], pr.v[1], pr.v[2], P, w);
				const int y_in = max(w.ys, band_first_y);
				const int y_out = min(w.ye - 1, band_last_y);
				float x0 = be.x0, x1 = be.x1;
				bool switched = (y_in > w.ys) && ((float)y_in >= w.c1y);
				float s1 = switched ? w.s2 : w.s1;
				uint32_t mask = 0;
				for (int rr = 0; rr < SWGL_TILE; rr++) S.span[rr][tid] = 0;
				for (int y = y_in; y <= y_out; y++)
				{
					int xa, xb;
					row_span(x0, x1, P, xa, xb);
					xa = max(xa, tile_x0) - tile_x0;
					xb = min(xb, tile_x0 + SWGL_TILE) - tile_x0;
					if (xa < xb)
					{
						const int rr = band_last_y - y;
						S.span[rr][tid] = (uint16_t)(xa | (xb << 8));
						mask |= 1u << rr;
					}
					if (!switched && (float)y + 1.0f >= w.c1y) { switched = true; s1 = w.s2; x1 = w.c1x; }
					x0 += w.s0; x1 += s1;
				}
				S.rowmask[tid] = mask;
				if (mask)
				{
					BaryConst k;
					bary_setup(pr.v[0], pr.v[1], pr.v[2], k);
					S.bc[0][tid] = k.ax; S.bc[1][tid] = k.ay; S.bc[2][tid] = k.v0x; S.bc[3][tid] = k.v0y;
					S.bc[4][tid] = k.v1x; S.bc[5][tid] = k.v1y; S.bc[6][tid] = k.d00; S.bc[7][tid] = k.d01;
					S.bc[8][tid] = k.d11; S.bc[9][tid] = k.denom; S.bc[10][tid] = k.w0; S.bc[11][tid] = k.w1;
					S.bc[12][tid] = k.w2; S.bc[13][tid] = k.z0; S.bc[14][tid] = k.z1; S.bc[15][tid] = k.z2;
					S.vid[0][tid] = pr.vid[0]; S.vid[1][tid] = pr.vid[1]; S.vid[2][tid] = pr.vid[2];
					if (FS != SWFS_GENERIC)
						for (int j = 0; j < 3; j++)
						{
							const float* src = P.vary + (size_t)pr.vid[j] * P.nvf + P.fs_slot;
							for (uint32_t c = 0; c < 4; c++) S.sv[j * 4 + c][tid] = (c < P.fs_slot_floats) ? src[c] : 0.0f;
						}
				}
			}
			__syncthreads();
			/* phase B: every thread walks the batch in order for its own four pixels */
			if (row_ok)
				for (uint32_t j = 0; j < nb; j++)
				{
					if (!((S.rowmask[j] >> r) & 1u)) continue;
					const uint32_t sp = S.span[r][j];
					const int xa = (int)(sp & 0xffu), xb = (int)(sp >> 8);
					const int lx0 = (int)(q << 2);
					if (xb <= lx0 || xa >= lx0 + 4) continue;
					BaryConst k;
					k.ax = S.bc[0][j]; k.ay = S.bc[1][j]; k.v0x = S.bc[2][j]; k.v0y = S.bc[3][j];
					k.v1x = S.bc[4][j]; k.v1y = S.bc[5][j]; k.d00 = S.bc[6][j]; k.d01 = S.bc[7][j];
					k.d11 = S.bc[8][j]; k.denom = S.bc[9][j]; k.w0 = S.bc[10][j]; k.w1 = S.bc[11][j];
					k.w2 = S.bc[12][j]; k.z0 = S.bc[13][j]; k.z1 = S.bc[14][j]; k.z2 = S.bc[15][j];
#pragma unroll
					for (int kk = 0; kk < 4; kk++)
					{
						const int lx = lx0 + kk;
						if (lx < xa || lx >= xb) continue;
						n_tested++;
						FragIn f;
						float z;
						frag_weights(k, (float)(px0 + kk), fy, f.u, f.v, f.w, z);
						const float cur = dep[kk];
						if (cur == 0.0f || cur >= z)   /* swgl.c:3387 */
						{
							dep[kk] = z;
							n_shaded++;
							f.vid0 = S.vid[0][j]; f.vid1 = S.vid[1][j]; f.vid2 = S.vid[2][j];
							f.a = &S.sv[0][j]; f.b = &S.sv[4][j]; f.c = &S.sv[8][j]; f.stride = SWGL_BATCH;
							const float4 o = run_fragment<FS>(P, f);
							col[kk] = blend_pack(o.x, o.y, o.z, o.w, col[kk]);
							dirty = true;
						}
					}
				}
		}
	}

	/* write-back: 128-bit stores of the finished strip */
	if (row_ok && dirty)
	{
		for (int k = 0; k < 4; k++) dep[k] = canon_nan(dep[k]);
		if (vec_ok)
		{
			*(uint4*)(P.color + pix) = make_uint4(col[0], col[1], col[2], col[3]);
			*(float4*)(P.depth + pix) = make_float4(dep[0], dep[1], dep[2], dep[3]);
			if (P.peer_color) *(uint4*)(P.peer_color + pix) = make_uint4(col[0], col[1], col[2], col[3]);
		}
		else
			for (int k = 0; k < 4; k++)
				if (px0 + k < (int)P.W)
				{
					P.color[pix + k] = col[k]; P.depth[pix + k] = dep[k];
					if (P.peer_color) P.peer_color[pix + k] = col[k];
				}
	}

	if (P.count_fragments)
	{
		unsigned long long a = n_tested, b = n_shaded;
		for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
		__syncthreads();
		if ((tid & 31u) == 0) { S.red[0][tid >> 5] = a; S.red[1][tid >> 5] = b; }
		__syncthreads();
		if (tid == 0)
		{
			unsigned long long ta = 0, tb = 0;
			for (int i = 0; i < SWGL_RASTER_THREADS / 32; i++) { ta += S.red[0][i]; tb += S.red[1][i]; }
			if (ta) atomicAdd(&P.ctr->tested[tile % SWGL_CTR_SLOTS], ta);
			if (tb) atomicAdd(&P.ctr->shaded[tile % SWGL_CTR_SLOTS], tb);
		}
	}
}

/* ========================================================================================
 * host side of the C ABI
 * ====================================================================================== */
#include "swgl_raster_frag.cuh"

/* raster_path: 1 = pixel-owner (k_raster), 2 = fragment-parallel (k_raster_frag), 0 = default */
template <int FS>
static void launch_raster(swgldev_ctx* c, const DrawParams& P)
{
	dim3 grid(c->tiles_x, c->tiles_y);
	const int path = c->opt_raster_path == 1 ? 1 : 2;
	c->last_raster_path = path;
	if (path == 1) k_raster<FS><<<grid, SWGL_RASTER_THREADS, sizeof(RasterShared), c->stream>>>(P);
	else k_raster_frag<FS><<<grid, FRAG_THREADS, sizeof(FragShared), c->stream>>>(P);
}

extern "C" {

swgldev_ctx* swgldev_create(int device, uint32_t width, uint32_t height)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return nullptr; }
	if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
	if (device >= n) device = device % n;
	if (cudaSetDevice(device) != cudaSuccess) return nullptr;

	swgldev_ctx* c = new swgldev_ctx();
	c->device = device;
	c->W = width; c->H = height;
	c->tiles_x = (width + SWGL_TILE - 1) / SWGL_TILE;
	c->tiles_y = (height + SWGL_TILE - 1) / SWGL_TILE;
	c->stream = nullptr; c->color = nullptr; c->depth = nullptr; c->h_color = nullptr; c->h_depth = nullptr;
	c->peer_color = nullptr; c->rank = 0; c->n_ranks = 1; c->band_rows = 1;
	c->clip = nullptr; c->cap_clip = 0; c->vary = nullptr; c->cap_vary = 0;
	c->prims = nullptr; c->cap_prims = 0; c->prim_band = nullptr; c->cap_prim_band = 0;
	c->bands = nullptr; c->cap_bands = 0; c->pairs = nullptr; c->cap_pairs = 0;
	c->tile_count = nullptr; c->tile_off = nullptr; c->ctr = nullptr; c->h_ctr = nullptr;
	c->ctr_pending = 0; c->last_draw_valid = 0; c->last_raster_path = 0;
	c->opt_fuse_clear = 1; c->opt_count_fragments = 1; c->opt_raster_path = 0; c->opt_stage_timing = 0;
	c->n_launches = 0; c->stage_draws = 0;
	for (int i = 0; i < 8; i++) { c->stage_ev[i] = nullptr; c->stage_us[i] = 0.0; }
	memset(&c->pending_clear, 0, sizeof(c->pending_clear));
	memset(&c->stats, 0, sizeof(c->stats));
	c->n_draws = 0; c->error[0] = 0;

	const size_t npx = (size_t)width * height;
	const size_t ntiles = (size_t)c->tiles_x * c->tiles_y;
	bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess
	       && cudaMalloc((void**)&c->color, (npx ? npx : 1) * 4) == cudaSuccess
	       && cudaMalloc((void**)&c->depth, (npx ? npx : 1) * 4) == cudaSuccess
	       && cudaMallocHost((void**)&c->h_color, (npx ? npx : 1) * 4) == cudaSuccess
	       && cudaMallocHost((void**)&c->h_depth, (npx ? npx : 1) * 4) == cudaSuccess
	       && cudaMalloc((void**)&c->tile_count, (ntiles + 1) * 4) == cudaSuccess
	       && cudaMalloc((void**)&c->tile_off, (ntiles + 2) * 4) == cudaSuccess
	       && cudaMalloc((void**)&c->ctr, sizeof(Counters)) == cudaSuccess
	       && cudaMallocHost((void**)&c->h_ctr, sizeof(Counters)) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->ctr_event, cudaEventDisableTiming) == cudaSuccess
	       && cudaEventCreate(&c->stage_ev[0]) == cudaSuccess && cudaEventCreate(&c->stage_ev[1]) == cudaSuccess
	       && cudaEventCreate(&c->stage_ev[2]) == cudaSuccess && cudaEventCreate(&c->stage_ev[3]) == cudaSuccess
	       && cudaEventCreate(&c->stage_ev[4]) == cudaSuccess && cudaEventCreate(&c->stage_ev[5]) == cudaSuccess
	       && cudaMemset(c->tile_count, 0, (ntiles + 1) * 4) == cudaSuccess
	       && cudaMemset(c->ctr, 0, sizeof(Counters)) == cudaSuccess;
	if (!ok)
	{
		fprintf(stderr, "swgl_b200: device context creation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
		swgldev_destroy(c);
		return nullptr;
	}
	memset(c->h_ctr, 0, sizeof(Counters));

	/* the raster kernels need more than the default 48 KB of dynamic shared memory */
	const int smem = (int)sizeof(RasterShared);
	cudaFuncSetAttribute(k_raster<SWFS_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	cudaFuncSetAttribute(k_raster<SWFS_VARYING>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	cudaFuncSetAttribute(k_raster<SWFS_TEXTURE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	const int smem_f = (int)sizeof(FragShared);
	cudaFuncSetAttribute(k_raster_frag<SWFS_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f);
	cudaFuncSetAttribute(k_raster_frag<SWFS_VARYING>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f);
	cudaFuncSetAttribute(k_raster_frag<SWFS_TEXTURE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f);
	return c;
}

void swgldev_destroy(swgldev_ctx* c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	for (void* p : c->allocations) cudaFree(p);
	for (auto& kv : c->code_cache) cudaFree(kv.second);
	cudaFree(c->color); cudaFree(c->depth); cudaFreeHost(c->h_color); cudaFreeHost(c->h_depth);
	cudaFree(c->tile_count); cudaFree(c->tile_off); cudaFree(c->ctr); cudaFreeHost(c->h_ctr);
	if (c->clip) cudaFree(c->clip);
	if (c->vary) cudaFree(c->vary);
	if (c->prims) cudaFree(c->prims);
	if (c->prim_band) cudaFree(c->prim_band);
	if (c->bands) cudaFree(c->bands);
	if (c->pairs) cudaFree(c->pairs);
	if (c->ctr_event) cudaEventDestroy(c->ctr_event);
	for (int i = 0; i < 8; i++) if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
	if (c->stream) cudaStreamDestroy(c->stream);
	cudaGetLastError();
	delete c;
}

const char* swgldev_last_error(swgldev_ctx* c)
{
	static char out[512];
	snprintf(out, sizeof(out), "%s", c->error);
	c->error[0] = 0;
	return out;
}

void* swgldev_stream(swgldev_ctx* c) { return (void*)c->stream; }

swgldev_ptr swgldev_alloc(swgldev_ctx* c, uint64_t bytes)
{
	void* p = nullptr;
	cudaSetDevice(c->device);
	cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
	if (e != cudaSuccess) { set_err(c, "cudaMalloc (buffer/texture upload)", e); return 0; }
	c->allocations.push_back(p);
	return (swgldev_ptr)(uintptr_t)p;
}

void swgldev_free(swgldev_ctx* c, swgldev_ptr p)
{
	void* q = (void*)(uintptr_t)p;
	for (size_t i = 0; i < c->allocations.size(); i++)
		if (c->allocations[i] == q)
		{
			c->allocations[i] = c->allocations.back();
			c->allocations.pop_back();
			cudaStreamSynchronize(c->stream); /* queued draws may still read it */
			cudaFree(q);
			return;
		}
}

int swgldev_upload(swgldev_ctx* c, swgldev_ptr dst, const void* src, uint64_t bytes)
{
	/* the caller may free `src` on return (swgl.c:3142-3144), so the copy completes here;
	 * pinned sources go at full PCIe rate, pageable ones through the driver's staging */
	CK(cudaMemcpyAsync((void*)(uintptr_t)dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

/* ---- deferred overflow check: wait for the counter snapshot of the previous draw and, when
 * its scratch was too small, grow and re-issue it before anything later is queued ---- */
static int launch_draw(swgldev_ctx* c, DrawParams& P);

static int settle_last_draw(swgldev_ctx* c)
{
	int guard = 0;
	while (c->ctr_pending)
	{
		CK(cudaEventSynchronize(c->ctr_event));
		c->ctr_pending = 0;
		const Counters& h = *c->h_ctr;
		c->stats.prims_out = h.prims_out;
		c->stats.tile_pairs = h.pair_total;
		c->stats.bands = h.band_cursor;
		if (!h.overflow || !c->last_draw_valid) break;
		if (++guard > 6) { set_err(c, "draw dropped: binning scratch could not be sized", cudaSuccess); break; }
		/* grow to what the dropped draw asked for, then run it again */
		size_t need_bands = (size_t)h.band_cursor + 1024, need_pairs = (size_t)h.pair_total + 1024;
		if (h.band_cursor > c->cap_bands) need_pairs = need_pairs < c->cap_pairs * 2 ? c->cap_pairs * 2 : need_pairs;
		if (grow(c, &c->bands, &c->cap_bands, need_bands)) return -1;
		if (grow(c, &c->pairs, &c->cap_pairs, need_pairs)) return -1;
		DrawParams P = c->last_draw;
		if (launch_draw(c, P)) return -1;
	}
	return 0;
}

static int flush_clear(swgldev_ctx* c)
{
	ClearParams cp = c->pending_clear;
	if (!cp.flags) return 0;
	c->pending_clear.flags = 0;
	if (cp.x1 <= cp.x0 || cp.y1 <= cp.y0) return 0;
	dim3 block(128), grid(((uint32_t)(cp.x1 - cp.x0) + 511u) / 512u, (uint32_t)(cp.y1 - cp.y0));
	k_clear<<<grid, block, 0, c->stream>>>(c->color, c->depth, c->W, cp);
	c->n_launches++;
	CK(cudaGetLastError());
	return 0;
}

int swgldev_sync(swgldev_ctx* c)
{
	cudaSetDevice(c->device);
	if (settle_last_draw(c)) return -1;
	if (flush_clear(c)) return -1;
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

int swgldev_clear(swgldev_ctx* c, uint32_t flags, uint32_t color_word, int32_t x0, int32_t y0, int32_t x1, int32_t y1)
{
	cudaSetDevice(c->device);
	if (!flags) return 0;
	if (settle_last_draw(c)) return -1;
	ClearParams& pc = c->pending_clear;
	if (pc.flags && pc.x0 == x0 && pc.y0 == y0 && pc.x1 == x1 && pc.y1 == y1)
	{
		/* same rectangle: the later clear wins per attachment */
		if (flags & 1u) pc.word = color_word;
		pc.flags |= flags;
	}
	else
	{
		if (flush_clear(c)) return -1;
		pc.flags = flags; pc.word = color_word; pc.x0 = x0; pc.y0 = y0; pc.x1 = x1; pc.y1 = y1;
	}
	if (!c->opt_fuse_clear) return flush_clear(c);
	return 0;
}

void swgldev_fill(swgldev_ctx* c, uint32_t color_word, float depth)
{
	cudaSetDevice(c->device);
	settle_last_draw(c);
	c->pending_clear.flags = 0;
	size_t n = (size_t)c->W * c->H;
	k_fill_fb<<<1184, 256, 0, c->stream>>>(c->color, c->depth, n, color_word, depth);
}

static const swgl_ir_op* upload_code(swgldev_ctx* c, uint64_t id, const swgl_ir_code* code)
{
	auto it = c->code_cache.find(id);
	if (it != c->code_cache.end()) return it->second;
	swgl_ir_op* d = nullptr;
	size_t bytes = sizeof(swgl_ir_op) * (code->n_ops ? code->n_ops : 1);
	if (cudaMalloc((void**)&d, bytes) != cudaSuccess) return nullptr;
	cudaMemcpy(d, code->ops, sizeof(swgl_ir_op) * code->n_ops, cudaMemcpyHostToDevice);
	c->code_cache[id] = d;
	return d;
}

static int launch_draw(swgldev_ctx* c, DrawParams& P)
{
	P.cap_bands = (uint32_t)(c->cap_bands > 0xffffffffull ? 0xffffffffull : c->cap_bands);
	P.cap_pairs = (uint32_t)(c->cap_pairs > 0xffffffffull ? 0xffffffffull : c->cap_pairs);
	P.bands = c->bands; P.pairs = c->pairs;
	c->last_draw = P; c->last_draw_valid = 1;

	const bool timing = c->opt_stage_timing != 0;
#define STAGE(i) do { if (timing) cudaEventRecord(c->stage_ev[i], c->stream); } while (0)
	const uint32_t vb = (P.n_shade + 255u) / 256u;
	STAGE(0);
	if (P.vs_kind == SWVS_PASS) k_vertex<SWVS_PASS><<<vb ? vb : 1, 256, 0, c->stream>>>(P);
	else if (P.vs_kind == SWVS_MATRIX) k_vertex<SWVS_MATRIX><<<vb ? vb : 1, 256, 0, c->stream>>>(P);
	else k_vertex<SWVS_GENERIC><<<vb ? vb : 1, 256, 0, c->stream>>>(P);
	STAGE(1);
	k_setup_bin<<<(P.ntri + 127u) / 128u, 128, 0, c->stream>>>(P);
	STAGE(2);
	k_scan_tiles<<<1, 1024, 0, c->stream>>>(P);
	STAGE(3);
	CK(cudaMemcpyAsync(c->h_ctr, c->ctr, 16, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaEventRecord(c->ctr_event, c->stream));
	c->ctr_pending = 1;
	k_fill_bins<<<148 * 8, 256, 0, c->stream>>>(P);
	STAGE(4);
	if (P.fs_kind == SWFS_VARYING) launch_raster<SWFS_VARYING>(c, P);
	else if (P.fs_kind == SWFS_TEXTURE) launch_raster<SWFS_TEXTURE>(c, P);
	else launch_raster<SWFS_GENERIC>(c, P);
	STAGE(5);
#undef STAGE
	c->n_launches += 5;
	CK(cudaGetLastError());
	if (timing)
	{
		CK(cudaEventSynchronize(c->stage_ev[5]));
		for (int i = 0; i < 5; i++)
		{
			float ms = 0.0f;
			if (cudaEventElapsedTime(&ms, c->stage_ev[i], c->stage_ev[i + 1]) == cudaSuccess) c->stage_us[i] += 1000.0 * ms;
		}
		c->stage_draws++;
	}
	return 0;
}

int swgldev_draw_triangles(swgldev_ctx* c, const swgldev_draw* d)
{
	cudaSetDevice(c->device);
	if (settle_last_draw(c)) return -1;
	c->last_draw_valid = 0;

	/* The tile mapping needs storage row = VH-1+2*VY-y to be a bijection on the viewport rows,
	 * i.e. the viewport lies inside the framebuffer vertically (otherwise the reference clamps
	 * several raster rows onto row Height-1, swgl.c:3386). */
	if (d->vy < 0 || (uint64_t)d->vy + d->vh > c->H || d->vw > 0x7fffffffu || d->vh > 0x7fffffffu
	    || c->tiles_x > 2047u || c->tiles_y > 1023u)
	{
		set_err(c, "draw skipped: viewport must lie inside the framebuffer rows (0 <= y, y+height <= Height)", cudaSuccess);
		return flush_clear(c);
	}
	const uint32_t ntri = (d->count + 2u) / 3u;
	if (ntri == 0 || d->vh == 0) return flush_clear(c);
	if (d->varying_floats > SWGL_MAX_VARYING_FLOATS || d->vs_words > SWGL_MAX_VAR_WORDS || d->fs_words > SWGL_MAX_VAR_WORDS)
	{
		set_err(c, "draw skipped: shader interface too large", cudaSuccess);
		return flush_clear(c);
	}

	DrawParams P;
	memset(&P, 0, sizeof(P));
	P.color = c->color; P.depth = c->depth; P.peer_color = c->peer_color; P.W = c->W; P.H = c->H;
	P.vx = d->vx; P.vy = d->vy; P.vw = d->vw; P.vh = d->vh;
	P.hw = (float)(d->vw / 2u); P.hh = (float)(d->vh / 2u);
	P.fvx = (float)d->vx; P.fvy = (float)d->vy;
	P.xlimit = (float)(uint32_t)((uint32_t)d->vx + d->vw);
	P.ylimit = (float)(uint32_t)((uint32_t)d->vy + d->vh);
	P.ytop = (int32_t)d->vh - 1 + 2 * d->vy;
	P.tiles_x = c->tiles_x; P.tiles_y = c->tiles_y;
	P.rank = c->rank; P.n_ranks = c->n_ranks; P.band_rows = c->band_rows ? c->band_rows : 1;
	P.vbo = (const uint8_t*)(uintptr_t)d->vbo; P.vbo_bytes = d->vbo_bytes;
	P.ibo = (const uint32_t*)(uintptr_t)d->ibo; P.ibo_count = d->ibo_bytes / 4u;
	P.first = d->first; P.count = d->count; P.ntri = ntri;
	P.n_shade = d->ibo ? d->n_vertices : 3u * ntri;
	P.nvf = d->varying_floats;
	P.clip_vid_base = P.n_shade;
	P.vs_kind = d->vs_kind; P.fs_kind = d->fs_kind;
	P.vs_words = d->vs_words; P.fs_words = d->fs_words;
	P.pos_word = d->pos_word; P.out_word = d->out_word; P.out_floats = d->out_floats;
	memcpy(P.fetch, d->fetch, sizeof(P.fetch)); P.n_fetch = d->n_fetch;
	memcpy(P.varying, d->varying, sizeof(P.varying)); P.n_varying = d->n_varying;
	P.pos_src_offset = d->pos_src_offset; P.pos_src_stride = d->pos_src_stride; P.pos_src_floats = d->pos_src_floats;
	memcpy(P.pos_matrix, d->pos_matrix, sizeof(P.pos_matrix));
	P.fs_slot = d->fs_slot; P.fs_slot_floats = d->fs_slot_floats;
	P.fs_swz_u = d->fs_swz_u; P.fs_swz_v = d->fs_swz_v; P.fs_tex_unit = d->fs_tex_unit;
	for (int u = 0; u < SWGL_MAX_TEX_UNITS; u++)
	{
		P.tex[u].data = (const void*)(uintptr_t)d->tex[u].data;
		P.tex[u].w = d->tex[u].width; P.tex[u].h = d->tex[u].height; P.tex[u].fpp = d->tex[u].fpp;
		P.tex[u].is_float = d->tex[u].is_float; P.tex[u].rep_s = d->tex[u].wrap_s_repeat; P.tex[u].rep_t = d->tex[u].wrap_t_repeat;
	}
	if (d->vs_image) memcpy(P.vs_image, d->vs_image, 4u * d->vs_words);
	if (d->fs_image) memcpy(P.fs_image, d->fs_image, 4u * d->fs_words);
	P.count_fragments = (uint32_t)c->opt_count_fragments;

	if (P.vs_kind == SWVS_GENERIC)
	{
		P.vs_ops = upload_code(c, d->vs_code_id, d->vs_code); P.vs_nops = d->vs_code->n_ops;
		if (!P.vs_ops) { set_err(c, "vertex shader upload failed", cudaGetLastError()); return -1; }
	}
	if (P.fs_kind == SWFS_GENERIC)
	{
		P.fs_ops = upload_code(c, d->fs_code_id, d->fs_code); P.fs_nops = d->fs_code->n_ops;
		if (!P.fs_ops) { set_err(c, "fragment shader upload failed", cudaGetLastError()); return -1; }
	}

	/* scratch */
	const size_t n_prims = 2ull * ntri;
	if (grow(c, &c->clip, &c->cap_clip, (size_t)P.n_shade)) return -1;
	if (grow(c, &c->vary, &c->cap_vary, ((size_t)P.n_shade + n_prims) * (P.nvf ? P.nvf : 1) + 4)) return -1;
	if (grow(c, &c->prims, &c->cap_prims, n_prims)) return -1;
	if (grow(c, &c->prim_band, &c->cap_prim_band, n_prims)) return -1;
	if (grow(c, &c->bands, &c->cap_bands, n_prims * 2 + (1u << 20))) return -1;
	if (grow(c, &c->pairs, &c->cap_pairs, n_prims * 3 + (1u << 21))) return -1;
	P.clip = c->clip; P.vary = c->vary; P.prims = c->prims; P.prim_band = c->prim_band;
	P.tile_count = c->tile_count; P.tile_off = c->tile_off; P.ctr = c->ctr;

	/* fused clear: the raster kernel starts the covered pixels from the clear value */
	P.clear = c->pending_clear;
	c->pending_clear.flags = 0;

	c->n_draws++;
	c->stats.draws = c->n_draws;
	c->stats.triangles_in = ntri;
	return launch_draw(c, P);
}

uint32_t* swgldev_map_color(swgldev_ctx* c)
{
	if (swgldev_sync(c)) return c->h_color;
	cudaMemcpyAsync(c->h_color, c->color, (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->stream);
	cudaStreamSynchronize(c->stream);
	return c->h_color;
}

float* swgldev_map_depth(swgldev_ctx* c)
{
	if (swgldev_sync(c)) return c->h_depth;
	cudaMemcpyAsync(c->h_depth, c->depth, (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->stream);
	cudaStreamSynchronize(c->stream);
	return c->h_depth;
}

swgldev_ptr swgldev_color_devptr(swgldev_ctx* c) { return (swgldev_ptr)(uintptr_t)c->color; }
swgldev_ptr swgldev_depth_devptr(swgldev_ctx* c) { return (swgldev_ptr)(uintptr_t)c->depth; }

void swgldev_get_stats(swgldev_ctx* c, swgldev_stats* out)
{
	swgldev_sync(c);
	Counters h;
	if (cudaMemcpy(&h, c->ctr, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess)
	{
		unsigned long long a = 0, b = 0;
		for (int i = 0; i < SWGL_CTR_SLOTS; i++) { a += h.tested[i]; b += h.shaded[i]; }
		c->stats.tested = a; c->stats.shaded = b;
		c->stats.prims_out = h.prims_out; c->stats.tile_pairs = h.pair_total; c->stats.bands = h.band_cursor;
	}
	*out = c->stats;
}

void swgldev_set_stripe(swgldev_ctx* c, uint32_t rank, uint32_t n_ranks, uint32_t band_tile_rows)
{
	swgldev_sync(c);
	c->rank = rank; c->n_ranks = n_ranks ? n_ranks : 1; c->band_rows = band_tile_rows ? band_tile_rows : 1;
}

void swgldev_set_peer_color(swgldev_ctx* c, swgldev_ptr peer_color)
{
	swgldev_sync(c);
	c->peer_color = (uint32_t*)(uintptr_t)peer_color;
}

int swgldev_ipc_export_color(swgldev_ctx* c, void* handle64)
{
	cudaIpcMemHandle_t h;
	CK(cudaIpcGetMemHandle(&h, c->color));
	memcpy(handle64, &h, sizeof(h));
	return 0;
}

swgldev_ptr swgldev_ipc_open(swgldev_ctx* c, const void* handle64)
{
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof(h));
	void* p = nullptr;
	cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
	if (e != cudaSuccess) { set_err(c, "cudaIpcOpenMemHandle", e); return 0; }
	return (swgldev_ptr)(uintptr_t)p;
}

void swgldev_ipc_close(swgldev_ctx* c, swgldev_ptr p)
{
	swgldev_sync(c);
	if (c->peer_color == (uint32_t*)(uintptr_t)p) c->peer_color = nullptr;
	cudaIpcCloseMemHandle((void*)(uintptr_t)p);
}

void swgldev_set_option(swgldev_ctx* c, const char* name, int64_t value)
{
	swgldev_sync(c);
	if (!strcmp(name, "fuse_clear")) c->opt_fuse_clear = (int)value;
	else if (!strcmp(name, "count_fragments")) c->opt_count_fragments = (int)value;
	else if (!strcmp(name, "raster_path")) c->opt_raster_path = (int)value;
	else if (!strcmp(name, "stage_timing"))
	{
		c->opt_stage_timing = (int)value;
		for (int i = 0; i < 8; i++) c->stage_us[i] = 0.0;
		c->stage_draws = 0;
	}
}

int64_t swgldev_get_option(swgldev_ctx* c, const char* name)
{
	if (!strcmp(name, "fuse_clear")) return c->opt_fuse_clear;
	if (!strcmp(name, "count_fragments")) return c->opt_count_fragments;
	if (!strcmp(name, "raster_path")) return c->opt_raster_path;
	if (!strcmp(name, "last_raster_path")) return c->last_raster_path;
	if (!strcmp(name, "tile_size")) return SWGL_TILE;
	if (!strcmp(name, "kernel_launches")) return (int64_t)c->n_launches;
	if (!strcmp(name, "stage_draws")) return (int64_t)c->stage_draws;
	if (!strncmp(name, "stage_ns_", 9)) { int i = atoi(name + 9); return (i >= 0 && i < 5) ? (int64_t)(c->stage_us[i] * 1000.0) : -1; }
	if (!strcmp(name, "device")) return c->device;
	if (!strcmp(name, "sizeof_draw_params")) return (int64_t)sizeof(DrawParams);
	return -1;
}

} /* extern "C" */
