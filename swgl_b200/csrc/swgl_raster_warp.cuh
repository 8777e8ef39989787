/*
 * swgl_raster_warp.cuh -- per-tile rasteriser, warp-autonomous form (included by swgl_dev.cu).
 *
 * One WARP owns one 32x8 tile from its list to its write-back; there is no block-level barrier
 * anywhere after the prologue.  Per batch of 32 primitives (ascending primitive id):
 *
 *   A  lane = primitive: gather the primitive, replay the span walk over the tile's 8 rows
 *      (swgl.c:3356-3361, 3466-3471) with each row's span in a register, and stage the constants
 *      of the fragment arithmetic (prim_consts) in shared memory;
 *   S  one warp shuffle scan of both counts, then every non-empty span is appended to the batch's
 *      span list (first fragment, row, first column, primitive lane) and its first fragment is
 *      marked in a bit map (one word per step of phase B);
 *   B  lane = fragment, dense steps of 32: the step's word of the bit map and a population count
 *      give the span of each lane (no search), Barycentric + perspective + z (swgl.c:3365-3382)
 *      with shared reciprocals (frag_weights_fast), fragment shader, then the ordered part: lanes
 *      that hit the same pixel are found with __match_any_sync and commit (depth test + blend,
 *      swgl.c:3387-3462) in lane order = submission order.
 *
 * Compared with the CTA-per-32x32-tile kernel (k_raster_frag) this trades ~40 % more
 * (tile, primitive) pairs for: no __syncthreads in the hot loop, shuffle scans instead of block
 * scans, 32-element shuffle sorts instead of 256-element merges, and no fragment->primitive map.
 */
#ifndef SWGL_RASTER_WARP_CUH
#define SWGL_RASTER_WARP_CUH

#ifndef WT_FIRST_TURN_FAST
#define WT_FIRST_TURN_FAST 1
#endif
#ifndef WT_PREFETCH_LEAN
#define WT_PREFETCH_LEAN 0   /* measured: prefetching the element indices of the next batch's record-less primitives 142.3 -> 144.3 us */
#endif
#ifndef WT_LUT
#define WT_LUT 0             /* 1: byte / 255 through a shared-memory table (round 1); 0: arithmetically (byte_over_255) */
#endif
#ifndef WT_MAGIC_I2F
#define WT_MAGIC_I2F 0   /* measured: no gain (tools/ab_bench.sh, r02) */
#endif
/* Tiles are 32 pixels wide and TH = 8, 4 or 2 rows tall (template parameter THS = log2 TH).  8 rows is the
 * choice whenever the framebuffer has at least one resident wave of such tiles (148 SMs x 32 warps); smaller
 * framebuffers -- or the share of one rank of a sort-first group -- get shorter tiles, i.e. more and shorter
 * warps, because a tile is one warp's serial work and the slowest warp sets the kernel's time. */
#if WT_LUT
#define WT_BLEND(col, cur) blend_pack_lut((col).x, (col).y, (col).z, (col).w, (cur), S.lut)
#else
#define WT_BLEND(col, cur) blend_pack_arith((col).x, (col).y, (col).z, (col).w, (cur))
#endif
#define WT_H_SHIFT  3        /* the default tile height (and the only one of run-time compiled kernels) */
#define WT_H_SHIFT_MIN 1
#define WT_WARPS    4        /* tiles per CTA */
#define WT_SORT_CAP 256      /* list entries a warp sorts in shared memory */
#define WT_MAX_FRAGS 2048                     /* fragments of one batch: a batch that would produce more is cut short */

/* span entry: (tile pixel of the first fragment - its fragment index) mod 2^13 | primitive lane << 21
 * | primitive passed prim_fast_ok() << 26 */
#define WT_PC_WORDS (4 * PC_VEC4)             /* per-primitive constants of phase B, see prim_consts() */
template <int TH>
struct WarpTile
{
	uint32_t color[SWGL_TILE * TH];
	float    depth[SWGL_TILE * TH];
	union
	{
		uint32_t ids[WT_SORT_CAP];               /* scratch of the list sort (before the first batch) */
		struct
		{
			uint32_t span[32 * TH];              /* this batch's non-empty (primitive, row) spans in fragment order */
			uint32_t start_bits[WT_MAX_FRAGS / 32];  /* bit f: fragment f is the first of a span */
		} b;
	} u;
	float4   pc[PC_VEC4 * PC_STRIDE];        /* [field][primitive lane]: constants staged by phase A (prim_consts) */
};

template <int TH>
struct WarpShared
{
#if WT_LUT
	float    lut[256];             /* byte / 255.0f (swgl.c:3434-3437) */
#endif
	WarpTile<TH> w[WT_WARPS];
};

/* lane masks from the special registers: one instruction where the compiler re-derives (1 << lane) - 1 from the
 * thread index on every use to save a register */
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ uint32_t lanemask_le() { uint32_t m; asm("mov.u32 %0, %%lanemask_le;" : "=r"(m)); return m; }

/* ascending sort of one value per lane (padding 0xffffffff sinks to the top lanes) */
__device__ __forceinline__ uint32_t warp_sort32(uint32_t x, uint32_t lane)
{
#pragma unroll
	for (uint32_t k = 2; k <= 32; k <<= 1)
#pragma unroll
		for (uint32_t j = k >> 1; j > 0; j >>= 1)
		{
			const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
			const bool keep_min = ((lane & k) == 0) == ((lane & j) == 0);
			x = keep_min ? min(x, y) : max(x, y);
		}
	return x;
}

/* a bitonic sequence of 32 (one value per lane) -> ascending */
__device__ __forceinline__ uint32_t warp_bitonic_clean(uint32_t x, uint32_t lane)
{
#pragma unroll
	for (uint32_t j = 16; j > 0; j >>= 1)
	{
		const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
		x = (lane & j) ? max(x, y) : min(x, y);
	}
	return x;
}

/* two ascending runs of 32 (a, b) -> one ascending run of 64 (a = lower half, b = upper half) */
__device__ __forceinline__ void warp_merge64(uint32_t& a, uint32_t& b, uint32_t lane)
{
	const uint32_t rb = __shfl_sync(0xffffffffu, b, 31u - lane);
	const uint32_t lo = min(a, rb), hi = max(a, rb);
	a = warp_bitonic_clean(lo, lane);
	b = warp_bitonic_clean(hi, lane);
}

/* In-place ascending sort of n ids by one warp (shared or global memory).  Bitonic network in its
 * "flip + half-cleaner" form: every compare-exchange moves the smaller key to the lower index, so
 * the virtual 0xffffffff padding beyond n never has to move and n need not be a power of two. */
__device__ __noinline__ void warp_sort_mem(uint32_t* ids, uint32_t n, uint32_t lane)
{
	uint32_t n_pow2 = 1; while (n_pow2 < n) n_pow2 <<= 1;
	for (uint32_t k = 2; k <= n_pow2; k <<= 1)
	{
		for (uint32_t i = lane; i < n; i += 32)
		{
			const uint32_t l = i ^ (k - 1u);
			if (l > i && l < n) { const uint32_t a = ids[i], b = ids[l]; if (a > b) { ids[i] = b; ids[l] = a; } }
		}
		__syncwarp();
		for (uint32_t j = k >> 2; j > 0; j >>= 1)
		{
			for (uint32_t i = lane; i < n; i += 32)
			{
				const uint32_t l = i ^ j;
				if (l > i && l < n) { const uint32_t a = ids[i], b = ids[l]; if (a > b) { ids[i] = b; ids[l] = a; } }
			}
			__syncwarp();
		}
	}
}

/* N (8, 4 or 2) consecutive pixels of one tile row: load (or take the pending clear value) */
template <int N>
__device__ __forceinline__ void wt_load(const DrawParams& P, const ClearParams& cp, int row, int px0,
                                        uint32_t* c8, float* d8, bool& touched)
{
#pragma unroll
	for (int k = 0; k < N; k++) { c8[k] = 0u; d8[k] = 0.0f; }
	if (row >= (int)P.H) return;
	const size_t pix = (size_t)row * P.W + (size_t)px0;
	const bool in_row = cp.flags && row >= cp.y0 && row < cp.y1;
	const bool all_in = in_row && px0 >= cp.x0 && px0 + N <= cp.x1;
	const bool none_in = !in_row || px0 + N <= cp.x0 || px0 >= cp.x1;
	const bool need_c = !(all_in && (cp.flags & 1u)), need_d = !(all_in && (cp.flags & 2u));
	if (N >= 4 && px0 + N - 1 < (int)P.W && ((pix & 3u) == 0))
	{
#pragma unroll
		for (int q = 0; q < N / 4; q++)
		{
			if (need_c) { const uint4 a = *(const uint4*)(P.color + pix + 4 * q); c8[4 * q] = a.x; c8[4 * q + 1] = a.y; c8[4 * q + 2] = a.z; c8[4 * q + 3] = a.w; }
			if (need_d) { const float4 a = *(const float4*)(P.depth + pix + 4 * q); d8[4 * q] = a.x; d8[4 * q + 1] = a.y; d8[4 * q + 2] = a.z; d8[4 * q + 3] = a.w; }
		}
	}
	else if (N == 2 && px0 + 1 < (int)P.W && ((pix & 1u) == 0))
	{
		if (need_c) { const uint2 a = *(const uint2*)(P.color + pix); c8[0] = a.x; c8[1] = a.y; }
		if (need_d) { const float2 a = *(const float2*)(P.depth + pix); d8[0] = a.x; d8[1] = a.y; }
	}
	else
	{
#pragma unroll
		for (int k = 0; k < N; k++)
			if (px0 + k < (int)P.W) { if (need_c) c8[k] = P.color[pix + k]; if (need_d) d8[k] = P.depth[pix + k]; }
	}
	if (!none_in)
	{
#pragma unroll
		for (int k = 0; k < N; k++)
		{
			const bool inside = (px0 + k) >= cp.x0 && (px0 + k) < cp.x1;
			if (inside && (cp.flags & 1u)) c8[k] = cp.word;
			if (inside && (cp.flags & 2u)) d8[k] = 0.0f;
			touched |= inside;
		}
	}
}

/* frag_weights() for a fragment outside the fast domain, from the list entry: out of line, the hot loop
 * only carries the call */
__device__ __noinline__ float4 weights_slow_entry(const DrawParams& P, uint32_t entry, float px, float py)
{
	const Prim q = load_prim(P, entry);
	BaryConst k;
	bary_setup(q.v[0], q.v[1], q.v[2], k);
	float4 r;
	frag_weights(k, px, py, r.x, r.y, r.z, r.w);
	return r;
}

/* Shading of a fragment that failed the depth test before the ordered part and passes in it (an
 * earlier fragment of the same step stored exactly 0.0 = "empty"): rare, kept out of the hot loop. */
template <int FS>
__device__ __noinline__ float4 shade_late(const DrawParams& P, uint32_t pid, int tile_x0, int band_last_y, uint32_t pix)
{
	/* the pixel's coordinates are worked out here, not by the caller: the hot loop only passes what it has */
	const float px = (float)(tile_x0 + (int)(pix & (SWGL_TILE - 1))), py = (float)(band_last_y - (int)(pix >> SWGL_TILE_SHIFT));
	const Prim qv = load_prim(P, pid);
	const Prim* q = &qv;
	BaryConst k;
	bary_setup(q->v[0], q->v[1], q->v[2], k);
	FragIn fi;
	float z2;
	frag_weights(k, px, py, fi.u, fi.v, fi.w, z2);
	fi.vid0 = q->vid[0]; fi.vid1 = q->vid[1]; fi.vid2 = q->vid[2];
	fi.a = P.vary + (size_t)fi.vid0 * P.nvf + P.fs_slot;
	fi.b = P.vary + (size_t)fi.vid1 * P.nvf + P.fs_slot;
	fi.c = P.vary + (size_t)fi.vid2 * P.nvf + P.fs_slot;
	fi.stride = 1;
	fi.lod = ((FS == SWFS_GENERIC || FS == SWFS_JIT) && P.mip_lod) ? mip_level(q->v[0].x, q->v[0].y, q->v[1].x, q->v[1].y, q->v[2].x, q->v[2].y) : 0.0f;
	return clamp_color(run_fragment<FS>(P, fi));
}

template <int FS, int THS>
__global__ void __launch_bounds__(WT_WARPS * 32, 8) k_raster_warp(const __grid_constant__ DrawParams P)
{
	constexpr int TH = 1 << THS;                 /* tile rows */
	__shared__ WarpShared<TH> S;
	const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	/* the grid covers the tile rows this rank owns: owned row i -> tile row (sort-first bands of
	 * band_rows x 32 framebuffer rows dealt round-robin, see owns_tile_row) */
	const uint32_t tx = blockIdx.x * WT_WARPS + wid;     /* grid: (tile columns / WT_WARPS, owned tile rows) */
	uint32_t ty = blockIdx.y;
	bool mine = tx < P.tiles_x;
	if (P.n_ranks > 1)
	{
		const uint32_t per = P.band_rows << (5u - THS);      /* tile rows per ownership group */
		ty = ((ty / per) * P.n_ranks + P.rank) * per + ty % per;
		mine = mine && ty < P.tiles_y && owns_tile_row(P, ty);
	}
	const uint32_t tile = ty * P.tiles_x + tx;
	/* the list length is requested first: its round trip runs under the table load and the barrier */
	uint32_t n_raw = 0;
	if (lane == 0 && mine) n_raw = *(volatile const uint32_t*)(P.tile_count + tile);
#if WT_LUT
	for (uint32_t i = threadIdx.x; i < 256; i += WT_WARPS * 32) S.lut[i] = __ldg(P.lut255 + i);
	__syncthreads();            /* the only block-level barrier */
#endif
	if (!mine) return;
	WarpTile<TH>& T = S.w[wid];

	/* the cursor is re-armed for the next draw */
	if (lane == 0)
	{
		if (n_raw) P.tile_count[tile] = 0u;
		if (n_raw > P.bin_cap) atomicMax(&P.ctr->max_list, n_raw);
		if (n_raw && P.count_fragments && (n_raw <= P.bin_cap || P.ov_cap)) atomicAdd(&P.ctr->pair_total, (unsigned long long)n_raw);
	}
	n_raw = __shfl_sync(0xffffffffu, n_raw, 0);
	if (P.ctr->overflow || P.diag) return;
	/* n_raw > K: a deep tile, its list continues in the overflow pool (assembled below); without a pool the
	 * draw has been flagged and never gets here */
	uint32_t n_list = (n_raw > P.bin_cap && !P.ov_cap) ? P.bin_cap : n_raw;
	const ClearParams cp = P.clear;
	if (n_list == 0 && !cp.flags) return;

	const int tile_x0 = (int)(tx << SWGL_TILE_SHIFT), tile_r0 = (int)(ty << THS);
	const int band_last_y = P.ytop - tile_r0;                /* raster y of tile row 0 */
	const int band_first_y = band_last_y - (TH - 1);

	/* ---- stage the tile: every lane TH consecutive pixels (32 / TH lanes per row) ---- */
	bool dirty = false;
	{
		const int r = (int)(lane >> (5 - THS)), seg = (int)(lane & ((32u >> THS) - 1u));
		const int px0 = tile_x0 + (seg << THS);
		const uint32_t s = (uint32_t)r * SWGL_TILE + (uint32_t)(seg << THS);
		/* the usual frame: a pending clear of both attachments covers the whole tile, nothing is read */
		const bool whole = (cp.flags & 3u) == 3u && cp.x0 <= tile_x0 && cp.x1 >= tile_x0 + SWGL_TILE
		                   && cp.y0 <= tile_r0 && cp.y1 >= tile_r0 + TH
		                   && tile_x0 + SWGL_TILE <= (int)P.W && tile_r0 + TH <= (int)P.H;
		uint32_t c8[TH]; float d8[TH];
		if (whole)
		{
#pragma unroll
			for (int k = 0; k < TH; k++) { c8[k] = cp.word; d8[k] = 0.0f; }
			dirty = true;
		}
		else wt_load<TH>(P, cp, tile_r0 + r, px0, c8, d8, dirty);
		if (TH >= 4)
		{
#pragma unroll
			for (int q = 0; q < TH / 4; q++)
			{
				*(uint4*)&T.color[s + 4 * q] = make_uint4(c8[4 * q], c8[4 * q + 1], c8[4 * q + 2], c8[4 * q + 3]);
				*(float4*)&T.depth[s + 4 * q] = make_float4(d8[4 * q], d8[4 * q + 1], d8[4 * q + 2], d8[4 * q + 3]);
			}
		}
		else
		{
			*(uint2*)&T.color[s] = make_uint2(c8[0], c8[1]);
			*(float2*)&T.depth[s] = make_float2(d8[0], d8[1]);
		}
	}
	__syncwarp();

	uint32_t n_tested = 0, n_shaded = 0;
	/* column limits of a row's span folded per tile (fast row walk of phase A): viewport, framebuffer
	 * and tile; exact when the viewport limits are integers a float holds exactly */
	const bool tile_fast = fabsf(P.fvx) < 16777216.0f && P.xlimit < 16777216.0f;
	const int t_lo = max(max((int)P.fvx, 0), tile_x0);
	const int t_hi = min(min((int)ceilf(P.xlimit), (int)P.W), tile_x0 + SWGL_TILE);
	if (n_list > 0)
	{
		/* ---- ascending primitive id = submission order ---- */
		uint32_t* gl_ids = P.pairs + (size_t)tile * P.bin_cap;
		if (n_list > P.bin_cap)
		{
			/* deep tile: K inline entries + this tile's entries of the pool, gathered into hot_store (the set-up
			 * kernels have checked that the pool and hot_store are large enough) and sorted there */
			uint32_t hb = 0;
			if (lane == 0) hb = atomicAdd(&P.ctr->hot_cursor, n_raw);
			hb = __shfl_sync(0xffffffffu, hb, 0);
			uint32_t* dst = P.hot_store + hb;
			for (uint32_t i = lane; i < P.bin_cap; i += 32) dst[i] = gl_ids[i];
			uint32_t wr = P.bin_cap;
			const uint32_t tot = min(P.ctr->ov_cursor, P.ov_cap);
			for (uint32_t i0 = 0; i0 < tot; i0 += 32)
			{
				const uint32_t i = i0 + lane;
				uint2 e = make_uint2(0xffffffffu, 0u);
				if (i < tot) e = P.ov_pool[i];
				const bool mine = e.x == tile;
				const uint32_t bal = __ballot_sync(0xffffffffu, mine);
				if (mine) dst[wr + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = e.y;
				wr += (uint32_t)__popc(bal);
			}
			__syncwarp();
			n_list = min(wr, n_raw);
			gl_ids = dst;
		}
		const uint32_t* sorted = gl_ids;
		uint32_t xs0 = 0xffffffffu, xs1 = 0xffffffffu, xs2 = 0xffffffffu, xs3 = 0xffffffffu;
		if (n_list <= 128)
		{
			/* up to four runs of 32 sorted with the shuffle network and merged with bitonic merges:
			 * the whole list stays in registers (element 32 * c + lane in xs[c]) */
			xs0 = warp_sort32(lane < n_list ? gl_ids[lane] : 0xffffffffu, lane);
			if (n_list > 32)
			{
				xs1 = warp_sort32(32u + lane < n_list ? gl_ids[32u + lane] : 0xffffffffu, lane);
				warp_merge64(xs0, xs1, lane);
			}
			if (n_list > 64)
			{
				xs2 = warp_sort32(64u + lane < n_list ? gl_ids[64u + lane] : 0xffffffffu, lane);
				if (n_list > 96)
				{
					xs3 = warp_sort32(96u + lane < n_list ? gl_ids[96u + lane] : 0xffffffffu, lane);
					warp_merge64(xs2, xs3, lane);
				}
				/* 64 + 64: reverse the upper run, compare-exchange, then clean both bitonic halves */
				const uint32_t r2 = __shfl_sync(0xffffffffu, xs3, 31u - lane), r3 = __shfl_sync(0xffffffffu, xs2, 31u - lane);
				const uint32_t l0 = min(xs0, r2), l1 = min(xs1, r3), h0 = max(xs0, r2), h1 = max(xs1, r3);
				xs0 = warp_bitonic_clean(min(l0, l1), lane); xs1 = warp_bitonic_clean(max(l0, l1), lane);
				xs2 = warp_bitonic_clean(min(h0, h1), lane); xs3 = warp_bitonic_clean(max(h0, h1), lane);
			}
		}
		else if (n_list <= WT_SORT_CAP)
		{
			for (uint32_t i = lane; i < n_list; i += 32) T.u.ids[i] = gl_ids[i];
			__syncwarp();
			warp_sort_mem(T.u.ids, n_list, lane);
			for (uint32_t i = lane; i < n_list; i += 32) gl_ids[i] = T.u.ids[i];   /* the scratch is reused by the batches */
			__syncwarp();
		}
		else warp_sort_mem(gl_ids, n_list, lane);

		uint32_t take = 32;                      /* primitives consumed by the batch */
		for (uint32_t base = 0; base < n_list; base += take)
		{
			const uint32_t nb = min(32u, n_list - base);
			take = 32;
			/* ---- phase A: lane = primitive.  Replay the walk over the tile's rows; row r's span
			 * (first column | length << 8) stays in a register ---- */
			uint32_t pid;
			if (n_list > 128) pid = lane < nb ? sorted[base + lane] : 0xffffffffu;
			else if ((base & 31u) == 0) pid = base == 0 ? xs0 : base == 32 ? xs1 : base == 64 ? xs2 : xs3;
			else
			{   /* after a batch that was cut short: element base + lane of the register-resident list */
				const uint32_t i = base + lane, sl = i & 31u, sr = i >> 5;
				const uint32_t v0 = __shfl_sync(0xffffffffu, xs0, sl), v1 = __shfl_sync(0xffffffffu, xs1, sl);
				const uint32_t v2 = __shfl_sync(0xffffffffu, xs2, sl), v3 = __shfl_sync(0xffffffffu, xs3, sl);
				pid = sr == 0 ? v0 : sr == 1 ? v1 : sr == 2 ? v2 : sr == 3 ? v3 : 0xffffffffu;
			}
			uint32_t cnt = 0, nsp = 0, fast = 0;
			uint32_t row0 = 0;                   /* tile row of the primitive's first walked row; sp[k] is the span k rows above it */
			uint32_t sp[TH];
#pragma unroll
			for (int r = 0; r < TH; r++) sp[r] = 0u;
			if (lane < nb)
			{
				const PrimRef q = prim_ref(P, pid);
				const float4 a = *q.a, b = *q.b, c = *q.c;
				const uint32_t band = q.band;
				TriWalk w;
				tri_setup(a, b, c, P, w);
				const int y_in = max(w.ys, band_first_y), y_out = min(w.ye - 1, band_last_y);
				if (y_out >= y_in)
				{
					row0 = (uint32_t)(band_last_y - y_in);
					fast = prim_fast_ok(a, b, c) ? 1u : 0u;
					/* built-in shapes: the three words are the float offsets of the consumed varying in the packed
					 * records (the host has checked that they fit 32 bits); IR shaders get the record ids */
					if (FS == SWFS_VARYING || FS == SWFS_TEXTURE)
						prim_consts(a, b, c, q.vid0 * P.nvf + P.fs_slot, q.vid1 * P.nvf + P.fs_slot, q.vid2 * P.nvf + P.fs_slot, pid, &T.pc[lane]);
					else
						prim_consts(a, b, c, q.vid0, q.vid1, q.vid2, pid, &T.pc[lane]);
					float x0, x1, s1;
					bool switched;
					walk_to_row(P, w, band, ty, y_in, x0, x1, s1, switched);
					if (fast && tile_fast)
					{
						/* Bounded coordinates (prim_fast_ok) and a viewport whose limits are exact in float: the
						 * row's span is trunc(lo) and ceil(hi) clamped once against limits folded per tile --
						 * trunc and ceil are monotonic, so they commute with the reference's MAX/MIN against
						 * the viewport (swgl.c:3358-3361) -- and the edge switch is an integer compare. */
						const int c1yi = (int)w.c1y;
#pragma unroll
						for (int k = 0; k < TH; k++)
						{
							const int y = y_in + k;
							if (y <= y_out)
							{
								const int xa = min(max(__float2int_rz(fminf(x0, x1)), t_lo), tile_x0 + SWGL_TILE);
								const int xb = max(min(__float2int_ru(fmaxf(x0, x1)), t_hi), tile_x0);
								if (xb > xa)
								{
									sp[k] = (uint32_t)(xa - tile_x0) | ((uint32_t)(xb - xa) << 8);
									cnt += (uint32_t)(xb - xa);
									nsp++;
								}
								if (!switched && y + 1 >= c1yi) { switched = true; s1 = w.s2; x1 = w.c1x; }
								x0 += w.s0; x1 += s1;
							}
						}
					}
					else
					{
#pragma unroll
						for (int k = 0; k < TH; k++)
						{
							const int y = y_in + k;
							if (y <= y_out)
							{
								int xa, xb;
								row_span(x0, x1, P, xa, xb);
								xa = min(max(xa, tile_x0), tile_x0 + SWGL_TILE) - tile_x0;
								xb = min(max(xb, tile_x0), tile_x0 + SWGL_TILE) - tile_x0;
								if (xb > xa)
								{
									sp[k] = (uint32_t)xa | ((uint32_t)(xb - xa) << 8);
									cnt += (uint32_t)(xb - xa);
									nsp++;
								}
								if (!switched && (float)y + 1.0f >= w.c1y) { switched = true; s1 = w.s2; x1 = w.c1x; }
								x0 += w.s0; x1 += s1;
							}
						}
					}
				}
			}
			/* ---- S: one exclusive scan for fragment counts (low half) and span counts (high half) ---- */
			const uint32_t both = cnt | (nsp << 16);
			uint32_t incl = both;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += y; }
			const uint32_t excl = incl - both;
			const uint32_t last = __shfl_sync(0xffffffffu, incl, 31);
			uint32_t total = last & 0xffffu;             /* at most 32 x 256 = 8192 fragments, 256 spans */
			uint32_t excl_k = excl;
			if (total > WT_MAX_FRAGS)
			{
				/* more fragments than the start-bit map holds: keep the leading primitives that fit (at
				 * least 8, a primitive has at most 256 fragments here) and start the next batch behind them */
				take = (uint32_t)__popc(__ballot_sync(0xffffffffu, (incl & 0xffffu) <= WT_MAX_FRAGS));
				total = __shfl_sync(0xffffffffu, incl, take - 1u) & 0xffffu;
				if (lane >= take)
				{
#pragma unroll
					for (int r = 0; r < TH; r++) sp[r] = 0u;
				}
			}
			n_tested += (lane == 0) ? total : 0u;
			if (total == 0) continue;
			for (uint32_t j = lane; (j << 5) < total; j += 32) T.u.b.start_bits[j] = 0u;
			__syncwarp();
			{
				uint32_t run = excl_k & 0xffffu, k = excl_k >> 16;
				const uint32_t tag = (lane << 21) | (fast << 26);
#pragma unroll
				for (int j = 0; j < TH; j++)
				{
					const uint32_t len = sp[j] >> 8;
					if (len)
					{
						/* low 13 bits: tile pixel of the span's first fragment minus its fragment index (mod 2^13) */
						T.u.b.span[k++] = (((row0 - (uint32_t)j) * SWGL_TILE + (sp[j] & 31u) - run) & 0x1fffu) | tag;
						atomicOr(&T.u.b.start_bits[run >> 5], 1u << (run & 31u));
						run += len;
					}
				}
			}
			__syncwarp();

			/* the next batch's primitive records are requested now and arrive while this batch is shaded */
			if (base + 32u < n_list && take == 32u && (base & 31u) == 0)
			{
				const uint32_t nxt = (n_list <= 128) ? (base == 0 ? xs1 : base == 32 ? xs2 : xs3)
				                                     : (base + 32u + lane < n_list ? sorted[base + 32u + lane] : 0xffffffffu);
				if (nxt != 0xffffffffu && (nxt & 1u))
				{
					const Prim* q = prim_at(P, nxt >> 1);
					asm volatile("prefetch.global.L1 [%0];" :: "l"(q));
					asm volatile("prefetch.global.L1 [%0];" :: "l"((const char*)q + 32));
				}
#if WT_PREFETCH_LEAN
				else if (nxt != 0xffffffffu && P.ibo)
				{
					/* record-less primitive: the first link of its gather chain (element indices -> vertices) */
					const uint32_t* ix = P.ibo + ((unsigned long long)(long long)P.first + 3u * (nxt >> 2));
					asm volatile("prefetch.global.L1 [%0];" :: "l"(ix));
				}
#endif
			}

			/* ---- phase B: lane = fragment, dense steps of 32 in submission order ---- */
			uint32_t spans_before = 0;           /* spans that start before this step */
			for (uint32_t t0 = 0; t0 < total; t0 += 32)
			{
				const uint32_t t = t0 + lane;
				const bool active = t < total;
				/* the span of fragment t: the last one that starts at or before it */
				const uint32_t starts = T.u.b.start_bits[t0 >> 5];
				const uint32_t e = T.u.b.span[spans_before + (uint32_t)__popc(starts & lanemask_le()) - 1u];
				spans_before += (uint32_t)__popc(starts);
				const float4* pc = &T.pc[(e >> 21) & 31u];
				const uint32_t tpix = (e + t) & 0x1fffu;         /* pixel of the tile: row * 32 + column */
				const uint32_t r = tpix >> SWGL_TILE_SHIFT, lx = tpix & (SWGL_TILE - 1u);
				const uint32_t pix = active ? tpix : (0x80000000u | lane);
				/* lanes on the same pixel commit in lane order; the match is issued before the arithmetic
				 * it does not depend on */
				const uint32_t peers = __match_any_sync(0xffffffffu, pix);
				bool pending = active;
				float z = 0.0f;
				float4 col = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				bool shaded_early = false;
				if (active)
				{
					float u, v, w;
#if WT_MAGIC_I2F
					/* pixel coordinates are below 2^23 here (tiles_x <= 2047, |y| < 2^22 is checked by the host): the
					 * int -> float conversion is an OR into the mantissa of 2^23 and one exact subtraction */
					const float px = __int_as_float(0x4b000000 | (tile_x0 + (int)lx)) - 8388608.0f;
					const float py = (float)(band_last_y - (int)r);
#else
					const float px = (float)(tile_x0 + (int)lx), py = (float)(band_last_y - (int)r);
#endif
					if (!(frag_weights_fast(pc, px, py, u, v, w, z) && ((e >> 26) & 1u)))
					{
						const float4 s4 = weights_slow_entry(P, __float_as_uint(pc[5 * PC_STRIDE].w), px, py);
						u = s4.x; v = s4.y; w = s4.z; z = s4.w;
					}
					/* the fragment shader does not read the framebuffer: run it before the ordered
					 * part unless the fragment already fails against the stored depth (then it can only
					 * pass later if an earlier fragment stores exactly 0.0 = "empty": shaded late) */
					const float cur = T.depth[pix];
					if (cur == 0.0f || cur >= z)
					{
						float4 va = make_float4(0.0f, 0.0f, 0.0f, 0.0f), vb = va, vc = va;
						if (FS != SWFS_GENERIC && FS != SWFS_JIT)
						{
							const float4 ids = pc[5 * PC_STRIDE];
							const float* pa = P.vary + __float_as_uint(ids.x);
							const float* pb = P.vary + __float_as_uint(ids.y);
							const float* pv = P.vary + __float_as_uint(ids.z);
							if (FS == SWFS_VARYING)
							{
								/* the host lays vec4 varyings out on 16-byte boundaries (the device layer demotes the draw to
								 * the IR path otherwise): three 128-bit loads, no per-component alternative in the loop */
								va = __ldg((const float4*)pa); vb = __ldg((const float4*)pb); vc = __ldg((const float4*)pv);
							}
							else
							{
								va.x = __ldg(pa + P.fs_swz_u); va.y = __ldg(pa + P.fs_swz_v);
								vb.x = __ldg(pb + P.fs_swz_u); vb.y = __ldg(pb + P.fs_swz_v);
								vc.x = __ldg(pv + P.fs_swz_u); vc.y = __ldg(pv + P.fs_swz_v);
							}
						}
						if (FS == SWFS_VARYING)      /* InterpolateLinearEx (swgl.c:3270-3297) */
							col = make_float4(va.x * u + vb.x * v + vc.x * w, va.y * u + vb.y * v + vc.y * w,
							                  va.z * u + vb.z * v + vc.z * w, va.w * u + vb.w * v + vc.w * w);
						else if (FS == SWFS_TEXTURE)
							col = sample_nearest(P.tex[P.fs_tex_unit], va.x * u + vb.x * v + vc.x * w, va.y * u + vb.y * v + vc.y * w);
						else
						{
							const float4 ids = pc[5 * PC_STRIDE];
							FragIn fi;
							fi.u = u; fi.v = v; fi.w = w;
							fi.vid0 = __float_as_uint(ids.x); fi.vid1 = __float_as_uint(ids.y); fi.vid2 = __float_as_uint(ids.z);
							fi.a = P.vary + (size_t)fi.vid0 * P.nvf + P.fs_slot;
							fi.b = P.vary + (size_t)fi.vid1 * P.nvf + P.fs_slot;
							fi.c = P.vary + (size_t)fi.vid2 * P.nvf + P.fs_slot;
							fi.stride = 1;
							fi.lod = P.mip_lod ? prim_lod(P, __float_as_uint(ids.w)) : 0.0f;
							col = run_fragment<FS>(P, fi);
						}
						col = clamp_color(col);
						shaded_early = true;
					}
				}
				/* ---- ordered commit ---- */
				const uint32_t my_turn = (uint32_t)__popc(peers & lanemask_lt());
				const uint32_t turns = __reduce_max_sync(0xffffffffu, my_turn);
#if WT_FIRST_TURN_FAST
				/* the first fragment of each pixel group sees the depth it was tested against above (nothing has
				 * written the pixel since): its test is decided, no reload, no late shading */
				__syncwarp();                /* the early depth reads of the later fragments of a group come first */
				if (pending && my_turn == 0u)
				{
					if (shaded_early)
					{
						T.depth[pix] = z;
						n_shaded++;
						T.color[pix] = WT_BLEND(col, T.color[pix]);
						dirty = true;
					}
					pending = false;
				}
				__syncwarp();
				for (uint32_t turn = 1; turn <= turns; turn++)
#else
				for (uint32_t turn = 0; turn <= turns; turn++)
#endif
				{
					if (pending && my_turn == turn)
					{
						const float cur = T.depth[pix];
						if (cur == 0.0f || cur >= z)     /* swgl.c:3387 */
						{
							T.depth[pix] = z;
							n_shaded++;
							if (!shaded_early)
								col = shade_late<FS>(P, __float_as_uint(pc[5 * PC_STRIDE].w), tile_x0, band_last_y, pix);
							T.color[pix] = WT_BLEND(col, T.color[pix]);
							dirty = true;
						}
						pending = false;
					}
					__syncwarp();
				}
			}
			__syncwarp();
		}
	}

	/* ---- write-back: each store instruction covers four whole 128-byte row segments (the mirror
	 * target may be a peer GPU or pinned host memory, where partial lines cost full transactions) ---- */
	if (__any_sync(0xffffffffu, dirty))
	{
		const int px0 = tile_x0 + (int)((lane & 7u) << 2);
#pragma unroll
		for (int i = 0; i < (TH + 3) / 4; i++)
		{
			const int r = (int)(lane >> 3) + 4 * i, row = tile_r0 + r;
			if (r >= TH || row >= (int)P.H) continue;
			const size_t pixg = (size_t)row * P.W + (size_t)px0;
			const uint32_t s = (uint32_t)r * SWGL_TILE + ((lane & 7u) << 2);
			const uint4 c4 = *(const uint4*)&T.color[s];
			float4 d4 = *(const float4*)&T.depth[s];
			d4.x = canon_nan(d4.x); d4.y = canon_nan(d4.y); d4.z = canon_nan(d4.z); d4.w = canon_nan(d4.w);
			if (px0 + 3 < (int)P.W && ((pixg & 3u) == 0))
			{
				*(uint4*)(P.color + pixg) = c4;
				*(float4*)(P.depth + pixg) = d4;
				if (P.peer_color) *(uint4*)(P.peer_color + pixg) = c4;
			}
			else
			{
				const uint32_t cc[4] = { c4.x, c4.y, c4.z, c4.w };
				const float dd[4] = { d4.x, d4.y, d4.z, d4.w };
				for (int k = 0; k < 4; k++)
					if (px0 + k < (int)P.W)
					{
						P.color[pixg + k] = cc[k]; P.depth[pixg + k] = dd[k];
						if (P.peer_color) P.peer_color[pixg + k] = cc[k];
					}
			}
		}
	}

	if (P.count_fragments)
	{
		for (int o = 16; o > 0; o >>= 1) { n_tested += __shfl_down_sync(0xffffffffu, n_tested, o); n_shaded += __shfl_down_sync(0xffffffffu, n_shaded, o); }
		if (lane == 0)
		{
			if (n_tested) atomicAdd(&P.ctr->tested[tile % SWGL_CTR_SLOTS], (unsigned long long)n_tested);
			if (n_shaded) atomicAdd(&P.ctr->shaded[tile % SWGL_CTR_SLOTS], (unsigned long long)n_shaded);
		}
	}
}

#endif
