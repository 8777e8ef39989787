/*
 * swgl_raster_warp.cuh -- per-tile rasteriser, warp-autonomous form (included by swgl_dev.cu).
 *
 * One WARP owns one 32x8 tile from its list to its write-back; there is no block-level barrier
 * anywhere after the prologue.  Per batch of 32 primitives (ascending primitive id):
 *
 *   A  lane = primitive: replay the span walk over the tile's 8 rows (swgl.c:3356-3361,
 *      3466-3471), spans to shared memory, fragment count in a register;
 *   S  warp shuffle scan of the counts;
 *   B  lane = fragment, dense steps of 32: shuffle binary search for the owning lane, row walk
 *      over at most 8 spans, Barycentric + perspective + z (swgl.c:3365-3382), fragment shader,
 *      then the ordered part: lanes that hit the same pixel are found with __match_any_sync and
 *      commit (depth test + blend, swgl.c:3387-3462) in lane order = submission order.
 *
 * Compared with the CTA-per-32x32-tile kernel (k_raster_frag) this trades ~40 % more
 * (tile, primitive) pairs for: no __syncthreads in the hot loop, shuffle scans instead of block
 * scans, 32-element shuffle sorts instead of 256-element merges, and no fragment->primitive map.
 */
#ifndef SWGL_RASTER_WARP_CUH
#define SWGL_RASTER_WARP_CUH

#define WT_H        8        /* tile rows */
#define WT_H_SHIFT  3
#define WT_PIX      (SWGL_TILE * WT_H)
#define WT_WARPS    4        /* tiles per CTA */
#define WT_SORT_CAP 256      /* list entries a warp sorts in shared memory */

struct WarpTile
{
	uint32_t color[WT_PIX];
	float    depth[WT_PIX];
	uint32_t ids[WT_SORT_CAP];
	uint16_t span[WT_H][32];       /* [row][lane]: xa | xb << 8 */
};

struct WarpShared
{
	float    lut[256];             /* byte / 255.0f (swgl.c:3434-3437) */
	WarpTile w[WT_WARPS];
};

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

/* In-place ascending sort of n ids by one warp (shared or global memory).  Bitonic network in its
 * "flip + half-cleaner" form: every compare-exchange moves the smaller key to the lower index, so
 * the virtual 0xffffffff padding beyond n never has to move and n need not be a power of two. */
__device__ __forceinline__ void warp_sort_mem(uint32_t* ids, uint32_t n, uint32_t lane)
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

/* 8 consecutive pixels of one tile row: load (or take the pending clear value) */
__device__ __forceinline__ void wt_load8(const DrawParams& P, const ClearParams& cp, int row, int px0,
                                         uint32_t* c8, float* d8, bool& touched)
{
	for (int k = 0; k < 8; k++) { c8[k] = 0u; d8[k] = 0.0f; }
	if (row >= (int)P.H) return;
	const size_t pix = (size_t)row * P.W + (size_t)px0;
	const bool in_row = cp.flags && row >= cp.y0 && row < cp.y1;
	const bool all_in = in_row && px0 >= cp.x0 && px0 + 8 <= cp.x1;
	const bool none_in = !in_row || px0 + 8 <= cp.x0 || px0 >= cp.x1;
	const bool need_c = !(all_in && (cp.flags & 1u)), need_d = !(all_in && (cp.flags & 2u));
	if (px0 + 7 < (int)P.W && ((pix & 3u) == 0))
	{
		if (need_c)
		{
			const uint4 a = *(const uint4*)(P.color + pix), b = *(const uint4*)(P.color + pix + 4);
			c8[0] = a.x; c8[1] = a.y; c8[2] = a.z; c8[3] = a.w; c8[4] = b.x; c8[5] = b.y; c8[6] = b.z; c8[7] = b.w;
		}
		if (need_d)
		{
			const float4 a = *(const float4*)(P.depth + pix), b = *(const float4*)(P.depth + pix + 4);
			d8[0] = a.x; d8[1] = a.y; d8[2] = a.z; d8[3] = a.w; d8[4] = b.x; d8[5] = b.y; d8[6] = b.z; d8[7] = b.w;
		}
	}
	else
		for (int k = 0; k < 8; k++)
			if (px0 + k < (int)P.W) { if (need_c) c8[k] = P.color[pix + k]; if (need_d) d8[k] = P.depth[pix + k]; }
	if (!none_in)
		for (int k = 0; k < 8; k++)
		{
			const bool inside = (px0 + k) >= cp.x0 && (px0 + k) < cp.x1;
			if (inside && (cp.flags & 1u)) c8[k] = cp.word;
			if (inside && (cp.flags & 2u)) d8[k] = 0.0f;
			touched |= inside;
		}
}

template <int FS>
__global__ void __launch_bounds__(WT_WARPS * 32) k_raster_warp(const __grid_constant__ DrawParams P)
{
	__shared__ WarpShared S;
	const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	for (uint32_t i = threadIdx.x; i < 256; i += WT_WARPS * 32) S.lut[i] = (float)i / 255.0f;
	__syncthreads();            /* the only block-level barrier */

	const uint32_t n_tiles = P.tiles_x * P.tiles_y;
	const uint32_t tile = blockIdx.x * WT_WARPS + wid;
	if (tile >= n_tiles) return;
	const uint32_t tx = tile % P.tiles_x, ty = tile / P.tiles_x;
	if (!owns_tile_row(P, ty)) return;
	WarpTile& T = S.w[wid];

	/* list length; the cursor is re-armed for the next draw */
	uint32_t n_raw = 0;
	if (lane == 0)
	{
		n_raw = P.tile_count[tile];
		if (n_raw) P.tile_count[tile] = 0u;
		if (n_raw > P.bin_cap) atomicMax(&P.ctr->max_list, n_raw);
		else if (n_raw && P.count_fragments) atomicAdd(&P.ctr->pair_total, (unsigned long long)n_raw);
	}
	n_raw = __shfl_sync(0xffffffffu, n_raw, 0);
	if (P.ctr->overflow || P.diag) return;
	const uint32_t n_list = min(n_raw, P.bin_cap);
	const ClearParams cp = P.clear;
	if (n_list == 0 && !cp.flags) return;

	const int tile_x0 = (int)(tx << SWGL_TILE_SHIFT), tile_r0 = (int)(ty << WT_H_SHIFT);
	const int band_last_y = P.ytop - tile_r0;                /* raster y of tile row 0 */
	const int band_first_y = band_last_y - (WT_H - 1);

	/* ---- stage the tile: lane -> row lane/4, 8 pixels from column 8*(lane%4) ---- */
	bool dirty = false;
	{
		uint32_t c8[8]; float d8[8];
		const int r = (int)(lane >> 2), px0 = tile_x0 + (int)((lane & 3u) << 3);
		wt_load8(P, cp, tile_r0 + r, px0, c8, d8, dirty);
		const uint32_t s = (uint32_t)r * SWGL_TILE + ((lane & 3u) << 3);
		*(uint4*)&T.color[s] = make_uint4(c8[0], c8[1], c8[2], c8[3]);
		*(uint4*)&T.color[s + 4] = make_uint4(c8[4], c8[5], c8[6], c8[7]);
		*(float4*)&T.depth[s] = make_float4(d8[0], d8[1], d8[2], d8[3]);
		*(float4*)&T.depth[s + 4] = make_float4(d8[4], d8[5], d8[6], d8[7]);
	}
	__syncwarp();

	uint32_t n_tested = 0, n_shaded = 0;
	if (n_list > 0)
	{
		/* ---- ascending primitive id = submission order ---- */
		uint32_t* gl_ids = P.pairs + (size_t)tile * P.bin_cap;
		const uint32_t* sorted = gl_ids;
		uint32_t first_batch_id = 0xffffffffu;
		if (n_list <= 32)
			first_batch_id = warp_sort32(lane < n_list ? gl_ids[lane] : 0xffffffffu, lane);
		else if (n_list <= 128)
		{
			/* up to four runs of 32, each sorted with the shuffle network; an id's final place is its
			 * position in its own run plus the ids below it in the other runs (branch-free lower
			 * bounds, all runs in flight); ranks stay in registers until every lane has finished
			 * reading the runs, then the ids are scattered to their final places */
			const uint32_t n_runs = (n_list + 31u) >> 5;
			uint32_t x[4], rank[4];
#pragma unroll
			for (uint32_t c = 0; c < 4; c++)
			{
				x[c] = 0xffffffffu;
				if (c < n_runs)
				{
					const uint32_t i = (c << 5) + lane;
					x[c] = warp_sort32(i < n_list ? gl_ids[i] : 0xffffffffu, lane);
					T.ids[i] = x[c];
				}
			}
			__syncwarp();
#pragma unroll
			for (uint32_t c = 0; c < 4; c++)
			{
				rank[c] = lane;
				if (c < n_runs && x[c] != 0xffffffffu)
				{
					uint32_t cnt[4] = { 0u, 0u, 0u, 0u };
#pragma unroll
					for (uint32_t st = 32; st > 0; st >>= 1)
#pragma unroll
						for (uint32_t r = 0; r < 4; r++)
							if (r != c && r < n_runs && cnt[r] + st <= 32u && T.ids[(r << 5) + cnt[r] + st - 1u] < x[c]) cnt[r] += st;
					rank[c] += cnt[0] + cnt[1] + cnt[2] + cnt[3];
				}
			}
			__syncwarp();
#pragma unroll
			for (uint32_t c = 0; c < 4; c++)
				if (c < n_runs && x[c] != 0xffffffffu) T.ids[rank[c]] = x[c];
			__syncwarp();
			sorted = T.ids;
		}
		else if (n_list <= WT_SORT_CAP)
		{
			for (uint32_t i = lane; i < n_list; i += 32) T.ids[i] = gl_ids[i];
			__syncwarp();
			warp_sort_mem(T.ids, n_list, lane);
			sorted = T.ids;
		}
		else warp_sort_mem(gl_ids, n_list, lane);

		for (uint32_t base = 0; base < n_list; base += 32)
		{
			const uint32_t nb = min(32u, n_list - base);
			/* ---- phase A: lane = primitive ---- */
			const uint32_t pid = (n_list <= 32) ? first_batch_id : (lane < nb ? sorted[base + lane] : 0xffffffffu);
			uint32_t cnt = 0, rowinfo = 0;
#pragma unroll
			for (int r = 0; r < WT_H; r++) T.span[r][lane] = 0;
			if (lane < nb)
			{
				const Prim* q = P.prims + pid;
				const float4 a = q->v[0], b = q->v[1], c = q->v[2];
				const uint32_t band = q->band;
				TriWalk w;
				tri_setup(a, b, c, P, w);
				const int y_in = max(w.ys, band_first_y), y_out = min(w.ye - 1, band_last_y);
				if (y_out >= y_in)
				{
					float x0, x1, s1;
					bool switched;
					walk_to_row(P, w, band, ty, y_in, x0, x1, s1, switched);
					for (int y = y_in; y <= y_out; y++)
					{
						int xa, xb;
						row_span(x0, x1, P, xa, xb);
						xa = min(max(xa, tile_x0), tile_x0 + SWGL_TILE) - tile_x0;
						xb = min(max(xb, tile_x0), tile_x0 + SWGL_TILE) - tile_x0;
						if (xb < xa) xb = xa;
						T.span[band_last_y - y][lane] = (uint16_t)(xa | (xb << 8));
						cnt += (uint32_t)(xb - xa);
						if (!switched && (float)y + 1.0f >= w.c1y) { switched = true; s1 = w.s2; x1 = w.c1x; }
						x0 += w.s0; x1 += s1;
					}
					rowinfo = (uint32_t)(band_last_y - y_out) | ((uint32_t)(y_out - y_in + 1) << 8);
				}
			}
			/* ---- S: exclusive scan of the fragment counts ---- */
			uint32_t incl = cnt;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += y; }
			const uint32_t off = incl - cnt;
			const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
			n_tested += (lane == 0) ? total : 0u;
			__syncwarp();

			/* ---- phase B: lane = fragment, dense steps of 32 in submission order ---- */
			for (uint32_t t0 = 0; t0 < total; t0 += 32)
			{
				const uint32_t t = t0 + lane;
				const bool active = t < total;
				/* owner = last lane whose offset is <= t (empty lanes share the offset of their
				 * successor, so the last one is the lane that really owns fragment t) */
				uint32_t owner = 0;
#pragma unroll
				for (uint32_t st = 16; st > 0; st >>= 1)
				{
					const uint32_t probe = min(owner + st, 31u);
					const uint32_t v = __shfl_sync(0xffffffffu, off, probe);
					if (owner + st < 32u && v <= t) owner += st;
				}
				const uint32_t o_off = __shfl_sync(0xffffffffu, off, owner);
				const uint32_t o_rows = __shfl_sync(0xffffffffu, rowinfo, owner);
				const uint32_t o_pid = __shfl_sync(0xffffffffu, pid, owner);
				bool pending = active;
				uint32_t pix = 0;
				float z = 0.0f;
				float4 col = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				bool shaded_early = false;
				if (active)
				{
					uint32_t g = t - o_off, r = o_rows & 0xffu, sp = 0;
					const uint32_t r_end = r + (o_rows >> 8);
					for (; r < r_end; r++)
					{
						sp = T.span[r][owner];
						const uint32_t len = (sp >> 8) - (sp & 0xffu);
						if (g < len) break;
						g -= len;
					}
					const uint32_t lx = (sp & 0xffu) + g;
					pix = r * SWGL_TILE + lx;
					const Prim* q = P.prims + o_pid;
					const float4 a = q->v[0], b = q->v[1], c = q->v[2];
					BaryConst k;
					bary_setup(a, b, c, k);
					FragIn fi;
					frag_weights(k, (float)(tile_x0 + (int)lx), (float)(band_last_y - (int)r), fi.u, fi.v, fi.w, z);
					/* the fragment shader does not read the framebuffer: run it before the ordered
					 * part unless the fragment already fails against the stored depth (then it can only
					 * pass later if an earlier fragment stores exactly 0.0 = "empty": shaded late) */
					const float cur = T.depth[pix];
					if (cur == 0.0f || cur >= z)
					{
						fi.vid0 = q->vid[0]; fi.vid1 = q->vid[1]; fi.vid2 = q->vid[2];
						fi.a = P.vary + (size_t)fi.vid0 * P.nvf + P.fs_slot;
						fi.b = P.vary + (size_t)fi.vid1 * P.nvf + P.fs_slot;
						fi.c = P.vary + (size_t)fi.vid2 * P.nvf + P.fs_slot;
						fi.stride = 1;
						col = clamp_color(run_fragment<FS>(P, fi));
						shaded_early = true;
					}
				}
				/* ---- ordered commit: lanes on the same pixel go in lane order ---- */
				const uint32_t peers = __match_any_sync(0xffffffffu, active ? pix : (0x80000000u | lane));
				uint32_t my_turn = (uint32_t)__popc(peers & ((1u << lane) - 1u));
				uint32_t turns = my_turn;
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) turns = max(turns, __shfl_xor_sync(0xffffffffu, turns, o));
				for (uint32_t turn = 0; turn <= turns; turn++)
				{
					if (pending && my_turn == turn)
					{
						const float cur = T.depth[pix];
						if (cur == 0.0f || cur >= z)     /* swgl.c:3387 */
						{
							T.depth[pix] = z;
							n_shaded++;
							if (!shaded_early)
							{
								const Prim* q = P.prims + o_pid;
								BaryConst k;
								bary_setup(q->v[0], q->v[1], q->v[2], k);
								FragIn fi;
								float z2;
								frag_weights(k, (float)(tile_x0 + (int)(pix & (SWGL_TILE - 1))), (float)(band_last_y - (int)(pix >> SWGL_TILE_SHIFT)), fi.u, fi.v, fi.w, z2);
								fi.vid0 = q->vid[0]; fi.vid1 = q->vid[1]; fi.vid2 = q->vid[2];
								fi.a = P.vary + (size_t)fi.vid0 * P.nvf + P.fs_slot;
								fi.b = P.vary + (size_t)fi.vid1 * P.nvf + P.fs_slot;
								fi.c = P.vary + (size_t)fi.vid2 * P.nvf + P.fs_slot;
								fi.stride = 1;
								col = clamp_color(run_fragment<FS>(P, fi));
							}
							T.color[pix] = blend_pack_lut(col.x, col.y, col.z, col.w, T.color[pix], S.lut);
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
		for (int i = 0; i < WT_H / 4; i++)
		{
			const int r = (int)(lane >> 3) + 4 * i, row = tile_r0 + r;
			if (row >= (int)P.H) continue;
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
