/*
 * swgl_raster_frag.cuh -- per-tile rasteriser, fragment-parallel form (included by swgl_dev.cu).
 *
 * The pixel-owner kernel (k_raster) spends a whole warp on every primitive that touches its
 * four rows; with ~6-pixel triangles that leaves ~8 % of the lanes busy.  Here the tile's
 * colour and depth are staged in shared memory and the work is re-shaped in three phases per
 * batch of primitives:
 *
 *   A  one thread per primitive: replay the span walk over the tile's rows (swgl.c:3356-3361,
 *      3466-3471) and record [xa, xb) per row plus a running fragment count;
 *   S  block-wide exclusive scan of the fragment counts;
 *   B  one thread per FRAGMENT (dense, chunks of 256): binary-search the owning primitive and
 *      row, evaluate Barycentric + perspective correction + z (swgl.c:3365-3382), then commit.
 *
 * Commit keeps the reference's per-pixel submission order: fragments are numbered in
 * (primitive, row, x) order, so among the fragments of one chunk that hit the same pixel the
 * lowest thread index is the earliest primitive.  Each round every pending fragment does an
 * atomicMin of its thread index on the pixel's slot; the winner runs the depth test, shades,
 * blends into shared memory and releases the slot; losers retry.  Small meshes need two
 * rounds (the reference double-covers shared diagonals), large triangles one.
 *
 * The finished tile leaves with 128-bit coalesced stores, to local HBM and, when a peer
 * colour target is set, straight into rank 0's framebuffer over NVLink.
 */
#ifndef SWGL_RASTER_FRAG_CUH
#define SWGL_RASTER_FRAG_CUH

#define FRAG_THREADS   256
#define FRAG_BATCH     256      /* primitives per batch */
#define FRAG_SPAN_POOL 3072     /* (primitive, row) spans per batch */
#define FRAG_SORT_CAP  2048     /* lists up to this length are sorted in shared memory */
#define FRAG_MAX_FRAGS 16384    /* fragments per batch (start-bit map size) */

struct FragShared
{
	uint32_t color[SWGL_TILE * SWGL_TILE];
	float    depth[SWGL_TILE * SWGL_TILE];
	uint32_t owner[SWGL_TILE * SWGL_TILE];     /* commit arbitration slot per pixel */
	float    lut[256];                         /* byte / 255.0f (swgl.c:3434-3437) */
	uint32_t ids[FRAG_BATCH];                  /* this batch, ascending primitive id */
	uint32_t frag_off[FRAG_BATCH + 1];         /* exclusive scan of fragment counts */
	uint32_t span_base[FRAG_BATCH + 1];        /* exclusive scan of row counts */
	uint32_t row0[FRAG_BATCH];                 /* first tile row of the primitive's spans */
	/* fragment -> primitive map: bit f is set where a primitive's fragments start; own0[w] is the
	 * rank (among primitives with fragments) of the owner of fragment 32*w; nonempty[rank] = slot */
	uint32_t start_bits[FRAG_MAX_FRAGS / 32];
	uint16_t own0[FRAG_MAX_FRAGS / 32];
	uint16_t nonempty[FRAG_BATCH];
	union
	{
		struct
		{
			uint16_t span[FRAG_SPAN_POOL];     /* xa | xb << 8, tile-local columns */
			uint16_t row_pre[FRAG_SPAN_POOL];  /* fragments of the primitive before this row */
		} s;
		uint32_t sort_buf[FRAG_SORT_CAP];      /* whole-list sort, before the first batch */
	} u;
	uint32_t scan_tmp[FRAG_THREADS / 32];
	uint32_t scan_total;
	uint32_t batch_count;
	unsigned long long red[2][FRAG_THREADS / 32];
};

/* exclusive block scan of one value per thread; returns the exclusive prefix, total in *total */
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t* warp_tmp, uint32_t* total_out)
{
	const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	uint32_t x = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((int)lane >= o) x += y; }
	if (lane == 31) warp_tmp[wid] = x;
	__syncthreads();
	if (wid == 0)
	{
		uint32_t s = lane < (FRAG_THREADS / 32) ? warp_tmp[lane] : 0u;
#pragma unroll
		for (int o = 1; o < FRAG_THREADS / 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if ((int)lane >= o) s += y; }
		if (lane < (FRAG_THREADS / 32)) warp_tmp[lane] = s;
		if (lane == (FRAG_THREADS / 32) - 1) *total_out = s;
	}
	__syncthreads();
	const uint32_t before = wid ? warp_tmp[wid - 1] : 0u;
	__syncthreads();            /* the next scan rewrites warp_tmp: every warp must have read its prefix first */
	return before + (x - v);
}

template <int FS>
__global__ void __launch_bounds__(FRAG_THREADS, 6) k_raster_frag(const __grid_constant__ DrawParams P)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	FragShared& S = *reinterpret_cast<FragShared*>(smem_raw);

	const uint32_t tx = blockIdx.x, ty = blockIdx.y;
	if (!owns_tile_row(P, ty)) return;
	const uint32_t tile = ty * P.tiles_x + tx;
	const uint32_t tid = threadIdx.x;
	const uint32_t n_list = take_tile_list(P, tile);
	if (P.ctr->overflow || P.diag) return;
	const size_t list_off = (size_t)tile * P.bin_cap;
	const ClearParams cp = P.clear;
	if (n_list == 0 && !cp.flags) return;

	const int tile_x0 = (int)(tx << SWGL_TILE_SHIFT), tile_r0 = (int)(ty << SWGL_TILE_SHIFT);

	/* ---- stage the tile: 4x1 strip per thread, 4 passes of 8 rows ---- */
	bool tile_dirty = false;
	{
		const uint32_t q = tid & 7u;
		for (uint32_t r = tid >> 3; r < SWGL_TILE; r += FRAG_THREADS / 8)
		{
			const int row = tile_r0 + (int)r, px0 = tile_x0 + (int)(q << 2);
			uint32_t c4[4] = { 0u, 0u, 0u, 0u }; float d4[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
			if (row < (int)P.H)
			{
				const size_t pix = (size_t)row * P.W + (size_t)px0;
				const bool in_row = cp.flags && row >= cp.y0 && row < cp.y1;
				bool all_c = true, all_d = true;
				for (int k = 0; k < 4; k++)
				{
					const bool inside = in_row && (px0 + k) >= cp.x0 && (px0 + k) < cp.x1;
					tile_dirty |= inside;
					all_c &= inside && (cp.flags & 1u);
					all_d &= inside && (cp.flags & 2u);
				}
				if (px0 + 3 < (int)P.W && ((pix & 3u) == 0))
				{
					if (!all_c) { uint4 v = *(const uint4*)(P.color + pix); c4[0] = v.x; c4[1] = v.y; c4[2] = v.z; c4[3] = v.w; }
					if (!all_d) { float4 v = *(const float4*)(P.depth + pix); d4[0] = v.x; d4[1] = v.y; d4[2] = v.z; d4[3] = v.w; }
				}
				else
					for (int k = 0; k < 4; k++)
						if (px0 + k < (int)P.W) { c4[k] = P.color[pix + k]; d4[k] = P.depth[pix + k]; }
				for (int k = 0; k < 4; k++)
				{
					const bool inside = in_row && (px0 + k) >= cp.x0 && (px0 + k) < cp.x1;
					if (inside && (cp.flags & 1u)) c4[k] = cp.word;
					if (inside && (cp.flags & 2u)) d4[k] = 0.0f;
				}
			}
			const uint32_t s = r * SWGL_TILE + (q << 2);
			*(uint4*)&S.color[s] = make_uint4(c4[0], c4[1], c4[2], c4[3]);
			*(float4*)&S.depth[s] = make_float4(d4[0], d4[1], d4[2], d4[3]);
			*(uint4*)&S.owner[s] = make_uint4(0u, 0u, 0u, 0u);
		}
		S.lut[tid] = (float)tid / 255.0f;
	}

	uint32_t n_tested = 0, n_shaded = 0, commit_round = 0;

	if (n_list > 0)
	{
		/* ---- ascending primitive id = submission order ---- */
		uint32_t* gl_ids = P.pairs + list_off;
		if (n_list <= FRAG_BATCH)
		{
			/* every warp sorts 32 ids with a shuffle network, then each id finds its final place as
			 * (its position in its own run) + (ids below it in the other runs, by binary search) */
			const uint32_t lane = tid & 31u, wid = tid >> 5;
			const uint32_t n_runs = (n_list + 31u) >> 5;
			uint32_t x = tid < n_list ? gl_ids[tid] : 0xffffffffu;
			if (wid < n_runs)
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
				S.u.sort_buf[tid] = x;
			}
			__syncthreads();
			if (x != 0xffffffffu)
			{
				/* lower bounds in all other runs at once (6 independent shared loads per step) */
				uint32_t cnt[FRAG_BATCH / 32];
#pragma unroll
				for (uint32_t r = 0; r < FRAG_BATCH / 32; r++) cnt[r] = 0;
#pragma unroll
				for (uint32_t st = 32; st > 0; st >>= 1)
#pragma unroll
					for (uint32_t r = 0; r < FRAG_BATCH / 32; r++)
						if (r < n_runs && r != wid && cnt[r] + st <= 32u && S.u.sort_buf[(r << 5) + cnt[r] + st - 1u] < x) cnt[r] += st;
				uint32_t rank = lane;
#pragma unroll
				for (uint32_t r = 0; r < FRAG_BATCH / 32; r++) rank += cnt[r];
				S.ids[rank] = x;
				gl_ids[rank] = x;        /* a batch cut short by a pool limit reloads from here */
			}
		}
		else if (n_list <= FRAG_SORT_CAP)
		{
			for (uint32_t i = tid; i < n_list; i += FRAG_THREADS) S.u.sort_buf[i] = gl_ids[i];
			uint32_t n_pow2 = 1; while (n_pow2 < n_list) n_pow2 <<= 1;
			sort_ids_shared(S.u.sort_buf, n_list, n_pow2);
			for (uint32_t i = tid; i < n_list; i += FRAG_THREADS) gl_ids[i] = S.u.sort_buf[i];
		}
		else sort_ids_global(gl_ids, n_list);
		__syncthreads();

		const int band_last_y = P.ytop - tile_r0;                 /* raster y of tile row 0 */
		const int band_first_y = band_last_y - (SWGL_TILE - 1);

		uint32_t base = 0;
		while (base < n_list)
		{
			uint32_t nb = min((uint32_t)FRAG_BATCH, n_list - base);
			if (n_list > FRAG_BATCH || base > 0)
			{
				if (tid < nb) S.ids[tid] = gl_ids[base + tid];
				__syncthreads();
			}

			/* ---- phase A1: rows each primitive has inside this tile ---- */
			Prim pr;
			TriWalk w;
			int y_in = 0, y_out = -1;
			if (tid < nb)
			{
				const uint32_t pid = S.ids[tid];
				pr = load_prim(P, pid);
				tri_setup(pr.v[0], pr.v[1], pr.v[2], P, w);
				y_in = max(w.ys, band_first_y);
				y_out = min(w.ye - 1, band_last_y);
			}
			const uint32_t my_rows = (tid < nb && y_out >= y_in) ? (uint32_t)(y_out - y_in + 1) : 0u;
			uint32_t rows_total;
			const uint32_t my_span_base = block_scan_excl(my_rows, S.scan_tmp, &S.scan_total);
			rows_total = S.scan_total;
			if (rows_total > FRAG_SPAN_POOL)
			{
				/* too many spans for one batch: keep the longest prefix that fits (always >= 1
				 * primitive, a primitive has at most 32 rows here) */
				if (tid == 0) S.batch_count = nb;
				__syncthreads();
				if (tid < nb && my_span_base + my_rows > FRAG_SPAN_POOL) atomicMin(&S.batch_count, tid);
				__syncthreads();
				nb = S.batch_count;
			}

			/* ---- phase A2: replay the walk, record spans and fragment counts ---- */
			uint32_t my_frags = 0;
			if (tid < nb)
			{
				float x0, x1, s1;
				bool switched;
				walk_to_row(P, w, pr.band, ty, y_in, x0, x1, s1, switched);
				S.row0[tid] = (uint32_t)(band_last_y - y_out);       /* smallest tile row index */
				for (int y = y_in; y <= y_out; y++)
				{
					int xa, xb;
					row_span(x0, x1, P, xa, xb);
					xa = min(max(xa, tile_x0), tile_x0 + SWGL_TILE) - tile_x0;   /* both ends inside [0, 32]: they share a 16-bit word */
					xb = min(max(xb, tile_x0), tile_x0 + SWGL_TILE) - tile_x0;   /* xb may be INT_MIN: clamp before subtracting */
					if (xb < xa) xb = xa;
					/* spans are stored by ascending tile row = descending y */
					const uint32_t slot = my_span_base + (uint32_t)(y_out - y);
					S.u.s.span[slot] = (uint16_t)(xa | (xb << 8));
					my_frags += (uint32_t)(xb - xa);
					if (!switched && (float)y + 1.0f >= w.c1y) { switched = true; s1 = w.s2; x1 = w.c1x; }
					x0 += w.s0; x1 += s1;
				}
				S.span_base[tid] = my_span_base;
			}
			/* one scan carries both the fragment offsets (low 20 bits) and the rank among the
			 * primitives that have fragments (high bits) */
			S.start_bits[tid] = 0u; S.start_bits[tid + FRAG_THREADS] = 0u;
			uint32_t frag_total;
			const uint32_t packed = block_scan_excl(my_frags | (my_frags ? (1u << 20) : 0u), S.scan_tmp, &S.scan_total);
			uint32_t my_frag_off = packed & 0xfffffu;
			const uint32_t my_rank = packed >> 20;
			frag_total = S.scan_total & 0xfffffu;
			if (frag_total > FRAG_MAX_FRAGS)
			{
				/* keep the longest prefix whose fragments fit the start-bit map (>= 1 primitive) */
				if (tid == 0) S.batch_count = nb;
				__syncthreads();
				if (tid < nb && my_frag_off + my_frags > FRAG_MAX_FRAGS) atomicMin(&S.batch_count, tid);
				__syncthreads();
				nb = S.batch_count;
			}
			if (tid < nb)
			{
				S.frag_off[tid] = my_frag_off;
				/* per-row prefix inside the primitive, in the stored (ascending tile row) order */
				uint32_t acc = 0;
				for (uint32_t k = 0; k < my_rows; k++)
				{
					const uint32_t sp = S.u.s.span[my_span_base + k];
					S.u.s.row_pre[my_span_base + k] = (uint16_t)acc;
					acc += (sp >> 8) - (sp & 0xffu);
				}
				if (my_frags)
				{
					atomicOr(&S.start_bits[my_frag_off >> 5], 1u << (my_frag_off & 31u));
					const uint32_t w_last = (my_frag_off + my_frags - 1u) >> 5;
					for (uint32_t wd = (my_frag_off + 31u) >> 5; wd <= w_last; wd++) S.own0[wd] = (uint16_t)my_rank;
					S.nonempty[my_rank] = (uint16_t)tid;
				}
			}
			if (tid == nb) { S.frag_off[nb] = my_frag_off; S.span_base[nb] = my_span_base; }
			if (nb == FRAG_THREADS && tid == 0) { S.frag_off[nb] = frag_total; S.span_base[nb] = rows_total; }
			__syncthreads();
			frag_total = S.frag_off[nb];

			/* ---- phase B: one thread per fragment, chunks of 256 in submission order ---- */
			for (uint32_t fbase = 0; fbase < frag_total; fbase += FRAG_THREADS)
			{
				const uint32_t f = fbase + tid;
				bool pending = f < frag_total;
				uint32_t pix = 0;
				float z = 0.0f;
				float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				bool shaded_early = false;
				uint32_t pid = 0;
				if (pending)
				{
					/* owning primitive from the start-bit map */
					const uint32_t wd = f >> 5, bt = f & 31u;
					const uint32_t rank = (uint32_t)S.own0[wd] + __popc(S.start_bits[wd] & 0xfffffffeu & (0xffffffffu >> (31u - bt)));
					const uint32_t lo = S.nonempty[rank];
					const uint32_t g = f - S.frag_off[lo];
					/* row: largest k with row_pre[k] <= g among rows that are not empty past g */
					const uint32_t sb = S.span_base[lo], nr = S.span_base[lo + 1] - sb;
					uint32_t rl = 0, rh = nr;
					while (rh - rl > 1) { const uint32_t mid = (rl + rh) >> 1; if (S.u.s.row_pre[sb + mid] <= g) rl = mid; else rh = mid; }
					const uint32_t sp = S.u.s.span[sb + rl];
					const uint32_t lx = (sp & 0xffu) + (g - S.u.s.row_pre[sb + rl]);
					const uint32_t r = S.row0[lo] + rl;
					pix = r * SWGL_TILE + lx;
					pid = S.ids[lo];
					const Prim qv = load_prim(P, pid);
					const Prim* q = &qv;
					const float4 a = q->v[0], b = q->v[1], c = q->v[2];
					BaryConst k;
					bary_setup(a, b, c, k);
					FragIn fi;
					frag_weights(k, (float)(tile_x0 + (int)lx), (float)(band_last_y - (int)r), fi.u, fi.v, fi.w, z);
					n_tested++;
					/* The fragment shader does not read the framebuffer, so it runs here, before the
					 * ordered part (depth test + blend).  It is skipped when the fragment fails against
					 * the depth stored right now -- it then almost always fails in submission order too
					 * (a stored depth only decreases); the exception (an earlier fragment of this chunk
					 * storing exactly 0.0 = "empty") is caught in the commit loop, which shades late. */
					const float cur = S.depth[pix];
					if (cur == 0.0f || cur >= z)
					{
						fi.vid0 = q->vid[0]; fi.vid1 = q->vid[1]; fi.vid2 = q->vid[2];
						fi.a = P.vary + (size_t)fi.vid0 * P.nvf + P.fs_slot;
						fi.b = P.vary + (size_t)fi.vid1 * P.nvf + P.fs_slot;
						fi.c = P.vary + (size_t)fi.vid2 * P.nvf + P.fs_slot;
						fi.stride = 1;
						fi.lod = (FS == SWFS_GENERIC && P.mip_lod) ? mip_level(a.x, a.y, b.x, b.y, c.x, c.y) : 0.0f;
						o = clamp_color(run_fragment<FS>(P, fi));
						shaded_early = true;
					}
				}
				/* ordered commit */
				while (__syncthreads_or(pending ? 1 : 0))
				{
					/* the lowest thread index of this round wins the pixel: keys carry the round number in
					 * their high bits, so a slot never has to be released (a stale key always loses) */
					commit_round++;
					const uint32_t key = (commit_round << 9) | (511u - tid);
					if (pending) atomicMax(&S.owner[pix], key);
					__syncthreads();
					if (pending && S.owner[pix] == key)
					{
						const float cur = S.depth[pix];
						if (cur == 0.0f || cur >= z)     /* swgl.c:3387 */
						{
							S.depth[pix] = z;
							n_shaded++;
							if (!shaded_early)
							{
								/* rare: recompute the weights (same arithmetic, same bits) and shade now */
								const Prim qv2 = load_prim(P, pid);
								const Prim* q_prim = &qv2;
								BaryConst k;
								bary_setup(q_prim->v[0], q_prim->v[1], q_prim->v[2], k);
								FragIn fi;
								float z2;
								frag_weights(k, (float)(tile_x0 + (int)(pix & (SWGL_TILE - 1))), (float)(band_last_y - (int)(pix >> SWGL_TILE_SHIFT)), fi.u, fi.v, fi.w, z2);
								fi.vid0 = q_prim->vid[0]; fi.vid1 = q_prim->vid[1]; fi.vid2 = q_prim->vid[2];
								fi.a = P.vary + (size_t)fi.vid0 * P.nvf + P.fs_slot;
								fi.b = P.vary + (size_t)fi.vid1 * P.nvf + P.fs_slot;
								fi.c = P.vary + (size_t)fi.vid2 * P.nvf + P.fs_slot;
								fi.stride = 1;
								fi.lod = (FS == SWFS_GENERIC && P.mip_lod) ? mip_level(q_prim->v[0].x, q_prim->v[0].y, q_prim->v[1].x, q_prim->v[1].y, q_prim->v[2].x, q_prim->v[2].y) : 0.0f;
								o = clamp_color(run_fragment<FS>(P, fi));
							}
							S.color[pix] = blend_pack_lut(o.x, o.y, o.z, o.w, S.color[pix], S.lut);
							tile_dirty = true;
						}
						pending = false;
					}
				}
			}
			base += nb;
			__syncthreads();
		}
	}

	/* ---- write-back: 128-bit stores of the finished tile ---- */
	const bool dirty_any = __syncthreads_or(tile_dirty ? 1 : 0);
	if (dirty_any)
	{
		const uint32_t q = tid & 7u;
		for (uint32_t r = tid >> 3; r < SWGL_TILE; r += FRAG_THREADS / 8)
		{
			const int row = tile_r0 + (int)r, px0 = tile_x0 + (int)(q << 2);
			if (row >= (int)P.H) continue;
			const size_t pix = (size_t)row * P.W + (size_t)px0;
			const uint32_t s = r * SWGL_TILE + (q << 2);
			uint4 c4 = *(const uint4*)&S.color[s];
			float4 d4 = *(const float4*)&S.depth[s];
			d4.x = canon_nan(d4.x); d4.y = canon_nan(d4.y); d4.z = canon_nan(d4.z); d4.w = canon_nan(d4.w);
			if (px0 + 3 < (int)P.W && ((pix & 3u) == 0))
			{
				*(uint4*)(P.color + pix) = c4;
				*(float4*)(P.depth + pix) = d4;
				if (P.peer_color) *(uint4*)(P.peer_color + pix) = c4;
			}
			else
			{
				const uint32_t cc[4] = { c4.x, c4.y, c4.z, c4.w };
				const float dd[4] = { d4.x, d4.y, d4.z, d4.w };
				for (int k = 0; k < 4; k++)
					if (px0 + k < (int)P.W)
					{
						P.color[pix + k] = cc[k]; P.depth[pix + k] = dd[k];
						if (P.peer_color) P.peer_color[pix + k] = cc[k];
					}
			}
		}
	}

	if (P.count_fragments)
	{
		unsigned long long a = n_tested, b = n_shaded;
		for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
		if ((tid & 31u) == 0) { S.red[0][tid >> 5] = a; S.red[1][tid >> 5] = b; }
		__syncthreads();
		if (tid == 0)
		{
			unsigned long long ta = 0, tb = 0;
			for (int i = 0; i < FRAG_THREADS / 32; i++) { ta += S.red[0][i]; tb += S.red[1][i]; }
			if (ta) atomicAdd(&P.ctr->tested[tile % SWGL_CTR_SLOTS], ta);
			if (tb) atomicAdd(&P.ctr->shaded[tile % SWGL_CTR_SLOTS], tb);
		}
	}
}

#endif
