/*
 * swgl_dev.cu -- CUDA device layer of libswgl_b200.so (sm_100a), behind the C ABI of
 * include/swgl_dev.h.  It replaces the reference's per-draw work:
 *
 *   k_clear          glClear when it cannot be fused into the next draw      swgl.c:3183-3214
 *   k_vertex         attribute fetch + VS + varying capture + divide/viewport snap, one thread
 *                    per vertex                          swgl.c:3618-3666, 3683-3692
 *   k_setup_bin      near clip, triangle set-up and single-pass binning into fixed-capacity
 *                    per-tile lists, one thread per input triangle; short narrow primitives are
 *                    binned by vertex extent, tall/wide ones get per-tile-row walk states
 *                                                        swgl.c:499-697, 3316-3361, 3466-3471
 *   k_bin_tall       spans and list insertion of tall/wide primitives, a warp per tile row
 *                    (big-triangle draws only)            swgl.c:3356-3361
 *   k_raster_warp    (swgl_raster_warp.cuh) per tile: sort the list by primitive id, replay the
 *                    span walk, then per fragment Barycentric, perspective correction, depth
 *                    test, varying interpolation, fragment shader, blend, pack; fused clear;
 *                    128-bit write-back                  swgl.c:3358-3462
 *   k_raster_frag, k_raster   two older CTA-per-32x32-tile rasterisers, kept as independent
 *                    cross-checks (raster_path option)
 *   k_points_claim / k_points_write   GL_POINTS            swgl.c:3496-3608
 *
 * Order dependence: the reference's depth test (LEQUAL with 0.0f = empty) and its
 * unconditional blend make the result depend on submission order, so every pixel sees its
 * fragments in ascending primitive id -- lists are sorted, and same-pixel fragments of a step
 * commit in lane order.
 *
 * No tensor cores: the path is scan/scatter shaped, not a contraction.  Compiled with
 * -fmad=false (see swgl_dev_math.cuh).  There is no CPU fallback anywhere in this file.
 */
#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "swgl_dev_common.cuh"
#include "swgl_jit.h"

#define SWGL_MAX_GROUP 16
#ifndef SWGL_BIN_TALL_CTAS_PER_SM
#define SWGL_BIN_TALL_CTAS_PER_SM 8   /* a warp handles its band entries one after the other, three dependent loads each: many warps, few entries per warp */
#endif
#define SWGL_MAX_HOT_TILES 64       /* tiles whose list may exceed K and continue in the overflow pool */

/* ========================================================================================
 * device group workers
 * ====================================================================================== */
/* A device group is driven by the application's one thread, but a frame is a few dozen CUDA API calls PER
 * DEVICE (uploads, event plumbing, three launches, a counter snapshot) and those calls cost microseconds
 * each: issued one device after the other they -- not the GPUs -- set the frame time at eight devices.
 * Member i > 0 therefore has a worker thread that issues member i's calls; run(f) executes f(i) for every
 * member (f(0) on the caller) and returns when all are done.  Workers spin briefly between the fan-outs of
 * a frame and sleep on a condition variable otherwise. */
struct GroupPool
{
	int n = 0;
	std::vector<std::thread> threads;
	std::mutex mu;
	std::condition_variable cv;
	std::atomic<uint64_t> epoch{0};
	std::atomic<int> pending{0};
	const std::function<void(int)>* fn = nullptr;
	std::atomic<bool> quit{false};

	void start(int count, const int* devices)
	{
		n = count;
		for (int i = 1; i < count; i++) threads.emplace_back([this, i, d = devices[i]] { worker(i, d); });
	}
	void worker(int i, int device)
	{
		cudaSetDevice(device);
		uint64_t seen = 0;
		for (;;)
		{
			int spins = 0;
			while (epoch.load(std::memory_order_acquire) == seen && !quit.load(std::memory_order_acquire))
			{
				if (++spins < 20000) { __builtin_ia32_pause(); continue; }
				std::unique_lock<std::mutex> lk(mu);
				cv.wait_for(lk, std::chrono::milliseconds(50), [&] { return epoch.load(std::memory_order_acquire) != seen || quit.load(std::memory_order_acquire); });
			}
			if (quit.load(std::memory_order_acquire)) return;
			seen = epoch.load(std::memory_order_acquire);
			(*fn)(i);
			pending.fetch_sub(1, std::memory_order_acq_rel);
		}
	}
	void run(const std::function<void(int)>& f)
	{
		fn = &f;
		pending.store(n - 1, std::memory_order_release);
		{
			std::lock_guard<std::mutex> lk(mu);
			epoch.fetch_add(1, std::memory_order_acq_rel);
		}
		cv.notify_all();
		f(0);
		while (pending.load(std::memory_order_acquire) > 0) __builtin_ia32_pause();
	}
	void stop()
	{
		quit.store(true, std::memory_order_release);
		{ std::lock_guard<std::mutex> lk(mu); }
		cv.notify_all();
		for (auto& t : threads) t.join();
		threads.clear();
	}
};

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
	/* Write-through host mirror: when the application reads every frame back, the raster kernels
	 * store finished tiles into h_color as well (posted PCIe writes under the kernel) and
	 * glGetFramePtr only has to wait.  mirror_synced: h_color equals the device colour once the
	 * stream has drained. */
	int mirror_synced, wt_predict, color_exposed;
	uint64_t wt_draws;
	/* frame pipelining: two mirrors, h_color is the active one; swgldev_frame_submit
	 * stamps the frame and flips, swgldev_frame_wait hands the finished mirror out */
	uint32_t* h_mirror[2];
	cudaEvent_t frame_ev[2];
	cudaStream_t copy;                   /* frame copies of swgldev_frame_submit (DMA, overlaps the next upload) */
	cudaEvent_t frame_done;
	int copy_inflight;                   /* the copy stream may still read the colour attachment (slot + 1) */
	uint64_t frame_serial;
	uint32_t* rgba_staging;
	/* uploads run on their own stream and wait only for the draws that read the destination, so the
	 * next frame's buffers can cross PCIe while this frame is rasterised (every draw gets a serial and
	 * an event in a small ring) */
	cudaStream_t upload;
	cudaEvent_t draw_ev[8];
	uint64_t draw_serial;
	uint32_t* d_maxidx;
	uint32_t* h_maxidx;                  /* pinned: a pageable destination sends the 4-byte read-back through the driver's slow staging path */
	float* lut255;
	uint32_t draws_since_map;
	uint32_t* peer_color;
	/* Shared frame mirror (sort-first, one process per GPU): page-locked host memory that every rank has
	 * mapped; each rank's raster kernels store its finished tiles there over its own PCIe link, so the
	 * assembled frame is in host memory when the ranks have finished -- no gather, no copy on rank 0. */
	uint32_t* shared_mirror; uint32_t* shared_mirror_dev; size_t shared_mirror_bytes;
	uint32_t rank, n_ranks, band_rows;

	/* grow-only scratch */
	float4* clip; size_t cap_clip;
	float2* clip_xy; size_t cap_clip_xy;
	float* vary; size_t cap_vary;        /* floats */
	Prim* prims; size_t cap_prims;
	BandEntry* bands; size_t cap_bands;
	uint32_t* pairs; size_t cap_pairs;   /* tiles * bin_cap entries */
	uint2* ov_pool; size_t cap_ov;       /* overflow pool of the deep tiles: (tile, entry) */
	uint32_t* hot_store; size_t cap_hot; /* where the raster warps of deep tiles assemble their lists */
	uint32_t bin_cap;                    /* K */
	uint32_t* tile_count;
	uint32_t* winner;                    /* GL_POINTS arbitration, W*H words, kept zeroed between draws */
	Counters* ctr; Counters* h_ctr;      /* device counters, pinned snapshot */
	cudaStream_t side;                   /* counter snapshots travel here, off the critical path */
	cudaEvent_t setup_event;             /* set-up kernel of the last draw finished */
	cudaEvent_t app_event;               /* "whatever is queued on the stream now" (swgldev_max_index_after_stream) */
	cudaEvent_t ctr_event;
	int ctr_pending;                     /* a snapshot copy is in flight for last_draw */

	std::map<uint64_t, swgl_ir_op*> code_cache;
	std::vector<void*> allocations;
	std::map<uintptr_t, uint64_t> last_use;   /* allocation base -> serial of the last draw reading it */

	ClearParams pending_clear;
	DrawParams last_draw;                /* for re-issue after a scratch overflow */
	int last_draw_valid;
	int last_raster_path;

	/* options */
	int opt_fuse_clear, opt_count_fragments, opt_raster_path, opt_stage_timing, opt_diag, opt_host_mirror, opt_lean_prims;
	int opt_setup_big;                   /* 1 (default): draws of big triangles are set up by a warp per triangle (k_setup_big) */
	int opt_setup_pipelined;             /* 1: the software-pipelined set-up kernel (measured slower everywhere, kept as an option) */
	int opt_overflow_pool;               /* 1 (default): lists longer than K continue in the overflow pool; 0: K grows for every tile */
	int opt_tile_rows;                   /* warp rasteriser: 8, 4 or 2 rows per tile; 0 = chosen per draw (th_shift_of) */
	uint32_t cur_th_shift;               /* of the draw being issued (list capacity is per tile of that size) */
	int opt_jit;                         /* compile generic shaders to kernels at run time (default 1; 0: on-device interpreter) */
	int jit_failed;                      /* a compilation failed: reported once, the interpreter draws */
	int last_vs_kind, last_fs_kind;      /* of the last triangle draw, as launched */
	int opt_mip_lod;                     /* sample mip chains with the defined per-triangle LOD (default 0: bug-compatible, base level) */
	size_t opt_bin_limit;
	uint64_t n_launches;                 /* kernels launched since creation */
	int64_t selftest_mismatches;
	cudaEvent_t stage_ev[8];
	double stage_us[8];                  /* accumulated per-stage device time (stage timing mode) */
	uint64_t stage_draws;

	swgldev_stats stats;
	uint64_t n_draws;
	float* d_last_level;                 /* mip_lod: the reference's global MipMapLevel (what the last triangle left; GL_POINTS samples with it) */
	int in_fold;                         /* the draw being issued renders into the virtual framebuffer of draw_folded() */
	int64_t fold_shift;                  /* virtual storage row = real storage row - fold_shift (<= 0) */
	uint64_t draws_folded;
	uint64_t draws_refused;              /* draws skipped because the viewport leaves the framebuffer rows (error string set) */
	char error[512];

	/* Device group (swgldev_create_group): one host thread drives several GPUs.  The leader (member 0, the
	 * handle the host layer holds) carries the member list and the table that maps an allocation of member 0
	 * to its replicas on the other devices; every entry point called on the leader fans out to the members
	 * (sort-first bands, geometry replicated, every member stores its finished tiles into the leader's pinned
	 * frame mirror over its own PCIe link).  `solo` > 0: the call is the fan-out itself. */
	swgldev_ctx* group[SWGL_MAX_GROUP];
	int n_group, solo;
	GroupPool* pool;                     /* leader: one worker thread per further member */
	int group_peer_ok;                   /* every member can load from every other member's memory */
	std::map<uintptr_t, std::vector<swgldev_ptr>> replicas;
	int mirror_borrowed;                 /* shared_mirror is the leader's pinned mirror: nothing to unregister */
	cudaEvent_t slice_ev, gather_ev;     /* group uploads: this member's slice has arrived / its replica is complete */
	cudaStream_t gather;                 /* the device-to-device part runs here, so the next upload's host copy does not queue behind it */
	/* gathers in flight on `gather`, newest last used: (event, the leader's allocation it completes).  An upload
	 * into an allocation waits for the earlier gathers of the same allocation on every member; an entry that is
	 * recycled while it may still be pending makes later checks wait for the newest event (same stream: it covers all) */
	cudaEvent_t gather_ring_ev[4];
	swgldev_ptr gather_ring_dst[4];
	int gather_ring_pos, gather_pending, gather_evicted;
};

#define IS_GROUP(c) ((c)->n_group > 1 && !(c)->solo)
struct Solo { swgldev_ctx* m; explicit Solo(swgldev_ctx* m_) : m(m_) { m->solo++; } ~Solo() { m->solo--; } };

/* f(i, member i) for every member of the leader's group, each on the thread that drives that member; the
 * member is marked solo for the duration (its entry points then act on it alone).  Returns the OR of the results. */
static int group_each(swgldev_ctx* c, const std::function<int(int, swgldev_ctx*)>& f)
{
	std::atomic<int> rc{0};
	const std::function<void(int)> body = [&](int i) { swgldev_ctx* m = c->group[i]; Solo s(m); cudaSetDevice(m->device); if (f(i, m)) rc.fetch_or(1); };
	if (c->pool) c->pool->run(body);
	else for (int i = 0; i < c->n_group; i++) body(i);
	cudaSetDevice(c->device);
	return rc.load() ? -1 : 0;
}

static void set_err(swgldev_ctx* c, const char* what, cudaError_t e)
{
	if (c->error[0]) return; /* keep the first */
	if (e != cudaSuccess) snprintf(c->error, sizeof(c->error), "%.300s: %.200s", what, cudaGetErrorString(e));
	else snprintf(c->error, sizeof(c->error), "%.500s", what);
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
/* Sort-first ranks clear the rows of the bands they own only (the other rows of the local attachment are
 * nobody's, and in the assembled frame they belong to ranks that may be storing into them right now); `mirror`
 * is the assembled colour target of those rows (a peer's attachment or a shared host mirror), if any. */
__global__ void k_clear(uint32_t* __restrict__ color, float* __restrict__ depth, uint32_t* __restrict__ mirror, uint32_t W,
                        ClearParams cp, uint32_t rank, uint32_t n_ranks, uint32_t band_rows)
{
	int y = cp.y0 + (int)blockIdx.y;
	int x = cp.x0 + (int)(blockIdx.x * blockDim.x + threadIdx.x) * 4;
	if (y >= cp.y1) return;
	if (n_ranks > 1 && (((uint32_t)y >> 5) / band_rows) % n_ranks != rank) return;
	size_t base = (size_t)y * W;
	if (x + 3 < cp.x1 && (((base + (size_t)x) & 3u) == 0))
	{
		const uint4 w4 = make_uint4(cp.word, cp.word, cp.word, cp.word);
		if (cp.flags & 1u) { *(uint4*)(color + base + x) = w4; if (mirror) *(uint4*)(mirror + base + x) = w4; }
		if (cp.flags & 2u) *(float4*)(depth + base + x) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		return;
	}
	for (int k = 0; k < 4 && x + k < cp.x1; k++)
	{
		if (cp.flags & 1u) { color[base + x + k] = cp.word; if (mirror) mirror[base + x + k] = cp.word; }
		if (cp.flags & 2u) depth[base + x + k] = 0.0f;
	}
}

__global__ void k_init_lut255(float* lut) { lut[threadIdx.x] = (float)threadIdx.x / 255.0f; }

__global__ void k_fill_fb(uint32_t* __restrict__ color, float* __restrict__ depth, size_t n, uint32_t word, float d)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	size_t stride = (size_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride) { color[i] = word; depth[i] = d; }
}

/* ---- device group: the all-gather of a sharded upload.  Every member holds its own slice of the data in its
 * replica; this kernel, run by member `self`, pulls the other members' slices out of THEIR replicas (peer
 * loads over NVLink) into its own.  Slices are multiples of 256 bytes, replicas 16-byte aligned. ---- */
struct GatherArgs
{
	const char* src[SWGL_MAX_GROUP];
	char* dst;
	unsigned long long slice, bytes;
	int n, self;
};

__global__ void __launch_bounds__(256) k_group_gather(const __grid_constant__ GatherArgs a)
{
	const unsigned long long words = a.bytes >> 4, stride = (unsigned long long)gridDim.x * blockDim.x;
	const unsigned long long lo = (unsigned long long)a.self * a.slice, hi = lo + a.slice;
	for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += stride)
	{
		const unsigned long long at = w << 4;
		if (at >= lo && at < hi) continue;
		const unsigned long long j = at / a.slice;
		*(uint4*)(a.dst + at) = *(const uint4*)(a.src[j] + at);
	}
	/* tail shorter than 16 bytes */
	if (blockIdx.x == 0 && threadIdx.x < (a.bytes & 15ull))
	{
		const unsigned long long at = (words << 4) + threadIdx.x;
		if (!(at >= lo && at < hi)) a.dst[at] = a.src[at / a.slice][at];
	}
}

/* ---- glGenerateMipmap (swgl.c:2129-2171): one level of the 2x2 box chain, a thread per texel.  The
 * previous level is addressed with a row stride of 2 * CurWidth texels, as in the reference (for an odd
 * width that is NOT the previous level's own stride); the sum starts from 0.0f and adds the four texels
 * in the reference's order (sY outer, sX inner), then divides by 4.0f. ---- */
__global__ void __launch_bounds__(256) k_mipmap_box(const void* __restrict__ prev, int prev_is_u8, const float* __restrict__ lut255,
                                                    int fpp, int cw, int ch, float* __restrict__ out)
{
	const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y;
	if (x >= cw || y >= ch) return;
	for (int k = 0; k < fpp; k++)
	{
		float acc = 0.0f;
		for (int sy = 0; sy < 2; sy++)
			for (int sx = 0; sx < 2; sx++)
			{
				const size_t at = ((size_t)(x * 2 + sx) + (size_t)(y * 2 + sy) * (size_t)cw * 2u) * (size_t)fpp + (size_t)k;
				acc += prev_is_u8 ? lut255[((const uint8_t*)prev)[at]] : ((const float*)prev)[at];
			}
		out[((size_t)x + (size_t)y * (size_t)cw) * (size_t)fpp + (size_t)k] = acc / 4.0f;
	}
}

/* ---- largest index of an element buffer (decides how many vertices an indexed draw shades) ---- */
/* R<<24|G<<16|B<<8|A words (swgl.c:3455-3460) -> bytes R, G, B, A in memory */
__global__ void __launch_bounds__(256) k_pack_rgba8(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		out[i] = __byte_perm(in[i], 0u, 0x0123);
}

__global__ void __launch_bounds__(256) k_max_index(const uint32_t* __restrict__ idx, size_t n, uint32_t* out)
{
	uint32_t m = 0;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = max(m, __ldg(idx + i));
	for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
	if ((threadIdx.x & 31u) == 0 && m) atomicMax(out, m);
}

/* ---- self-test of the shared-reciprocal division (swgl_dev_math.cuh) against `/` on the three
 * operand domains frag_weights_fast() relies on; counts results that differ in any bit ---- */
__device__ __forceinline__ uint64_t splitmix64(uint64_t& s)
{
	uint64_t z = (s += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t test_mantissa(uint64_t r)
{
	const uint32_t m = (uint32_t)(r >> 8) & 0x7fffffu;
	switch ((uint32_t)r & 15u)
	{
	case 0: return 0u;
	case 1: return 0x7fffffu;
	case 2: return 0x7ffffeu;
	case 3: return 1u;
	case 4: return 0x400000u;
	case 5: return 1u << (m % 23u);
	case 6: return 0x7fffffu ^ (1u << (m % 23u));
	default: return m;
	}
}

__device__ __forceinline__ float test_float(uint32_t sign, int exp2, uint32_t mant, bool integer_valued)
{
	if (integer_valued && exp2 < 23) mant &= ~((1u << (23 - exp2)) - 1u);
	return __uint_as_float((sign << 31) | ((uint32_t)(exp2 + 127) << 23) | mant);
}

__global__ void __launch_bounds__(256) k_selftest_division(uint64_t n, uint64_t seed, unsigned long long* mismatches)
{
	unsigned long long bad = 0;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
	{
		uint64_t s = seed ^ (i * 0xD1342543DE82EF95ull);
		const uint64_t r0 = splitmix64(s), r1 = splitmix64(s), r2 = splitmix64(s);
		const uint32_t cls = (uint32_t)(i % 3u);
		/* dividend exponent range, divisor exponent range per domain (see frag_weights_fast) */
		const int xlo = cls == 0 ? 0 : cls == 1 ? -86 : -100, xhi = cls == 0 ? 63 : cls == 1 ? 64 : 78;
		const int ylo = cls == 0 ? 0 : cls == 1 ? -14 : -30, yhi = cls == 0 ? 62 : cls == 1 ? 13 : 23;
		const int ex = xlo + (int)((r2 >> 8) % (uint64_t)(xhi - xlo + 1));
		const int ey = ylo + (int)((r2 >> 24) % (uint64_t)(yhi - ylo + 1));
		float x = test_float((uint32_t)(r2 & 1u), ex, test_mantissa(r0), cls == 0);
		const float y = test_float((uint32_t)((r2 >> 1) & 1u), ey, test_mantissa(r1), cls == 0);
		if (((r2 >> 2) & 15u) == 0) x = __uint_as_float((uint32_t)(r2 & 1u) << 31);    /* +-0 */
		const float fast = div_shared(x, y, rcp_refined(y));
		const float ref = x / y;
		if (__float_as_uint(fast) != __float_as_uint(ref)) bad++;
	}
	for (int o = 16; o > 0; o >>= 1) bad += __shfl_down_sync(0xffffffffu, bad, o);
	if ((threadIdx.x & 31u) == 0 && bad) atomicAdd(mismatches, bad);
}

/* ---- near clip + snap + set-up + span walk + binning counts, one thread per triangle ---- */
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


/* An entry beyond the tile's K inline slots: into the overflow pool (a few deep tiles must not size every
 * list).  The draw is dropped and re-issued with larger scratch only when the pool itself is full or too many
 * tiles are deep; both are known when the set-up kernels have finished, before any pixel is touched. */
__device__ __noinline__ void bin_overflow(const DrawParams& P, uint32_t tile, uint32_t entry, uint32_t slot)
{
	if (slot == P.bin_cap && atomicAdd(&P.ctr->hot_tiles, 1u) >= P.max_hot) atomicOr(&P.ctr->overflow, 1u);
	const uint32_t o = atomicAdd(&P.ctr->ov_cursor, 1u);
	if (o < P.ov_cap) P.ov_pool[o] = make_uint2(tile, entry);
	else atomicOr(&P.ctr->overflow, 1u);
}

/* list slot for primitive `pid` in `tile`, from the tile's atomic cursor */
__device__ __forceinline__ void bin_insert(const DrawParams& P, uint32_t tile, uint32_t entry)
{
	const uint32_t slot = atomicAdd(&P.tile_count[tile], 1u);
	if (slot < P.bin_cap) P.pairs[(size_t)tile * P.bin_cap + slot] = entry;
	else bin_overflow(P, tile, entry, slot);
}

/* The walk of a primitive that is not binned by its vertex extent (swgl.c:3356-3361, 3466-3471): tall,
 * wide or near-clipped ones.  Out of line and by value: the common path of k_setup_bin (short narrow
 * primitives of a fine mesh) keeps none of this in registers; the slopes are only needed here. */
__device__ __noinline__ void setup_walk(const DrawParams& P, float c0x, float c0y, float c1x, float c1y, float c2x, float c2y,
                                        int ys, int ye, uint32_t band, uint32_t entry, uint32_t tr_hi)
{
	TriSorted ts;
	ts.c0x = c0x; ts.c0y = c0y; ts.c1x = c1x; ts.c1y = c1y; ts.c2x = c2x; ts.c2y = c2y; ts.ys = ys; ts.ye = ye;
	TriWalk w;
	tri_slopes(ts, w);
	float x0 = w.c0x, x1 = w.c0x, s1 = w.s1;
	bool switched = false;
	uint32_t tr = tr_hi;
	int band_last_y = P.ytop - (int)(tr << P.th_shift);   /* last raster row of this band */
	if (band != 0xffffffffu && !P.inline_tall)
	{
		/* Tall primitive of a big-triangle draw: only the serial part of the walk runs here -- the two
		 * float recurrences, sampled on entering every tile row.  The spans of each tile row (and the
		 * tile columns they touch) are worked out by k_bin_tall, one thread per (primitive, tile row),
		 * so a triangle hundreds of rows tall does not keep one thread busy for all of them.
		 *
		 * The state on entering row y is c0x (+) s0, (y - ys) times, and for the second edge c0x (+) s1
		 * up to the switch row, c1x (+) s2 after it (the switch happens after the first row y with
		 * (float)y + 1 >= c1y has been drawn, swgl.c:3466-3471); the additions are replayed one by one,
		 * in the reference's order, in loops that carry nothing else. */
		const int c1yi = (w.c1y >= 2147483648.0f) ? 0x7fffffff : (w.c1y <= -2147483648.0f) ? (int)0x80000000 : (int)w.c1y;
		const int ysw = (c1yi <= w.ys) ? w.ys : c1yi - 1;            /* row after which the switch happens */
		int y = w.ys;
		for (;;)
		{
			BandEntry e;
			e.x0 = x0; e.x1 = x1; e.prim = entry;
			e.cols = (owns_tile_row(P, tr) && !(P.diag & 1u)) ? tr : 0xffffffffu;
			P.bands[band + (tr_hi - tr)] = e;
			const int y_end = min(band_last_y, w.ye - 1);             /* last row of this band */
			if (y_end >= w.ye - 1) break;
			/* advance to the state on entering row y_end + 1 */
			int n = y_end - y + 1;
			if (ysw >= y && ysw <= y_end)
			{
				for (int i = ysw - y; i > 0; i--) { x0 += w.s0; x1 += s1; }
				s1 = w.s2; x1 = w.c1x;
				x0 += w.s0; x1 += s1;
				n = y_end - ysw;
			}
			for (int i = n; i > 0; i--) { x0 += w.s0; x1 += s1; }
			y = y_end + 1;
			tr--; band_last_y += 1 << P.th_shift;
		}
		return;
	}
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
			/* span-exact tile columns of this band */
			const bool hit = cmax >= 0 && owns_tile_row(P, tr) && !(P.diag & 1u);
			const uint32_t c0 = (uint32_t)max(cmin, 0) >> SWGL_TILE_SHIFT, c1 = (uint32_t)max(cmax, 0) >> SWGL_TILE_SHIFT;
			if (band != 0xffffffffu)
			{
				BandEntry e;
				e.x0 = ex0; e.x1 = ex1; e.prim = entry;
				e.cols = hit ? (c0 | (c1 << 11) | (tr << 22)) : 0xffffffffu;
				P.bands[band + (tr_hi - tr)] = e;
				/* few tall primitives in a mesh of small ones: inserted here, k_bin_tall is not launched */
				if (P.inline_tall && hit) for (uint32_t cx = c0; cx <= c1; cx++) bin_insert(P, tr * P.tiles_x + cx, entry);
			}
			else if (hit)
				for (uint32_t cx = c0; cx <= c1; cx++) bin_insert(P, tr * P.tiles_x + cx, entry);   /* short near-clipped primitive */
			tr--; band_last_y += 1 << P.th_shift;
			ex0 = x0; ex1 = x1; cmin = 0x7fffffff; cmax = -1;
		}
	}
}


/* One primitive: divide + viewport snap, set-up, span walk.  Returns 1 if the primitive is live
 * (reaches the rasteriser on this rank).
 *   tall (more than two tile heights): one BandEntry per tile row (walk state + tile columns);
 *        k_bin_tall turns the entries into list insertions, one thread per entry;
 *   short (at most 3 tile rows): the tile columns of each band come back in pk[0..2]
 *        (c0 | c1 << 16, 0xffffffff = none) for the warp-aggregated insertion of the caller,
 *        or are inserted right here when INLINE_INSERT. */
template <bool INLINE_INSERT>
__device__ __forceinline__ uint32_t setup_one_prim(const DrawParams& P, uint32_t pid, float4 a, float4 b, float4 c,
                                                   uint32_t va, uint32_t vb, uint32_t vc,
                                                   uint32_t& tr_top, uint32_t& pk0, uint32_t& pk1, uint32_t& pk2,
                                                   uint32_t& rec_band, bool& rec_pending)
{
	/* a, b, c are already ((float)X, (float)Y, z_clip, w_clip) */
	TriSorted ts;
	if (!tri_rows(a, b, c, P, ts)) return 0u;
	/* rows [ys, ye) -> storage rows ytop-ys (bottom-most) .. ytop-ye+1: tile rows hi..lo */
	const uint32_t tr_hi = (uint32_t)(P.ytop - ts.ys) >> P.th_shift;
	const uint32_t tr_lo = (uint32_t)(P.ytop - (ts.ye - 1)) >> P.th_shift;
	if (P.n_ranks > 1)
	{
		/* sort-first: a primitive none of whose tile rows belong to this rank is dropped here */
		bool mine = false;
		for (uint32_t tr = tr_lo; tr <= tr_hi && !mine; tr++) mine = owns_tile_row(P, tr);
		if (!mine) return 0u;
	}
	uint32_t band = 0xffffffffu;
	/* short = at most two tile heights = at most 3 tile rows; narrow = the three vertices within 64 columns.
	 * Unclipped primitives that are not both get band entries: their tile lists are filled per tile row
	 * (k_bin_tall, or inline for the few of a small-triangle draw) instead of one list at a time here. */
	const float xmin = fminf(fminf(a.x, b.x), c.x), xmax = fmaxf(fmaxf(a.x, b.x), c.x);
	const bool narrow = xmin >= -32768.0f && xmax <= 32768.0f && xmax - xmin <= 64.0f;
	if (ts.ye - ts.ys > (2 << P.th_shift) || (!INLINE_INSERT && !narrow))
	{
		const uint32_t nb = tr_hi - tr_lo + 1u;
		/* one cursor round trip for the lanes of the warp that are here together (a draw of big triangles sends
		 * every lane through this: a thousand same-address atomics that return a value take microseconds) */
		{
			const uint32_t act = __activemask(), lane_id = threadIdx.x & 31u;
			uint32_t before = 0, total = 0;
			for (uint32_t m = act; m; m &= m - 1u)
			{
				const uint32_t j = (uint32_t)__ffs(m) - 1u;
				const uint32_t v = __shfl_sync(act, nb, j);
				total += v;
				if (j < lane_id) before += v;
			}
			const uint32_t leader = (uint32_t)__ffs(act) - 1u;
			uint32_t base = 0;
			if (lane_id == leader) base = atomicAdd(&P.ctr->band_cursor, total);
			band = __shfl_sync(act, base, leader) + before;
		}
		if ((unsigned long long)band + nb > (unsigned long long)P.cap_bands) { atomicOr(&P.ctr->overflow, 2u); return 0u; }
	}
	/* list entry; the record is written unless the consumer can gather the primitive itself */
	const bool has_record = INLINE_INSERT || band != 0xffffffffu || !P.lean_prims || va >= P.n_shade || vb >= P.n_shade || vc >= P.n_shade;
	const uint32_t entry = (pid << 1) | (has_record ? 1u : 0u);
	if (has_record)
	{
		if (INLINE_INSERT)
		{
			Prim* out = prim_at(P, pid);
			out->v[0] = a; out->v[1] = b; out->v[2] = c;
			*(uint4*)out->vid = make_uint4(va, vb, vc, band);
		}
		else { rec_band = band; rec_pending = true; }   /* the caller stores the warp's records together */
	}
	tr_top = tr_hi;

	if (!INLINE_INSERT && band == 0xffffffffu)
	{
		/* Short and narrow (the bulk of a fine mesh): bin by the x extent of the three vertices instead of
		 * walking the spans here -- the raster kernels walk them anyway and clamp to their tile.  The
		 * walk's x0/x1 stay inside the hull of the vertex columns up to the rounding of at most two tile
		 * heights of slope additions: with |x| <= 2^15 and an extent of at most 64 pixels that is below
		 * 0.25 pixel, so [xmin - 1, xmax + 1] covers every pixel row_span() can produce. */
		{
			const int lo = max((int)xmin - 1, max((int)P.fvx, 0));
			const int hi = min((int)xmax + 1, min((int)ceilf(P.xlimit), (int)P.W) - 1);
			if (lo <= hi && !(P.diag & 1u))
			{
				const uint32_t pk = ((uint32_t)lo >> SWGL_TILE_SHIFT) | (((uint32_t)hi >> SWGL_TILE_SHIFT) << 16);
				if (owns_tile_row(P, tr_hi)) pk0 = pk;
				if (tr_hi >= tr_lo + 1u && owns_tile_row(P, tr_hi - 1u)) pk1 = pk;
				if (tr_hi >= tr_lo + 2u && owns_tile_row(P, tr_hi - 2u)) pk2 = pk;
			}
			return 1u;
		}
	}

	setup_walk(P, ts.c0x, ts.c0y, ts.c1x, ts.c1y, ts.c2x, ts.c2y, ts.ys, ts.ye, band, entry, tr_hi);
	return 1u;
}

/* near-plane clipping of a triangle that is not entirely inside (swgl.c:563-696): rare, kept out
 * of line so the common path stays in registers */
__device__ __noinline__ uint32_t setup_clipped(const DrawParams& P, uint32_t t, const float4* p, const uint32_t* sid, uint32_t in_mask)
{
	int in_idx[3], out_idx[3], n_in = 0, n_out = 0;
	for (int j = 0; j < 3; j++) { if ((in_mask >> j) & 1u) in_idx[n_in++] = j; else out_idx[n_out++] = j; }
	const uint32_t new0 = P.clip_vid_base + 2u * t, new1 = new0 + 1u;
	float t0, t1;
	uint32_t d0, d1, d2, d3, d4;
	bool d5;
	if (n_in == 1)
	{
		const int a = in_idx[0];
		const float4 q1 = near_intersect(p[a], p[out_idx[0]], t0);
		const float4 q2 = near_intersect(p[a], p[out_idx[1]], t1);
		lerp_vary(P, new0, sid[a], sid[out_idx[0]], t0);
		lerp_vary(P, new1, sid[a], sid[out_idx[1]], t1);
		return setup_one_prim<true>(P, 2u * t, to_screen(p[a], P), to_screen(q1, P), to_screen(q2, P), sid[a], new0, new1, d0, d1, d2, d3, d4, d5);
	}
	if (n_in == 2)
	{
		const int a = in_idx[0], b = in_idx[1], o = out_idx[0];
		const float4 q0 = near_intersect(p[a], p[o], t0);
		const float4 q1 = near_intersect(p[b], p[o], t1);
		lerp_vary(P, new0, sid[a], sid[o], t0);
		lerp_vary(P, new1, sid[b], sid[o], t1);
		const float4 sq0 = to_screen(q0, P);
		uint32_t live = setup_one_prim<true>(P, 2u * t, to_screen(p[a], P), to_screen(p[b], P), sq0, sid[a], sid[b], new0, d0, d1, d2, d3, d4, d5);
		live += setup_one_prim<true>(P, 2u * t + 1u, to_screen(p[b], P), sq0, to_screen(q1, P), sid[b], new0, new1, d0, d1, d2, d3, d4, d5);
		return live;
	}
	return 0u;
}

/* One group of 32 x warps triangles: triangle t of this thread with its vertices already loaded. */
__device__ __forceinline__ void setup_group(const DrawParams& P, float4 (*stage)[32][4], uint32_t t, uint32_t lane, uint32_t wid,
                                            float4 p0, float4 p1, float4 p2, uint32_t s0, uint32_t s1, uint32_t s2)
{
	uint32_t live = 0, tr_top = 0, pk[3] = { 0xffffffffu, 0xffffffffu, 0xffffffffu };
	uint32_t rec_band = 0xffffffffu;
	bool rec_pending = false;
	if (t < P.ntri)
	{
		if (P.diag & 4u) { if (p0.x + p1.x + p2.x == 12345.678f) P.ctr->prims_out = 1; return; }
		if (P.diag & 16u) { if (s0 + s1 + s2 == 0x12345678u) P.ctr->prims_out = 1; return; }
		/* ClipTriangleAgainstNearPlane (swgl.c:532-561): inside iff z >= -w */
		const uint32_t in_mask = (p0.z >= -p0.w ? 1u : 0u) | (p1.z >= -p1.w ? 2u : 0u) | (p2.z >= -p2.w ? 4u : 0u);
		if (in_mask == 7u) live = setup_one_prim<false>(P, 2u * t, p0, p1, p2, s0, s1, s2, tr_top, pk[0], pk[1], pk[2], rec_band, rec_pending);
		else if (in_mask != 0u)
		{
			/* back to clip space for the intersection arithmetic */
			const float2 zz = make_float2(0.0f, 0.0f);
			const float2 c0 = (s0 < P.n_shade) ? P.clip_xy[s0] : zz, c1 = (s1 < P.n_shade) ? P.clip_xy[s1] : zz, c2 = (s2 < P.n_shade) ? P.clip_xy[s2] : zz;
			const float4 p[3] = { make_float4(c0.x, c0.y, p0.z, p0.w), make_float4(c1.x, c1.y, p1.z, p1.w), make_float4(c2.x, c2.y, p2.z, p2.w) };
			const uint32_t sid[3] = { s0, s1, s2 };
			live = setup_clipped(P, t, p, sid, in_mask);
		}
	}

	/* a warp none of whose primitives is live (sort-first: all in other ranks' bands, or all culled)
	 * has nothing to store or insert */
	if (!__any_sync(0xffffffffu, live != 0u)) return;

	/* The unclipped primitives' 64-byte records leave as whole 512-byte runs: a thread storing its own
	 * record would touch 32 half-written sectors per store instruction. */
	{
		const uint32_t recs = __ballot_sync(0xffffffffu, rec_pending);
		if (recs)
		{
			if (rec_pending)
			{
				stage[wid][lane][0] = p0; stage[wid][lane][1] = p1; stage[wid][lane][2] = p2;
				stage[wid][lane][3] = make_float4(__uint_as_float(s0), __uint_as_float(s1), __uint_as_float(s2), __uint_as_float(rec_band));
			}
			__syncwarp();
			float4* out = (float4*)(P.prims + (t - lane));
#pragma unroll
			for (uint32_t i = 0; i < 4; i++)
			{
				const uint32_t idx = i * 32u + lane, rec = idx >> 2;
				if ((recs >> rec) & 1u) out[idx] = stage[wid][rec][idx & 3u];
			}
		}
	}

	/* warp-aggregated list insertion for the short, unclipped primitives: neighbouring triangles
	 * mostly land in the same tile, so lanes that want the same tile share one atomicAdd.  The first
	 * tile column of all three bands goes first, with the three atomics in flight together (the
	 * round trip of an atomic that returns a value is what this kernel waits for); further columns
	 * (primitives that straddle a tile boundary in x) follow one at a time. */
	const uint32_t short_entry = (4u * t) | (rec_pending ? 1u : 0u);   /* primitive 2t; pk[] is only set for short unclipped ones */
	uint32_t base[3], peers[3], tile[3];
#pragma unroll
	for (int k = 0; k < 3; k++)
	{
		const bool have = pk[k] != 0xffffffffu;
		tile[k] = have ? (tr_top - (uint32_t)k) * P.tiles_x + (pk[k] & 0xffffu) : 0xffffffffu;
		peers[k] = __match_any_sync(0xffffffffu, tile[k]);
		base[k] = 0;
		if (have && lane == (uint32_t)__ffs(peers[k]) - 1u) base[k] = atomicAdd(&P.tile_count[tile[k]], (uint32_t)__popc(peers[k]));
	}
#pragma unroll
	for (int k = 0; k < 3; k++)
	{
		const uint32_t b0 = __shfl_sync(0xffffffffu, base[k], __ffs(peers[k]) - 1);
		if (pk[k] != 0xffffffffu)
		{
			const uint32_t slot = b0 + (uint32_t)__popc(peers[k] & ((1u << lane) - 1u));
			if (slot < P.bin_cap) P.pairs[(size_t)tile[k] * P.bin_cap + slot] = short_entry;
			else bin_overflow(P, tile[k], short_entry, slot);
		}
	}
#pragma unroll
	for (int k = 0; k < 3; k++)
	{
		const uint32_t c1 = pk[k] >> 16;
		uint32_t cx = (pk[k] & 0xffffu) + 1u;
		const bool valid = pk[k] != 0xffffffffu;
		while (__any_sync(0xffffffffu, valid && cx <= c1))
		{
			const bool have = valid && cx <= c1;
			const uint32_t tl = have ? (tr_top - (uint32_t)k) * P.tiles_x + cx : 0xffffffffu;
			const uint32_t pr = __match_any_sync(0xffffffffu, tl);
			const uint32_t leader = (uint32_t)__ffs(pr) - 1u;
			uint32_t bs = 0;
			if (have && lane == leader) bs = atomicAdd(&P.tile_count[tl], (uint32_t)__popc(pr));
			bs = __shfl_sync(0xffffffffu, bs, leader);
			if (have)
			{
				const uint32_t slot = bs + (uint32_t)__popc(pr & ((1u << lane) - 1u));
				if (slot < P.bin_cap) P.pairs[(size_t)tl * P.bin_cap + slot] = short_entry;
				else bin_overflow(P, tl, short_entry, slot);
			}
			cx++;
		}
	}

	/* primitives that reached the rasteriser: one atomic per warp */
	for (int o = 16; o > 0; o >>= 1) live += __shfl_down_sync(0xffffffffu, live, o);
	if (lane == 0 && live) atomicAdd(&P.ctr->prims_out, live);
}

/* the element indices of triangle t (stream positions 3t .. 3t+2 for glDrawArrays) */
__device__ __forceinline__ void tri_indices(const DrawParams& P, uint32_t t, uint32_t& s0, uint32_t& s1, uint32_t& s2)
{
	s0 = 3u * t; s1 = s0 + 1u; s2 = s0 + 2u;
	if (P.ibo && t < P.ntri)
	{
		const unsigned long long at = (unsigned long long)(long long)P.first + s0;
		s0 = (at < P.ibo_count) ? __ldg(P.ibo + at) : 0xffffffffu;
		s1 = (at + 1 < P.ibo_count) ? __ldg(P.ibo + at + 1) : 0xffffffffu;
		s2 = (at + 2 < P.ibo_count) ? __ldg(P.ibo + at + 2) : 0xffffffffu;
	}
}

__device__ __forceinline__ void tri_clip(const DrawParams& P, uint32_t t, uint32_t s0, uint32_t s1, uint32_t s2, float4& p0, float4& p1, float4& p2)
{
	const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	p0 = p1 = p2 = zero;
	if (t >= P.ntri) return;
	p0 = (s0 < P.n_shade) ? P.clip[s0] : to_screen(zero, P);
	p1 = (s1 < P.n_shade) ? P.clip[s1] : to_screen(zero, P);
	p2 = (s2 < P.n_shade) ? P.clip[s2] : to_screen(zero, P);
}

/* The set-up kernel is a chain of three dependent global round trips per triangle (indices -> gathered
 * vertices -> list cursors) and was latency-bound at 45 % issue with one triangle per thread.  It is now a
 * resident grid whose threads walk the triangle stream in groups of (grid x 128) with a two-deep software
 * pipeline: while group g is set up and binned, the vertices of group g+1 and the indices of group g+2 are in
 * flight.  Launched with programmatic stream serialisation behind k_vertex: the first indices are requested
 * before the wait, clip[] is only read after it. */
#ifndef SETUP_CTAS_PER_SM
#define SETUP_CTAS_PER_SM 5
#endif
/* PIPELINED = false: one triangle per thread, the form every draw uses.  PIPELINED = true: the resident,
 * software-pipelined form, kept behind the "setup_pipelined" option as a measured alternative: C4 on one device
 * 36.7 us against 33.6 us, and for the share of one rank of 8 / 4 / 2 (most warps only load their triangles to
 * find that none is theirs) 23.6 / 26.4 / 33.7 us against 19.6 / 21.4 / 25.4 us (tools/stripe_probe.py) -- hiding
 * the first link of the chain costs the occupancy that hides the other two. */
template <bool PIPELINED>
__global__ void __launch_bounds__(128, PIPELINED ? SETUP_CTAS_PER_SM : 8) k_setup_bin(const __grid_constant__ DrawParams P)
{
	__shared__ float4 stage[4][32][4];       /* the warp's primitive records on their way out */
	const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	if (PIPELINED)
	{
	const uint32_t stride = gridDim.x * blockDim.x;
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	/* whole warps leave together: every shuffle / vote below is warp-wide */
	const uint32_t t_end = (P.ntri + 31u) & ~31u;
	uint32_t a0, a1, a2, b0, b1, b2;        /* indices of the group being processed next / the one after */
	float4 p0, p1, p2;
	tri_indices(P, t, a0, a1, a2);
	tri_indices(P, t + stride, b0, b1, b2);
	pdl_wait();                              /* everything below reads the vertex kernel's output */
	tri_clip(P, t, a0, a1, a2, p0, p1, p2);
	for (; (t & ~31u) < t_end; t += stride)
	{
		/* requests for the next two groups go out before this group's dependent work */
		float4 q0, q1, q2;
		tri_clip(P, t + stride, b0, b1, b2, q0, q1, q2);
		uint32_t c0, c1, c2;
		tri_indices(P, t + 2u * stride, c0, c1, c2);
		setup_group(P, stage, t, lane, wid, p0, p1, p2, a0, a1, a2);
		__syncwarp();                        /* the stage rows of this warp are reused by the next group */
		p0 = q0; p1 = q1; p2 = q2;
		a0 = b0; a1 = b1; a2 = b2;
		b0 = c0; b1 = c1; b2 = c2;
	}
	}
	else
	{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	float4 p0, p1, p2;
	uint32_t s0 = 0, s1 = 0, s2 = 0;
	if (t < P.ntri) tri_vertices<true>(P, t, p0, p1, p2, s0, s1, s2);
	setup_group(P, stage, t, lane, wid, p0, p1, p2, s0, s1, s2);
	}
}

/* The tile's list length; the cursor is re-armed (zeroed) for the next draw.  Whole CTA calls it. */
__device__ __forceinline__ uint32_t take_tile_list(const DrawParams& P, uint32_t tile)
{
	__shared__ uint32_t n_s;
	if (threadIdx.x == 0)
	{
		const uint32_t n = P.tile_count[tile];
		if (n) P.tile_count[tile] = 0u;
		if (n > P.bin_cap) atomicMax(&P.ctr->max_list, n);
		else if (n && P.count_fragments) atomicAdd(&P.ctr->pair_total, (unsigned long long)n);
		n_s = n;
	}
	__syncthreads();
	return min(n_s, P.bin_cap);
}

/* ---- tall or wide primitives of big-triangle draws: one WARP per (primitive, tile row) band entry.
 * Lane r replays the walk from the state the set-up kernel sampled on entering the band to its own
 * row (swgl.c:3356-3361, 3466-3471; at most 31 additions), the warp reduces the columns the spans
 * touch, and the lanes then insert the primitive into one tile list each ---- */
__global__ void __launch_bounds__(256) k_bin_tall(const __grid_constant__ DrawParams P)
{
	if (P.ctr->overflow & 2u) return;
	const uint32_t total = P.ctr->band_cursor;
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < total; e += warps)
	{
		const BandEntry be = P.bands[e];
		if (be.cols == 0xffffffffu) continue;
		const uint32_t tr = be.cols;
		const Prim* q = prim_at(P, be.prim >> 1);       /* these primitives always have a record */
		TriWalk w;
		if (!tri_setup(q->v[0], q->v[1], q->v[2], P, w)) continue;
		const int band_last_y = P.ytop - (int)(tr << P.th_shift);
		const int y_in = max(w.ys, band_last_y - ((1 << P.th_shift) - 1)), y_out = min(w.ye - 1, band_last_y);
		float x0, x1, s1;
		bool switched;
		walk_to_row(P, w, q->band, tr, y_in, x0, x1, s1, switched);
		int cmin = 0x7fffffff, cmax = -1;
		const int y = y_in + (int)lane;
		if (y <= y_out)
		{
			for (int yy = y_in; yy < y; yy++)
			{
				if (!switched && (float)yy + 1.0f >= w.c1y) { switched = true; s1 = w.s2; x1 = w.c1x; }
				x0 += w.s0; x1 += s1;
			}
			int xa, xb;
			row_span(x0, x1, P, xa, xb);
			if (xa < xb) { cmin = xa; cmax = xb - 1; }
		}
		cmin = __reduce_min_sync(0xffffffffu, cmin);
		cmax = __reduce_max_sync(0xffffffffu, cmax);
		if (cmax < 0) continue;
		const uint32_t c0 = (uint32_t)max(cmin, 0) >> SWGL_TILE_SHIFT, c1 = (uint32_t)max(cmax, 0) >> SWGL_TILE_SHIFT;
		for (uint32_t cx = c0 + lane; cx <= c1; cx += 32) bin_insert(P, tr * P.tiles_x + cx, be.prim);
	}
}

/* ---- set-up of a draw of BIG triangles (inline_tall = 0, e.g. BASELINE config 1): one WARP per input triangle.
 * A tall primitive crosses dozens of tile rows; with a thread per triangle each thread walked them one after the
 * other (and a second kernel, k_bin_tall, then worked through the band entries), on a grid of a few CTAs.  Here
 * lane b takes tile row b of the primitive: it replays the two float recurrences from the primitive's first row
 * to its own (the additions are the reference's, in its order, swgl.c:3356-3361, 3466-3471 -- only nobody waits
 * for anybody), stores the band entry the rasteriser starts from, walks the rows of its band for the columns the
 * spans touch and inserts the primitive into those tiles.  Near-clipped triangles (rare) keep the serial path on
 * lane 0, which inserts as it walks (inline_tall): k_bin_tall is not launched for these draws. ---- */
__global__ void __launch_bounds__(256) k_setup_big(const __grid_constant__ DrawParams P)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (t >= P.ntri) return;
	float4 p0, p1, p2;
	uint32_t s0, s1, s2;
	tri_vertices<true>(P, t, p0, p1, p2, s0, s1, s2);        /* every lane the same addresses: broadcast loads */
	const uint32_t in_mask = (p0.z >= -p0.w ? 1u : 0u) | (p1.z >= -p1.w ? 2u : 0u) | (p2.z >= -p2.w ? 4u : 0u);
	if (in_mask != 7u)
	{
		if (in_mask != 0u && lane == 0)
		{
			const float2 zz = make_float2(0.0f, 0.0f);
			const float2 c0 = (s0 < P.n_shade) ? P.clip_xy[s0] : zz, c1 = (s1 < P.n_shade) ? P.clip_xy[s1] : zz, c2 = (s2 < P.n_shade) ? P.clip_xy[s2] : zz;
			const float4 p[3] = { make_float4(c0.x, c0.y, p0.z, p0.w), make_float4(c1.x, c1.y, p1.z, p1.w), make_float4(c2.x, c2.y, p2.z, p2.w) };
			const uint32_t sid[3] = { s0, s1, s2 };
			const uint32_t live = setup_clipped(P, t, p, sid, in_mask);
			if (live) atomicAdd(&P.ctr->prims_out, live);
		}
		return;
	}
	TriSorted ts;
	if (!tri_rows(p0, p1, p2, P, ts)) return;
	const uint32_t tr_hi = (uint32_t)(P.ytop - ts.ys) >> P.th_shift;
	const uint32_t tr_lo = (uint32_t)(P.ytop - (ts.ye - 1)) >> P.th_shift;
	if (P.n_ranks > 1)
	{
		bool mine = false;
		for (uint32_t tr = tr_lo; tr <= tr_hi && !mine; tr++) mine = owns_tile_row(P, tr);
		if (!mine) return;
	}
	const uint32_t pid = 2u * t;
	const float xmin = fminf(fminf(p0.x, p1.x), p2.x), xmax = fmaxf(fmaxf(p0.x, p1.x), p2.x);
	const bool narrow = xmin >= -32768.0f && xmax <= 32768.0f && xmax - xmin <= 64.0f;
	if (lane == 0) atomicAdd(&P.ctr->prims_out, 1u);
	if (ts.ye - ts.ys <= (2 << P.th_shift) && narrow)
	{
		/* short and narrow: binned by the x extent of its vertices, like in k_setup_bin */
		const bool has_record = !P.lean_prims || s0 >= P.n_shade || s1 >= P.n_shade || s2 >= P.n_shade;
		const uint32_t entry = (pid << 1) | (has_record ? 1u : 0u);
		if (has_record && lane < 4u)
		{
			float4* out = (float4*)prim_at(P, pid);
			out[lane] = lane == 0 ? p0 : lane == 1 ? p1 : lane == 2 ? p2
			          : make_float4(__uint_as_float(s0), __uint_as_float(s1), __uint_as_float(s2), __uint_as_float(0xffffffffu));
		}
		const int lo = max((int)xmin - 1, max((int)P.fvx, 0));
		const int hi = min((int)xmax + 1, min((int)ceilf(P.xlimit), (int)P.W) - 1);
		if (lo > hi || (P.diag & 1u)) return;
		const uint32_t c0 = (uint32_t)lo >> SWGL_TILE_SHIFT, nc = ((uint32_t)hi >> SWGL_TILE_SHIFT) - c0 + 1u;
		const uint32_t rows = tr_hi - tr_lo + 1u;           /* at most 3 */
		for (uint32_t j = lane; j < rows * nc; j += 32)
		{
			const uint32_t tr = tr_hi - j / nc;
			if (owns_tile_row(P, tr)) bin_insert(P, tr * P.tiles_x + c0 + j % nc, entry);
		}
		return;
	}
	/* tall or wide: a band entry per tile row */
	const uint32_t nb = tr_hi - tr_lo + 1u;
	uint32_t band = 0;
	if (lane == 0) band = atomicAdd(&P.ctr->band_cursor, nb);
	band = __shfl_sync(0xffffffffu, band, 0);
	if ((unsigned long long)band + nb > (unsigned long long)P.cap_bands) { if (lane == 0) atomicOr(&P.ctr->overflow, 2u); return; }
	const uint32_t entry = (pid << 1) | 1u;
	if (lane < 4u)
	{
		float4* out = (float4*)prim_at(P, pid);
		out[lane] = lane == 0 ? p0 : lane == 1 ? p1 : lane == 2 ? p2
		          : make_float4(__uint_as_float(s0), __uint_as_float(s1), __uint_as_float(s2), __uint_as_float(band));
	}
	TriWalk w;
	tri_slopes(ts, w);
	const int th = 1 << P.th_shift;
	/* The state on entering row y is c0x (+) s0, (y - ys) times, and for the second edge c0x (+) s1 up to the switch
	 * row, c1x (+) s2 after it (the switch happens after the first row y with (float)y + 1 >= c1y has been drawn,
	 * swgl.c:3466-3471).  The additions are replayed one by one in loops that carry nothing else (two independent
	 * 4-cycle chains), split at the switch row; a lane's next band (32 tile rows on) continues from where its
	 * previous one started instead of from the primitive's first row. */
	const int c1yi = (w.c1y >= 2147483648.0f) ? 0x7fffffff : (w.c1y <= -2147483648.0f) ? (int)0x80000000 : (int)w.c1y;
	const int ysw = (c1yi <= w.ys) ? w.ys : c1yi - 1;            /* row after which the switch happens */
	float x0 = w.c0x, x1 = w.c0x, sl = w.s1;
	bool switched = false;
	int y_at = w.ys;                                             /* (x0, x1, sl, switched) is the state on entering y_at */
	for (uint32_t b = lane; b < nb; b += 32)
	{
		const uint32_t tr = tr_hi - b;
		const int band_last_y = P.ytop - (int)(tr << P.th_shift);
		const int y_in = max(w.ys, band_last_y - (th - 1)), y_out = min(w.ye - 1, band_last_y);
		{
			int n = y_in - y_at;
			if (!switched && ysw >= y_at && ysw < y_in)
			{
				for (int i = ysw - y_at; i > 0; i--) { x0 += w.s0; x1 += sl; }
				switched = true; sl = w.s2; x1 = w.c1x;
				x0 += w.s0; x1 += sl;
				n = y_in - ysw - 1;
			}
			for (int i = n; i > 0; i--) { x0 += w.s0; x1 += sl; }
			y_at = y_in;
		}
		const float ex0 = x0, ex1 = x1;
		const float esl = sl;
		const bool eswitched = switched;
		BandEntry e;
		e.x0 = ex0; e.x1 = ex1; e.prim = entry; e.cols = 0xffffffffu;      /* inserted right here: nothing left for k_bin_tall */
		P.bands[band + b] = e;
		if (!owns_tile_row(P, tr) || (P.diag & 1u)) continue;
		int cmin = 0x7fffffff, cmax = -1;
		{
			float a0 = ex0, a1 = ex1, as = esl;
			bool asw = eswitched;
			for (int y = y_in; y <= y_out; y++)
			{
				int xa, xb;
				row_span(a0, a1, P, xa, xb);
				if (xa < xb) { cmin = min(cmin, xa); cmax = max(cmax, xb - 1); }
				if (!asw && (float)y + 1.0f >= w.c1y) { asw = true; as = w.s2; a1 = w.c1x; }
				a0 += w.s0; a1 += as;
			}
		}
		if (cmax < 0) continue;
		const uint32_t c0 = (uint32_t)max(cmin, 0) >> SWGL_TILE_SHIFT, c1 = (uint32_t)max(cmax, 0) >> SWGL_TILE_SHIFT;
		/* eight list cursors in flight at a time: one at a time the lane would wait out a round trip per tile */
		for (uint32_t cb = c0; cb <= c1; cb += 8u)
		{
			uint32_t slot[8];
#pragma unroll
			for (uint32_t k = 0; k < 8u; k++) if (cb + k <= c1) slot[k] = atomicAdd(&P.tile_count[tr * P.tiles_x + cb + k], 1u);
#pragma unroll
			for (uint32_t k = 0; k < 8u; k++)
				if (cb + k <= c1)
				{
					const uint32_t tile = tr * P.tiles_x + cb + k;
					if (slot[k] < P.bin_cap) P.pairs[(size_t)tile * P.bin_cap + slot[k]] = entry;
					else bin_overflow(P, tile, entry, slot[k]);
				}
		}
	}
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
	/* fallback for very long lists: in-place bitonic network over global memory (L2-resident) in its
	 * "flip + half-cleaner" form -- every exchange moves the smaller key down, so the virtual
	 * 0xffffffff padding beyond n never moves and n need not be a power of two */
	uint32_t n_pow2 = 1; while (n_pow2 < n) n_pow2 <<= 1;
	for (uint32_t k = 2; k <= n_pow2; k <<= 1)
	{
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
		{
			const uint32_t l = i ^ (k - 1u);
			if (l > i && l < n) { const uint32_t a = ids[i], b = ids[l]; if (a > b) { ids[i] = b; ids[l] = a; } }
		}
		__syncthreads();
		for (uint32_t j = k >> 2; j > 0; j >>= 1)
		{
			for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
			{
				const uint32_t l = i ^ j;
				if (l > i && l < n) { const uint32_t a = ids[i], b = ids[l]; if (a > b) { ids[i] = b; ids[l] = a; } }
			}
			__syncthreads();
		}
	}
}

template <int FS>
__global__ void __launch_bounds__(SWGL_RASTER_THREADS) k_raster(const __grid_constant__ DrawParams P)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	RasterShared& S = *reinterpret_cast<RasterShared*>(smem_raw);

	const uint32_t tx = blockIdx.x, ty = blockIdx.y;
	if (!owns_tile_row(P, ty)) return;
	const uint32_t tile = ty * P.tiles_x + tx;
	const uint32_t tid = threadIdx.x;
	const uint32_t n_list = take_tile_list(P, tile);
	if (P.ctr->overflow || P.diag) return;
	const size_t list_off = (size_t)tile * P.bin_cap;
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
				const Prim pr = load_prim(P, pid);
				TriWalk w;
				tri_setup(pr.v[0], pr.v[1], pr.v[2], P, w);
				const int y_in = max(w.ys, band_first_y);
				const int y_out = min(w.ye - 1, band_last_y);
				float x0, x1, s1;
				bool switched;
				walk_to_row(P, w, pr.band, ty, y_in, x0, x1, s1, switched);
				uint32_t mask = 0;
				for (int rr = 0; rr < SWGL_TILE; rr++) S.span[rr][tid] = 0;
				for (int y = y_in; y <= y_out; y++)
				{
					int xa, xb;
					row_span(x0, x1, P, xa, xb);
					xa = min(max(xa, tile_x0), tile_x0 + SWGL_TILE) - tile_x0;
					xb = min(max(xb, tile_x0), tile_x0 + SWGL_TILE) - tile_x0;   /* xb may be INT_MIN: clamp before subtracting */
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
							f.lod = (FS == SWFS_GENERIC && P.mip_lod) ? prim_lod(P, ids[base + j]) : 0.0f;
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
/* ---- GL_POINTS (swgl.c:3496-3608) ----
 * One pixel per vertex: X = (int)(x/w * (VH/2) + (VW/2) + VX) -- the x scale really is VH/2
 * (swgl.c:3531) -- Y = (int)(y/w * (VH/2) + (VH/2) + VY), no near clip, no Y flip, depth and
 * colour are overwritten unconditionally.  Points are submitted in order, so the last point that
 * lands on a pixel wins: pass 1 takes the maximum point index per pixel, pass 2 lets the winner
 * run the fragment shader and store. */
__device__ __forceinline__ bool point_pixel(const DrawParams& P, uint32_t i, uint32_t& vid, uint32_t& pix, float& z)
{
	vid = i;
	if (P.ibo)
	{
		const unsigned long long at = (unsigned long long)(long long)P.first + i;
		vid = (at < P.ibo_count) ? __ldg(P.ibo + at) : 0xffffffffu;
	}
	float4 c = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	float2 xy = make_float2(0.0f, 0.0f);
	if (vid < P.n_shade) { c = P.clip[vid]; xy = P.clip_xy[vid]; }
	const int X = cvt_x86((fdiv(xy.x, c.w) * P.hh + P.hw) + P.fvx);
	const int Y = cvt_x86((fdiv(xy.y, c.w) * P.hh + P.hh) + P.fvy);
	if (X < 0 || X >= (int)P.W || Y < 0 || Y >= (int)P.H) return false;
	if (P.n_ranks > 1 && ((((uint32_t)Y >> 5) / P.band_rows) % P.n_ranks) != P.rank) return false;
	pix = (uint32_t)X + (uint32_t)Y * P.W;
	z = c.z;
	return true;
}

__global__ void __launch_bounds__(256) k_points_claim(const __grid_constant__ DrawParams P)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.count) return;
	uint32_t vid, pix; float z;
	if (point_pixel(P, i, vid, pix, z)) atomicMax(&P.winner[pix], i + 1u);
}

template <int FS>
__global__ void __launch_bounds__(256) k_points_write(const __grid_constant__ DrawParams P)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.count) return;
	uint32_t vid, pix; float z;
	if (!point_pixel(P, i, vid, pix, z)) return;
	if (P.winner[pix] != i + 1u) return;
	P.winner[pix] = 0u;                      /* re-armed for the next draw */
	/* varyings are copied, not interpolated (swgl.c:3537-3553) */
	const float* vv = P.vary + (size_t)vid * P.nvf;
	float o[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
	if (FS == SWFS_VARYING) { for (int k = 0; k < 4; k++) o[k] = vv[P.fs_slot + k]; }
	else if (FS == SWFS_TEXTURE)
	{
		const float4 t = sample_nearest(P.tex[P.fs_tex_unit], vv[P.fs_slot + P.fs_swz_u], vv[P.fs_slot + P.fs_swz_v]);
		o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
	}
	else
	{
		uint32_t V[SWGL_MAX_VAR_WORDS];
		for (uint32_t k = 0; k < P.fs_words; k++) V[k] = P.fs_image[k];
		for (uint32_t k = 0; k < P.n_varying; k++)
			for (uint32_t j = 0; j < P.varying[k].n_floats; j++)
				V[P.varying[k].fs_word + j] = __float_as_uint(vv[P.varying[k].slot + j]);
		ir_execute(P.fs_ops, P.fs_nops, V, P, P.mip_lod ? *P.last_level : 0.0f);   /* points: the level the last triangle left (k_last_level) */
		for (uint32_t k = 0; k < P.out_floats; k++) o[k] = __uint_as_float(V[P.out_word + k]);
	}
	const float r = RMIN(RMAX(o[0], 0.0f), 1.0f), g = RMIN(RMAX(o[1], 0.0f), 1.0f);
	const float b = RMIN(RMAX(o[2], 0.0f), 1.0f), a = RMIN(RMAX(o[3], 0.0f), 1.0f);
	uint32_t word = 0;                       /* swgl.c:3599-3603: no blend */
	word |= (uint32_t)__float2int_rz(r * 255.0f) << 24;
	word |= (uint32_t)__float2int_rz(g * 255.0f) << 16;
	word |= (uint32_t)__float2int_rz(b * 255.0f) << 8;
	word |= (uint32_t)__float2int_rz(a * 255.0f);
	P.depth[pix] = z;                        /* swgl.c:3582-3583: written as is, no test */
	P.color[pix] = word;
	if (P.peer_color) P.peer_color[pix] = word;
}

#include "swgl_raster_frag.cuh"
#include "swgl_raster_warp.cuh"

/* ---- viewports that leave the framebuffer vertically (swgl.c:3386) ----
 * The reference stores raster row y at row min(Height-1, (unsigned)(VH-1+2*VY-y)): every row whose storage row
 * would be negative or beyond the framebuffer lands on row Height-1, where the fragments of several rows of a
 * primitive -- and of all primitives -- then meet in (primitive, y) order.  The tile machinery needs one raster
 * row per storage row, so such a draw is rendered in two parts (draw_folded): the rows that map inside
 * [0, Height-2] through the ordinary kernels into a virtual framebuffer tall enough for the whole viewport, and
 * row Height-1 by this kernel: one thread per pixel column walks every primitive in submission order and, for
 * each, its aliased rows in ascending y, with the exact (slow-path) fragment arithmetic.  Correct, not fast:
 * a client with an oversized or offset viewport gets the reference's frame instead of nothing. */
__device__ __forceinline__ bool fold_aliased(const DrawParams& P, int y)
{
	const long long srow = (long long)P.ytop - (long long)y;
	return srow < 0 || srow >= (long long)P.H - 1;
}

template <int FS>
__device__ __forceinline__ void fold_prim(const DrawParams& P, int x, int count_from, const float4& a, const float4& b, const float4& c,
                                          uint32_t v0, uint32_t v1, uint32_t v2, uint32_t& col, float& dep,
                                          uint32_t& n_tested, uint32_t& n_shaded)
{
	TriWalk w;
	if (!tri_setup(a, b, c, P, w)) return;
	/* aliased rows: y > ytop (negative storage row) or y <= ytop - (H - 1) */
	const long long hi_from = (long long)P.ytop + 1, lo_to = (long long)P.ytop - ((long long)P.H - 1);
	if (!((long long)w.ye - 1 >= hi_from || (long long)w.ys <= lo_to)) return;
	BaryConst k;
	bary_setup(a, b, c, k);
	float x0 = w.c0x, x1 = w.c0x, s1 = w.s1;
	bool switched = false;
	for (int y = w.ys; y < w.ye; y++)
	{
		if (fold_aliased(P, y))
		{
			int xa, xb;
			row_span(x0, x1, P, xa, xb);
			if (x >= xa && x < xb)
			{
				if (y >= count_from) n_tested++;       /* the rows below count_from were walked -- and counted -- by the tile kernels */
				FragIn fi;
				float z;
				frag_weights(k, (float)x, (float)y, fi.u, fi.v, fi.w, z);
				if (dep == 0.0f || dep >= z)     /* swgl.c:3387 */
				{
					dep = z;
					n_shaded++;
					fi.vid0 = v0; fi.vid1 = v1; fi.vid2 = v2;
					fi.a = P.vary + (size_t)v0 * P.nvf + P.fs_slot;
					fi.b = P.vary + (size_t)v1 * P.nvf + P.fs_slot;
					fi.c = P.vary + (size_t)v2 * P.nvf + P.fs_slot;
					fi.stride = 1;
					fi.lod = (FS == SWFS_GENERIC && P.mip_lod) ? mip_level(a.x, a.y, b.x, b.y, c.x, c.y) : 0.0f;
					const float4 o = run_fragment<FS>(P, fi);
					col = blend_pack(o.x, o.y, o.z, o.w, col);
				}
			}
		}
		if (!switched && (float)y + 1.0f >= w.c1y) { switched = true; s1 = w.s2; x1 = w.c1x; }
		x0 += w.s0; x1 += s1;
	}
}

template <int FS>
__global__ void __launch_bounds__(128) k_fold_row(const __grid_constant__ DrawParams P, int x_lo, int x_hi, int count_from)
{
	const int x = x_lo + (int)(blockIdx.x * blockDim.x + threadIdx.x);
	if (x >= x_hi) return;
	const size_t pix = (size_t)(P.H - 1u) * P.W + (size_t)x;
	uint32_t col = P.color[pix];
	float dep = P.depth[pix];
	uint32_t n_tested = 0, n_shaded = 0;
	for (uint32_t t = 0; t < P.ntri; t++)
	{
		float4 p0, p1, p2;
		uint32_t s0, s1, s2;
		tri_vertices(P, t, p0, p1, p2, s0, s1, s2);
		const uint32_t in_mask = (p0.z >= -p0.w ? 1u : 0u) | (p1.z >= -p1.w ? 2u : 0u) | (p2.z >= -p2.w ? 4u : 0u);
		if (in_mask == 7u) { fold_prim<FS>(P, x, count_from, p0, p1, p2, s0, s1, s2, col, dep, n_tested, n_shaded); continue; }
		if (in_mask == 0u) continue;
		/* near clip again (swgl.c:563-696); the varyings of the new vertices were written by the set-up kernel */
		const float2 zz = make_float2(0.0f, 0.0f);
		const float2 c0 = (s0 < P.n_shade) ? P.clip_xy[s0] : zz, c1 = (s1 < P.n_shade) ? P.clip_xy[s1] : zz, c2 = (s2 < P.n_shade) ? P.clip_xy[s2] : zz;
		const float4 p[3] = { make_float4(c0.x, c0.y, p0.z, p0.w), make_float4(c1.x, c1.y, p1.z, p1.w), make_float4(c2.x, c2.y, p2.z, p2.w) };
		const uint32_t sid[3] = { s0, s1, s2 };
		int in_idx[3], out_idx[3], n_in = 0, n_out = 0;
		for (int j = 0; j < 3; j++) { if ((in_mask >> j) & 1u) in_idx[n_in++] = j; else out_idx[n_out++] = j; }
		const uint32_t new0 = P.clip_vid_base + 2u * t, new1 = new0 + 1u;
		float t0, t1;
		if (n_in == 1)
		{
			const int ia = in_idx[0];
			const float4 q1 = near_intersect(p[ia], p[out_idx[0]], t0), q2 = near_intersect(p[ia], p[out_idx[1]], t1);
			fold_prim<FS>(P, x, count_from, to_screen(p[ia], P), to_screen(q1, P), to_screen(q2, P), sid[ia], new0, new1, col, dep, n_tested, n_shaded);
		}
		else
		{
			const int ia = in_idx[0], ib = in_idx[1], io = out_idx[0];
			const float4 q0 = near_intersect(p[ia], p[io], t0), q1 = near_intersect(p[ib], p[io], t1);
			const float4 sq0 = to_screen(q0, P);
			fold_prim<FS>(P, x, count_from, to_screen(p[ia], P), to_screen(p[ib], P), sq0, sid[ia], sid[ib], new0, col, dep, n_tested, n_shaded);
			fold_prim<FS>(P, x, count_from, to_screen(p[ib], P), sq0, to_screen(q1, P), sid[ib], new0, new1, col, dep, n_tested, n_shaded);
		}
	}
	P.color[pix] = col;
	P.depth[pix] = canon_nan(dep);
	if (P.count_fragments && (n_tested | n_shaded))
	{
		atomicAdd(&P.ctr->tested[x % SWGL_CTR_SLOTS], (unsigned long long)n_tested);
		atomicAdd(&P.ctr->shaded[x % SWGL_CTR_SLOTS], (unsigned long long)n_shaded);
	}
}

/* mip_lod: MipMapLevel is a global of the reference that every DrawTriangle call sets first thing (swgl.c:3314-3316) and
 * that GL_POINTS draws then sample with (their path never sets it).  One thread finds the last triangle of the draw
 * that reaches DrawTriangle -- the last input triangle not entirely behind the near plane, and of the two triangles the
 * clipper may make of it the second -- and leaves its level where k_points_write reads it. */
__global__ void k_last_level(const __grid_constant__ DrawParams P)
{
	for (uint32_t t = P.ntri; t-- > 0; )
	{
		float4 p0, p1, p2;
		uint32_t s0, s1, s2;
		tri_vertices(P, t, p0, p1, p2, s0, s1, s2);
		const uint32_t in_mask = (p0.z >= -p0.w ? 1u : 0u) | (p1.z >= -p1.w ? 2u : 0u) | (p2.z >= -p2.w ? 4u : 0u);
		if (in_mask == 0u) continue;
		if (in_mask != 7u)
		{
			const float2 zz = make_float2(0.0f, 0.0f);
			const float2 c0 = (s0 < P.n_shade) ? P.clip_xy[s0] : zz, c1 = (s1 < P.n_shade) ? P.clip_xy[s1] : zz, c2 = (s2 < P.n_shade) ? P.clip_xy[s2] : zz;
			const float4 p[3] = { make_float4(c0.x, c0.y, p0.z, p0.w), make_float4(c1.x, c1.y, p1.z, p1.w), make_float4(c2.x, c2.y, p2.z, p2.w) };
			int in_idx[3], out_idx[3], n_in = 0, n_out = 0;
			for (int j = 0; j < 3; j++) { if ((in_mask >> j) & 1u) in_idx[n_in++] = j; else out_idx[n_out++] = j; }
			float t0, t1;
			if (n_in == 1)
			{
				p0 = to_screen(p[in_idx[0]], P);
				p1 = to_screen(near_intersect(p[in_idx[0]], p[out_idx[0]], t0), P);
				p2 = to_screen(near_intersect(p[in_idx[0]], p[out_idx[1]], t1), P);
			}
			else
			{
				p0 = to_screen(p[in_idx[1]], P);
				p1 = to_screen(near_intersect(p[in_idx[0]], p[out_idx[0]], t0), P);
				p2 = to_screen(near_intersect(p[in_idx[1]], p[out_idx[0]], t1), P);
			}
		}
		*P.last_level = mip_level(p0.x, p0.y, p1.x, p1.y, p2.x, p2.y);
		return;
	}
}

/* raster_path: 1 = pixel-owner CTA per 32x32 tile (k_raster), 2 = fragment-parallel CTA per 32x32
 * tile (k_raster_frag), 3 = warp per 32x8 tile (k_raster_warp).  0 = default = the warp kernel for
 * every draw.  All three produce identical bits. */
/* every draw gets a serial and an event; the allocations it reads remember the serial */
static int stamp_draw(swgldev_ctx* c, const swgldev_draw* d)
{
	c->draw_serial++;
	CK(cudaEventRecord(c->draw_ev[c->draw_serial & 7u], c->stream));
	const swgldev_ptr used[2] = { d->vbo, d->ibo };
	for (int i = 0; i < 2; i++)
	{
		auto it = c->last_use.find((uintptr_t)used[i]);
		if (it != c->last_use.end()) it->second = c->draw_serial;
	}
	for (int u = 0; u < SWGL_MAX_TEX_UNITS; u++)
		if (d->tex[u].data)
		{
			auto it = c->last_use.find((uintptr_t)d->tex[u].data);
			if (it != c->last_use.end()) it->second = c->draw_serial;
		}
	return 0;
}

static int raster_path_for(const swgldev_ctx* c, uint32_t ntri)
{
	if (c->opt_raster_path >= 1 && c->opt_raster_path <= 3) return c->opt_raster_path;
	(void)ntri;
	return 3;       /* the warp kernel wins on every measured scene, big triangles included (C1: 60 vs 100 us) */
}
/* meshes of small triangles have few tall or wide primitives: the set-up kernel inserts them itself */
static bool small_triangle_draw(const swgldev_ctx* c, uint32_t ntri)
{
	return (double)c->W * (double)c->H / (double)(ntri ? ntri : 1u) <= 64.0;
}
/* Tile height of a draw.  The CTA kernels use 32 rows.  The warp kernel uses 8 when the tiles this context owns
 * fill at least one resident wave of warps (148 SMs x 32), else 4, else 2 (a tile is one warp's serial work: with
 * few tiles the slowest warp sets the kernel's time); the "tile_rows" option pins it.  Tile rows must stay below
 * 1024 (band-entry packing). */
static uint32_t th_shift_of(const swgldev_ctx* c, int path)
{
	if (path != 3) return SWGL_TILE_SHIFT;
	if (c->opt_tile_rows == 8 || c->opt_tile_rows == 4 || c->opt_tile_rows == 2)
	{
		const uint32_t sh = c->opt_tile_rows == 8 ? 3u : c->opt_tile_rows == 4 ? 2u : 1u;
		if (((c->H + (1u << sh) - 1u) >> sh) <= 1023u) return sh;
	}
	/* measured (tools/tile_rows_probe.py, r02): C1 (640x480, 1 200 tiles of 8 rows) 79.6 us with 8 rows, 73.5 with 4,
	 * 85.3 with 2 (band entries and list insertions of its tall triangles grow faster than the raster kernel
	 * shrinks); C2 (8 100 tiles) 51.8 / 56.7 / 141.7.  So: 4 rows below one resident wave of 8-row tiles, never 2
	 * unless asked for.
	 * The share of one rank of a sort-first group is a different matter: 4 050 tiles per rank of C4 at N = 8 take
	 * 31.7 us with 8 rows and 32.8 with 4 (tools/stripe_probe.py): ranks keep 8 rows. */
	const size_t tiles8 = (size_t)c->tiles_x * ((c->H + 7u) >> 3);
	if (c->n_ranks <= 1 && tiles8 < 148u * 32u && ((c->H + 3u) >> 2) <= 1023u) return 2u;
	return WT_H_SHIFT;
}

template <int FS>
static void launch_raster(swgldev_ctx* c, const DrawParams& P)
{
	const int path = P.th_shift < SWGL_TILE_SHIFT ? 3 : (c->opt_raster_path == 1 ? 1 : 2);
	c->last_raster_path = path;
	dim3 grid(P.tiles_x, P.tiles_y);
	const dim3 wgrid((P.tiles_x + WT_WARPS - 1) / WT_WARPS, P.owned_tile_rows ? P.owned_tile_rows : 1);
	if (path == 1) k_raster<FS><<<grid, SWGL_RASTER_THREADS, sizeof(RasterShared), c->stream>>>(P);
	else if (path == 2) k_raster_frag<FS><<<grid, FRAG_THREADS, sizeof(FragShared), c->stream>>>(P);
	else if (P.th_shift == 3) k_raster_warp<FS, 3><<<wgrid, WT_WARPS * 32, 0, c->stream>>>(P);
	else if (P.th_shift == 2) k_raster_warp<FS, 2><<<wgrid, WT_WARPS * 32, 0, c->stream>>>(P);
	else k_raster_warp<FS, 1><<<wgrid, WT_WARPS * 32, 0, c->stream>>>(P);
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
	c->shared_mirror = nullptr; c->shared_mirror_dev = nullptr; c->shared_mirror_bytes = 0;
	c->clip = nullptr; c->cap_clip = 0; c->clip_xy = nullptr; c->cap_clip_xy = 0; c->vary = nullptr; c->cap_vary = 0;
	c->prims = nullptr; c->cap_prims = 0; c->bin_cap = 0; c->side = nullptr; c->setup_event = nullptr; c->app_event = nullptr;
	c->bands = nullptr; c->cap_bands = 0; c->pairs = nullptr; c->cap_pairs = 0;
	c->ov_pool = nullptr; c->cap_ov = 0; c->hot_store = nullptr; c->cap_hot = 0;
	c->tile_count = nullptr; c->winner = nullptr; c->ctr = nullptr; c->h_ctr = nullptr; c->ctr_event = nullptr;
	c->ctr_pending = 0; c->last_draw_valid = 0; c->last_raster_path = 0;
	c->opt_fuse_clear = 1; c->opt_count_fragments = 1; c->opt_raster_path = 0; c->opt_stage_timing = 0; c->opt_diag = 0; c->opt_bin_limit = (size_t)6 << 30;
	c->n_launches = 0; c->stage_draws = 0; c->selftest_mismatches = -1; c->opt_lean_prims = 1; c->opt_mip_lod = 0; c->opt_jit = 1; c->opt_overflow_pool = 1; c->opt_setup_pipelined = 0; c->opt_setup_big = 1; c->opt_tile_rows = 0; c->cur_th_shift = WT_H_SHIFT; c->jit_failed = 0; c->last_vs_kind = -1; c->last_fs_kind = -1;
	c->mirror_synced = 0; c->wt_predict = 0; c->draws_since_map = 0; c->opt_host_mirror = 1; c->color_exposed = 0; c->wt_draws = 0;
	c->h_mirror[0] = c->h_mirror[1] = nullptr; c->frame_ev[0] = c->frame_ev[1] = nullptr; c->frame_serial = 0; c->rgba_staging = nullptr; c->copy = nullptr; c->frame_done = nullptr; c->copy_inflight = 0;
	c->upload = nullptr; c->draw_serial = 0; c->d_maxidx = nullptr; c->h_maxidx = nullptr; c->lut255 = nullptr; c->d_last_level = nullptr;
	for (int i = 0; i < 8; i++) c->draw_ev[i] = nullptr;
	for (int i = 0; i < 8; i++) { c->stage_ev[i] = nullptr; c->stage_us[i] = 0.0; }
	memset(&c->pending_clear, 0, sizeof(c->pending_clear));
	memset(&c->stats, 0, sizeof(c->stats));
	c->n_draws = 0; c->draws_refused = 0; c->draws_folded = 0; c->in_fold = 0; c->fold_shift = 0; c->error[0] = 0;
	for (int i = 0; i < SWGL_MAX_GROUP; i++) c->group[i] = nullptr;
	c->n_group = 1; c->solo = 0; c->pool = nullptr; c->group_peer_ok = 0; c->mirror_borrowed = 0; c->slice_ev = nullptr; c->gather_ev = nullptr; c->gather_pending = 0; c->gather = nullptr; c->gather_ring_pos = 0; c->gather_evicted = 0;
	for (int i = 0; i < 4; i++) { c->gather_ring_ev[i] = nullptr; c->gather_ring_dst[i] = 0; }

	const size_t npx = (size_t)width * height;
	const size_t ntiles = (size_t)c->tiles_x * ((height + (1u << WT_H_SHIFT_MIN) - 1) >> WT_H_SHIFT_MIN);   /* finest tiling */
	/* the upload stream's small reduction kernel must not queue behind a whole raster grid */
	int prio_lo = 0, prio_hi = 0;
	cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
	bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess
	       && cudaMalloc((void**)&c->color, (npx ? npx : 1) * 4) == cudaSuccess
	       && cudaMalloc((void**)&c->depth, (npx ? npx : 1) * 4) == cudaSuccess
	       && cudaHostAlloc((void**)&c->h_mirror[0], (npx ? npx : 1) * 4, cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess   /* portable: the members of a device group write into it */
	       && cudaStreamCreateWithPriority(&c->upload, cudaStreamNonBlocking, prio_hi) == cudaSuccess
	       && cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->frame_done, cudaEventDisableTiming) == cudaSuccess
	       && cudaMalloc((void**)&c->d_maxidx, 4) == cudaSuccess
	       && cudaMallocHost((void**)&c->h_maxidx, 4) == cudaSuccess
	       && cudaMalloc((void**)&c->lut255, 256 * sizeof(float)) == cudaSuccess
	       && cudaMalloc((void**)&c->d_last_level, sizeof(float)) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->frame_ev[0], cudaEventDisableTiming) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->frame_ev[1], cudaEventDisableTiming) == cudaSuccess
	       && cudaHostAlloc((void**)&c->h_depth, (npx ? npx : 1) * 4, cudaHostAllocPortable) == cudaSuccess
	       && cudaMalloc((void**)&c->tile_count, (ntiles + 1) * 4) == cudaSuccess
	       && cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->setup_event, cudaEventDisableTiming) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->app_event, cudaEventDisableTiming) == cudaSuccess
	       && cudaMalloc((void**)&c->ctr, sizeof(Counters)) == cudaSuccess
	       && cudaMallocHost((void**)&c->h_ctr, sizeof(Counters)) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->ctr_event, cudaEventDisableTiming) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->slice_ev, cudaEventDisableTiming) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->gather_ev, cudaEventDisableTiming) == cudaSuccess
	       && cudaStreamCreateWithPriority(&c->gather, cudaStreamNonBlocking, prio_hi) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->gather_ring_ev[0], cudaEventDisableTiming) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->gather_ring_ev[1], cudaEventDisableTiming) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->gather_ring_ev[2], cudaEventDisableTiming) == cudaSuccess
	       && cudaEventCreateWithFlags(&c->gather_ring_ev[3], cudaEventDisableTiming) == cudaSuccess
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
	k_init_lut255<<<1, 256, 0, c->stream>>>(c->lut255);
	cudaMemsetAsync(c->d_last_level, 0, sizeof(float), c->stream);     /* the reference's global starts at 0: base level */
	c->h_color = c->h_mirror[0];
	for (int i = 0; i < 8; i++)
		if (cudaEventCreateWithFlags(&c->draw_ev[i], cudaEventDisableTiming) != cudaSuccess) { swgldev_destroy(c); return nullptr; }

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

/* glInit with swglSetDeviceCount(n): n device contexts on ordinals device .. device+n-1 driven by the calling
 * thread (SURVEY.md 7 step 3, 8b "one host thread drives all GPUs").  Sort-first: tile-row bands of 32 rows are
 * dealt round-robin (at least 16 bands per member), geometry and textures are replicated, and every member
 * stores the tiles it finishes into the leader's pinned frame mirror over its own PCIe link. */
swgldev_ctx* swgldev_create_group(int device, int count, uint32_t width, uint32_t height)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return nullptr; }
	if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
	if (count <= 1) return swgldev_create(device, width, height);
	/* SWGL_B200_GROUP_EMULATE=1 (tests on a box with fewer GPUs): members beyond the visible devices wrap around
	 * and share a device -- same code path, same bands, same copies, no speed-up */
	const char* emu = getenv("SWGL_B200_GROUP_EMULATE");
	const bool wrap = emu && emu[0] == '1';
	if (count > SWGL_MAX_GROUP || (!wrap && device + count > n))
	{
		fprintf(stderr, "swgl_b200: %d devices from ordinal %d requested, %d visible\n", count, device, n);
		return nullptr;
	}
	swgldev_ctx* members[SWGL_MAX_GROUP];
	int ord[SWGL_MAX_GROUP];
	/* Which devices: consecutive ordinals from `device`, unless fewer devices are asked for than are visible and
	 * the start is the default -- then every (visible / count)-th device in PCI bus order.  Neighbouring GPUs of an
	 * 8-GPU board share their path to host memory (measured on an HGX B200, C4 end to end: devices {0,1,2,3}
	 * 0.78 ms per frame, {0,2,4,6} 0.55 ms; {0,1} 0.82, {0,4} 0.71), and a group moves its geometry in and its frame
	 * out over every member's own link. */
	int spread[64], n_spread = 0;
	if (!wrap && device == 0 && count < n && n <= 64)
	{
		char bus[64][32];
		for (int i = 0; i < n; i++) { spread[i] = i; if (cudaDeviceGetPCIBusId(bus[i], 32, i) != cudaSuccess) { bus[i][0] = (char)('0' + i); bus[i][1] = 0; } }
		for (int i = 1; i < n; i++)
			for (int j = i; j > 0 && strcmp(bus[spread[j - 1]], bus[spread[j]]) > 0; j--) { const int t = spread[j]; spread[j] = spread[j - 1]; spread[j - 1] = t; }
		n_spread = n;
		cudaGetLastError();
	}
	for (int i = 0; i < count; i++)
	{
		ord[i] = n_spread ? spread[(i * (n_spread / count)) % n_spread] : (device + i) % n;
		members[i] = swgldev_create(ord[i], width, height);
		if (!members[i]) { for (int j = 0; j < i; j++) swgldev_destroy(members[j]); return nullptr; }
	}
	/* NVLink between the replicas: the all-gather of a sharded upload copies device to device */
	int peer_ok = 1;
	for (int i = 0; i < count; i++)
	{
		cudaSetDevice(ord[i]);
		for (int j = 0; j < count; j++)
			if (ord[j] != ord[i])
			{
				int can = 0;
				cudaError_t e = cudaDeviceCanAccessPeer(&can, ord[i], ord[j]);
				if (e == cudaSuccess && can)
				{
					e = cudaDeviceEnablePeerAccess(ord[j], 0);
					if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
				}
				else can = 0;
				if (!can) peer_ok = 0;      /* uploads fall back to cudaMemcpyPeerAsync (staged through the host if need be) */
				cudaGetLastError();
			}
	}
	members[0]->group_peer_ok = peer_ok;
	swgldev_ctx* c = members[0];
	const uint32_t tiles_y32 = (height + 31u) / 32u;
	uint32_t band = tiles_y32 / ((uint32_t)count * 16u);
	if (band == 0) band = 1;
	for (int i = 0; i < count; i++)
	{
		swgldev_ctx* m = members[i];
		c->group[i] = m;
		m->rank = (uint32_t)i; m->n_ranks = (uint32_t)count; m->band_rows = band;
		/* the leader's pinned mirror is every member's write-through target (page-locked, portable, mapped) */
		m->shared_mirror = c->h_mirror[0]; m->shared_mirror_dev = c->h_mirror[0];
		m->shared_mirror_bytes = (size_t)width * height * 4; m->mirror_borrowed = 1; m->mirror_synced = 0;
	}
	c->n_group = count;
	c->pool = new GroupPool();
	c->pool->start(count, ord);
	cudaSetDevice(c->device);
	return c;
}

void swgldev_destroy(swgldev_ctx* c)
{
	if (!c) return;
	if (c->n_group > 1)
	{
		/* every member drains before the leader's pinned mirror (which they write into) goes */
		for (int i = 0; i < c->n_group; i++) if (c->group[i] && c->group[i]->stream) { cudaSetDevice(c->group[i]->device); cudaStreamSynchronize(c->group[i]->stream); cudaStreamSynchronize(c->group[i]->upload); }
		if (c->pool) { c->pool->stop(); delete c->pool; c->pool = nullptr; }
		for (int i = 1; i < c->n_group; i++) { swgldev_destroy(c->group[i]); c->group[i] = nullptr; }
		c->n_group = 1;
	}
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->shared_mirror && !c->mirror_borrowed) cudaHostUnregister(c->shared_mirror);
	c->shared_mirror = nullptr;
	for (void* p : c->allocations) cudaFree(p);
	for (auto& kv : c->code_cache) cudaFree(kv.second);
	if (c->upload) { cudaStreamSynchronize(c->upload); cudaStreamDestroy(c->upload); }
	if (c->copy) { cudaStreamSynchronize(c->copy); cudaStreamDestroy(c->copy); }
	if (c->frame_done) cudaEventDestroy(c->frame_done);
	cudaFree(c->color); cudaFree(c->depth); cudaFreeHost(c->h_mirror[0]); cudaFreeHost(c->h_depth);
	if (c->h_mirror[1]) cudaFreeHost(c->h_mirror[1]);
	if (c->rgba_staging) cudaFree(c->rgba_staging);
	if (c->d_maxidx) cudaFree(c->d_maxidx);
	if (c->h_maxidx) cudaFreeHost(c->h_maxidx);
	if (c->lut255) cudaFree(c->lut255);
	if (c->d_last_level) cudaFree(c->d_last_level);
	for (int i = 0; i < 2; i++) if (c->frame_ev[i]) cudaEventDestroy(c->frame_ev[i]);
	for (int i = 0; i < 8; i++) if (c->draw_ev[i]) cudaEventDestroy(c->draw_ev[i]);
	cudaFree(c->tile_count); cudaFree(c->ctr); cudaFreeHost(c->h_ctr);
	if (c->winner) cudaFree(c->winner);
	if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
	if (c->setup_event) cudaEventDestroy(c->setup_event);
	if (c->app_event) cudaEventDestroy(c->app_event);
	if (c->clip) cudaFree(c->clip);
	if (c->clip_xy) cudaFree(c->clip_xy);
	if (c->vary) cudaFree(c->vary);
	if (c->prims) cudaFree(c->prims);
	if (c->bands) cudaFree(c->bands);
	if (c->pairs) cudaFree(c->pairs);
	if (c->ov_pool) cudaFree(c->ov_pool);
	if (c->hot_store) cudaFree(c->hot_store);
	if (c->ctr_event) cudaEventDestroy(c->ctr_event);
	if (c->slice_ev) cudaEventDestroy(c->slice_ev);
	if (c->gather_ev) cudaEventDestroy(c->gather_ev);
	if (c->gather) { cudaStreamSynchronize(c->gather); cudaStreamDestroy(c->gather); }
	for (int i = 0; i < 4; i++) if (c->gather_ring_ev[i]) cudaEventDestroy(c->gather_ring_ev[i]);
	for (int i = 0; i < 8; i++) if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
	if (c->stream) cudaStreamDestroy(c->stream);
	cudaGetLastError();
	delete c;
}

const char* swgldev_last_error(swgldev_ctx* c)
{
	static char out[512];
	out[0] = 0;
	if (IS_GROUP(c))
	{   /* the first member that has something to say; all are cleared */
		for (int i = 0; i < c->n_group; i++)
		{
			swgldev_ctx* m = c->group[i];
			if (!out[0] && m->error[0]) { if (i) snprintf(out, sizeof(out), "device %d: %.480s", m->device, m->error); else snprintf(out, sizeof(out), "%s", m->error); }
			m->error[0] = 0;
		}
		return out;
	}
	snprintf(out, sizeof(out), "%s", c->error);
	c->error[0] = 0;
	return out;
}

void* swgldev_stream(swgldev_ctx* c) { return (void*)c->stream; }

/* the replica of a leader allocation on member i (the leader's own pointer for i = 0 or an address it does not know) */
static swgldev_ptr member_ptr(swgldev_ctx* leader, int i, swgldev_ptr p)
{
	if (!p) return 0;
	auto it = leader->replicas.find((uintptr_t)p);
	if (it == leader->replicas.end()) return i == 0 ? p : 0;
	return it->second[(size_t)i];
}

swgldev_ptr swgldev_alloc(swgldev_ctx* c, uint64_t bytes)
{
	if (IS_GROUP(c))
	{
		std::vector<swgldev_ptr> reps((size_t)c->n_group, 0);
		group_each(c, [&](int i, swgldev_ctx* m) { reps[(size_t)i] = swgldev_alloc(m, bytes); return 0; });
		for (int i = 0; i < c->n_group; i++)
			if (!reps[(size_t)i])
			{
				group_each(c, [&](int j, swgldev_ctx* m) { if (reps[(size_t)j]) swgldev_free(m, reps[(size_t)j]); return 0; });
				return 0;
			}
		c->replicas[(uintptr_t)reps[0]] = reps;
		return reps[0];
	}
	void* p = nullptr;
	cudaSetDevice(c->device);
	cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
	if (e != cudaSuccess) { set_err(c, "cudaMalloc (buffer/texture upload)", e); return 0; }
	c->allocations.push_back(p);
	c->last_use[(uintptr_t)p] = 0;
	return (swgldev_ptr)(uintptr_t)p;
}

static int settle_last_draw(swgldev_ctx* c);

void swgldev_free(swgldev_ctx* c, swgldev_ptr p)
{
	if (IS_GROUP(c))
	{
		auto it = c->replicas.find((uintptr_t)p);
		if (it == c->replicas.end()) return;
		const std::vector<swgldev_ptr> reps = it->second;
		c->replicas.erase(it);
		swgldev_sync(c);      /* every member: a gather on another member may still be reading this member's replica */
		group_each(c, [&](int i, swgldev_ctx* m) { swgldev_free(m, reps[(size_t)i]); return 0; });
		return;
	}
	void* q = (void*)(uintptr_t)p;
	/* an overflowed draw is re-issued from the raw pointers it was launched with: resolve it while
	 * the memory it reads is still there */
	cudaSetDevice(c->device);
	settle_last_draw(c);
	for (size_t i = 0; i < c->allocations.size(); i++)
		if (c->allocations[i] == q)
		{
			c->allocations[i] = c->allocations.back();
			c->allocations.pop_back();
			c->last_use.erase((uintptr_t)q);
			cudaStreamSynchronize(c->stream); /* queued draws may still read it */
			cudaFree(q);
			return;
		}
}

static int group_upload(swgldev_ctx* c, swgldev_ptr dst, uint64_t offset, const void* src, uint64_t bytes, uint32_t* max_index);

static void queue_max_index(swgldev_ctx* c, swgldev_ptr indices, uint64_t bytes);

/* Device group: `bytes` of host memory into every member's replica of `dst` (+ offset).  Member i copies
 * slice i over ITS PCIe link, then one kernel per member pulls the other slices straight out of the peers'
 * replicas over NVLink (k_group_gather, peer loads; cudaMemcpyPeerAsync per slice where peer access is
 * missing).  The call returns when the host memory has been read (the caller may free it, swgl.c:3142); the
 * device-to-device part only holds back the members' own draw streams.  With `max_index` each member also
 * reduces the largest u32 of its slice (element data). */
static int group_upload(swgldev_ctx* c, swgldev_ptr dst, uint64_t offset, const void* src, uint64_t bytes, uint32_t* max_index)
{
	const int n = c->n_group;
	if (max_index) *max_index = 0;
	if (bytes == 0) return 0;
	uint64_t slice = (bytes + (uint64_t)n - 1) / (uint64_t)n;
	slice = (slice + 255u) & ~(uint64_t)255u;
	for (int i = 0; i < n; i++) if (!member_ptr(c, i, dst)) { set_err(c, "upload: not a buffer of this device group", cudaSuccess); return -1; }
	/* 1. hazards (a replica may still be read by its member's queued draws, or by another member's pull of the
	 * previous upload), then every member's own slice from the host */
	int rc = group_each(c, [&](int i, swgldev_ctx* m)
	{
		int r = 0;
		if (settle_last_draw(m)) r = -1;
		char* rep = (char*)(uintptr_t)member_ptr(c, i, dst);
		auto it = m->last_use.find((uintptr_t)rep);
		if (it != m->last_use.end() && it->second)
		{
			const uint64_t sr = (m->draw_serial - it->second < 8u) ? it->second : m->draw_serial;
			if (cudaStreamWaitEvent(m->upload, m->draw_ev[sr & 7u], 0) != cudaSuccess) r = -1;
		}
		/* the previous upload INTO THIS BUFFER may still be pulled from (or into) the replicas */
		for (int j = 0; j < n; j++)
		{
			const swgldev_ctx* o = c->group[j];
			if (!o->gather_pending) continue;
			if (o->gather_evicted) { cudaStreamWaitEvent(m->upload, o->gather_ev, 0); continue; }
			for (int q = 0; q < 4; q++) if (o->gather_ring_dst[q] == dst) cudaStreamWaitEvent(m->upload, o->gather_ring_ev[q], 0);
		}
		const uint64_t lo = (uint64_t)i * slice, hi = lo + slice < bytes ? lo + slice : bytes;
		if (lo < hi)
		{
			if (cudaMemcpyAsync(rep + offset + lo, (const char*)src + lo, hi - lo, cudaMemcpyHostToDevice, m->upload) != cudaSuccess) r = -1;
			if (max_index) queue_max_index(m, (swgldev_ptr)(uintptr_t)(rep + offset + lo), hi - lo);
		}
		else if (max_index) *m->h_maxidx = 0;
		if (cudaEventRecord(m->slice_ev, m->upload) != cudaSuccess) r = -1;
		return r;
	});
	/* 2. (every slice event has been recorded) the other slices, device to device; then the host waits for the
	 * member's own slice only */
	GatherArgs ga;
	memset(&ga, 0, sizeof(ga));
	for (int j = 0; j < n; j++) ga.src[j] = (const char*)(uintptr_t)member_ptr(c, j, dst) + offset;
	ga.slice = slice; ga.bytes = bytes; ga.n = n;
	const bool by_kernel = c->group_peer_ok && ((offset & 15u) == 0);
	std::atomic<uint32_t> mx{0};
	rc |= group_each(c, [&](int i, swgldev_ctx* m)
	{
		int r = 0;
		char* rep = (char*)(uintptr_t)member_ptr(c, i, dst);
		for (int k = 0; k < n; k++)
		{
			const int j = (i + k) % n;                 /* k = 0: the member's own slice and whatever the upload stream held before it */
			if (((uint64_t)j * slice < bytes || k == 0) && cudaStreamWaitEvent(m->gather, c->group[j]->slice_ev, 0) != cudaSuccess) r = -1;
		}
		if (by_kernel)
		{
			GatherArgs a = ga;
			a.dst = rep + offset; a.self = i;
			k_group_gather<<<148 * 2, 256, 0, m->gather>>>(a);
			m->n_launches++;
		}
		else
			for (int k = 1; k < n; k++)
			{
				const int j = (i + k) % n;                 /* staggered: at any time every member is pulled from once */
				const uint64_t lo = (uint64_t)j * slice, hi = lo + slice < bytes ? lo + slice : bytes;
				if (lo >= hi) continue;
				if (cudaMemcpyPeerAsync(rep + offset + lo, m->device, ga.src[j] + lo, c->group[j]->device, hi - lo, m->gather) != cudaSuccess) r = -1;
			}
		if (cudaEventRecord(m->gather_ev, m->gather) != cudaSuccess) r = -1;       /* the newest gather of this member */
		{
			int q = -1;
			for (int t = 0; t < 4; t++) if (m->gather_ring_dst[t] == dst) q = t;   /* the same allocation again: its entry moves forward */
			if (q < 0)
			{
				q = m->gather_ring_pos; m->gather_ring_pos = (q + 1) & 3;
				if (m->gather_pending && m->gather_ring_dst[q]) m->gather_evicted = 1;
			}
			m->gather_ring_dst[q] = dst;
			if (cudaEventRecord(m->gather_ring_ev[q], m->gather) != cudaSuccess) r = -1;
		}
		m->gather_pending = 1;
		if (cudaStreamWaitEvent(m->stream, m->gather_ev, 0) != cudaSuccess) r = -1;      /* the member's draws read the whole replica */
		if (cudaEventSynchronize(m->slice_ev) != cudaSuccess) r = -1;                      /* the host memory has been read */
		if (max_index)
		{
			uint32_t cur = mx.load();
			const uint32_t v = *m->h_maxidx;
			while (v > cur && !mx.compare_exchange_weak(cur, v)) { }
		}
		return r;
	});
	if (max_index) *max_index = mx.load();
	if (rc) set_err(c, "upload to the device group failed", cudaGetLastError());
	return rc;
}

int swgldev_upload(swgldev_ctx* c, swgldev_ptr dst, const void* src, uint64_t bytes)
{
	if (IS_GROUP(c))
	{
		return group_each(c, [&](int i, swgldev_ctx* m) { return swgldev_upload(m, member_ptr(c, i, dst), src, bytes); });
	}
	/* ... and before the contents it reads are replaced */
	cudaSetDevice(c->device);
	if (settle_last_draw(c)) return -1;
	/* the caller may free `src` on return (swgl.c:3142-3144), so the copy completes here;
	 * pinned sources go at full PCIe rate, pageable ones through the driver's staging */
	CK(cudaMemcpyAsync((void*)(uintptr_t)dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

/* Same contract as swgldev_upload, but the copy only waits for the draws that read the destination
 * allocation: it overlaps whatever else is queued (the frame being rasterised, for instance). */
int swgldev_upload_overlapped(swgldev_ctx* c, swgldev_ptr dst, const void* src, uint64_t bytes)
{
	if (IS_GROUP(c)) return group_upload(c, dst, 0, src, bytes, nullptr);
	cudaSetDevice(c->device);
	/* an overflowed draw is re-issued from its buffers: resolve it before they may change */
	if (settle_last_draw(c)) return -1;
	auto it = c->last_use.find((uintptr_t)dst);
	if (it == c->last_use.end()) return swgldev_upload(c, dst, src, bytes);   /* not an allocation base: plain ordering */
	const uint64_t last_use = it->second;
	if (last_use)
	{
		const uint64_t s = (c->draw_serial - last_use < 8u) ? last_use : c->draw_serial;   /* older than the ring: newest */
		CK(cudaStreamWaitEvent(c->upload, c->draw_ev[s & 7u], 0));
	}
	CK(cudaMemcpyAsync((void*)(uintptr_t)dst, src, bytes, cudaMemcpyHostToDevice, c->upload));
	CK(cudaStreamSynchronize(c->upload));
	return 0;
}

/* queue the max-index reduction and its read-back behind whatever the upload stream holds */
static void queue_max_index(swgldev_ctx* c, swgldev_ptr indices, uint64_t bytes)
{
	const size_t n = bytes / 4u;
	cudaMemsetAsync(c->d_maxidx, 0, 4, c->upload);
	if (n) k_max_index<<<148 * 4, 256, 0, c->upload>>>((const uint32_t*)(uintptr_t)indices, n, c->d_maxidx);
	cudaMemcpyAsync(c->h_maxidx, c->d_maxidx, 4, cudaMemcpyDeviceToHost, c->upload);
}

void* swgldev_host_alloc(uint64_t bytes, int write_combined)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess)
	{
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

void swgldev_host_free(void* p) { cudaFreeHost(p); }

uint32_t swgldev_max_index(swgldev_ctx* c, swgldev_ptr indices, uint64_t bytes)
{
	if (IS_GROUP(c)) { Solo s(c); return swgldev_max_index(c, indices, bytes); }   /* the replicas hold the same data */
	cudaSetDevice(c->device);
	if (bytes < 4u) return 0;
	/* on the upload stream: the buffer was just written there and nothing later has been queued */
	queue_max_index(c, indices, bytes);
	cudaStreamSynchronize(c->upload);
	return *c->h_maxidx;
}

/* Part of an allocation: `base` names it (for the read hazards), the copy goes to base + offset. */
int swgldev_upload_range(swgldev_ctx* c, swgldev_ptr base, uint64_t offset, const void* src, uint64_t bytes)
{
	if (IS_GROUP(c)) return group_upload(c, base, offset, src, bytes, nullptr);
	cudaSetDevice(c->device);
	if (settle_last_draw(c)) return -1;
	auto it = c->last_use.find((uintptr_t)base);
	if (it == c->last_use.end()) return swgldev_upload(c, base + offset, src, bytes);
	const uint64_t last_use = it->second;
	if (last_use)
	{
		const uint64_t s = (c->draw_serial - last_use < 8u) ? last_use : c->draw_serial;
		CK(cudaStreamWaitEvent(c->upload, c->draw_ev[s & 7u], 0));
	}
	CK(cudaMemcpyAsync((void*)(uintptr_t)(base + offset), src, bytes, cudaMemcpyHostToDevice, c->upload));
	CK(cudaStreamSynchronize(c->upload));
	return 0;
}

/* Largest index of element data that work queued on the library's stream (an all-gather of the
 * application's, say) has just written: the reduction runs behind that work. */
uint32_t swgldev_max_index_after_stream(swgldev_ctx* c, swgldev_ptr indices, uint64_t bytes)
{
	if (IS_GROUP(c)) { Solo s(c); return swgldev_max_index_after_stream(c, indices, bytes); }
	cudaSetDevice(c->device);
	if (bytes < 4u) return 0;
	if (cudaEventRecord(c->app_event, c->stream) != cudaSuccess) return 0;
	cudaStreamWaitEvent(c->upload, c->app_event, 0);
	queue_max_index(c, indices, bytes);
	cudaStreamSynchronize(c->upload);
	return *c->h_maxidx;
}

/* swgldev_upload_overlapped of element data plus the largest index in it, with one wait for both */
int swgldev_upload_indices(swgldev_ctx* c, swgldev_ptr dst, const void* src, uint64_t bytes, uint32_t* max_index)
{
	if (IS_GROUP(c)) return group_upload(c, dst, 0, src, bytes, max_index);
	cudaSetDevice(c->device);
	*max_index = 0;
	if (settle_last_draw(c)) return -1;
	auto it = c->last_use.find((uintptr_t)dst);
	if (it == c->last_use.end())
	{
		if (swgldev_upload(c, dst, src, bytes)) return -1;
		*max_index = swgldev_max_index(c, dst, bytes);
		return 0;
	}
	const uint64_t last_use = it->second;
	if (last_use)
	{
		const uint64_t s = (c->draw_serial - last_use < 8u) ? last_use : c->draw_serial;   /* older than the ring: newest */
		CK(cudaStreamWaitEvent(c->upload, c->draw_ev[s & 7u], 0));
	}
	CK(cudaMemcpyAsync((void*)(uintptr_t)dst, src, bytes, cudaMemcpyHostToDevice, c->upload));
	queue_max_index(c, dst, bytes);
	CK(cudaStreamSynchronize(c->upload));
	*max_index = bytes >= 4u ? *c->h_maxidx : 0u;
	return 0;
}

/* ---- deferred overflow check: wait for the counter snapshot of the previous draw and, when
 * its scratch was too small, grow and re-issue it before anything later is queued ---- */
static int launch_draw(swgldev_ctx* c, DrawParams& P);

/* per-tile lists may take up to opt_bin_limit bytes (default 6 GiB); beyond that draws are split */

/* Snapshot of the previous draw's counters (taken right after its set-up kernel, on the side
 * stream).  If its binning scratch was too small every kernel after set-up has exited without
 * touching the framebuffer; grow and run it again before anything later is queued.  A draw whose
 * lists cannot fit at all is split in two halves by triangles -- consecutive sub-draws keep the
 * per-pixel submission order, so the result is the same. */
static int resolve_and_reissue(swgldev_ctx* c, const DrawParams& P, int depth);

static int settle_last_draw(swgldev_ctx* c)
{
	if (!c->ctr_pending) return 0;
	CK(cudaEventSynchronize(c->ctr_event));
	c->ctr_pending = 0;
	if (!c->h_ctr->overflow || !c->last_draw_valid) return 0;
	const DrawParams P = c->last_draw;
	c->last_draw_valid = 0;
	return resolve_and_reissue(c, P, 0);
}

/* Launch `P` and wait for it; resolve overflows until it has been drawn. */
static int issue_sync(swgldev_ctx* c, DrawParams P, int depth)
{
	if (launch_draw(c, P)) return -1;
	c->ctr_pending = 0;
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaStreamSynchronize(c->side));
	if (!c->h_ctr->overflow) return 0;
	return resolve_and_reissue(c, P, depth + 1);
}

/* pool of `entries` (tile, entry) pairs + room for SWGL_MAX_HOT_TILES whole lists in hot_store */
static int ensure_overflow_pool(swgldev_ctx* c, size_t entries)
{
	if (entries > 0xfffffff0ull) entries = 0xfffffff0ull;
	if (entries > c->cap_ov)
	{
		uint2* np = nullptr;
		CK(cudaMallocAsync((void**)&np, entries * sizeof(uint2), c->stream));
		if (c->ov_pool) CK(cudaFreeAsync(c->ov_pool, c->stream));
		c->ov_pool = np; c->cap_ov = entries;
	}
	size_t hot = c->cap_ov + (size_t)SWGL_MAX_HOT_TILES * c->bin_cap;
	if (hot > 0xfffffff0ull) hot = 0xfffffff0ull;
	if (hot > c->cap_hot)
	{
		uint32_t* np = nullptr;
		CK(cudaMallocAsync((void**)&np, hot * sizeof(uint32_t), c->stream));
		if (c->hot_store) CK(cudaFreeAsync(c->hot_store, c->stream));
		c->hot_store = np; c->cap_hot = hot;
	}
	return 0;
}

/* Precondition: the last attempt at `P` was dropped by the device (overflow flag set). */
static int resolve_and_reissue(swgldev_ctx* c, const DrawParams& P, int depth)
{
	if (depth > 64) { set_err(c, "draw dropped: binning scratch could not be sized", cudaSuccess); return -1; }
	CK(cudaStreamSynchronize(c->stream));
	Counters h;
	CK(cudaMemcpy(&h, c->ctr, sizeof(h), cudaMemcpyDeviceToHost));
	if (h.overflow & 2u)
		if (grow(c, &c->bands, &c->cap_bands, (size_t)h.band_cursor + (size_t)h.band_cursor / 4 + 1024)) return -1;
	if ((h.overflow & 1u) && P.ov_cap && h.hot_tiles <= SWGL_MAX_HOT_TILES && (size_t)h.ov_cursor * 12u <= c->opt_bin_limit)
	{
		/* a few deep tiles: the pool was too small, every other list is fine */
		if (ensure_overflow_pool(c, (size_t)h.ov_cursor + (size_t)h.ov_cursor / 4 + 1024)) return -1;
	}
	else if (h.overflow & 1u)
	{
		const size_t ntiles = (size_t)c->tiles_x * ((c->H + (1u << P.th_shift) - 1) >> P.th_shift);   /* tiles of this draw */
		size_t want = (size_t)h.max_list + (size_t)h.max_list / 4 + 64;
		if (want < 2 * (size_t)c->bin_cap) want = 2 * (size_t)c->bin_cap;
		if (want * ntiles * 4 > c->opt_bin_limit)
		{
			if (P.ntri < 2) { set_err(c, "draw dropped: one primitive list exceeds the bin memory limit", cudaSuccess); return -1; }
			/* split by triangles: first half (keeps the fused clear), then second half */
			DrawParams A = P, B = P;
			A.ntri = P.ntri / 2; A.count = 3u * A.ntri;
			B.ntri = P.ntri - A.ntri; B.count = P.count - A.count; B.first = P.first + (int32_t)A.count;
			B.clear.flags = 0;
			if (!P.ibo) { A.n_shade = 3u * A.ntri; B.n_shade = 3u * B.ntri; A.clip_vid_base = A.n_shade; B.clip_vid_base = B.n_shade; }
			if (issue_sync(c, A, depth + 1)) return -1;
			return issue_sync(c, B, depth + 1);
		}
		if (grow(c, &c->pairs, &c->cap_pairs, want * ntiles)) return -1;
		c->bin_cap = (uint32_t)(c->cap_pairs / ntiles);
		if (ensure_overflow_pool(c, c->cap_ov)) return -1;      /* hot_store holds SWGL_MAX_HOT_TILES lists of the new K */
	}
	return issue_sync(c, P, depth);
}

/* a frame copy of swgldev_frame_submit may still be reading the colour attachment: kernels that
 * write it queue behind that copy (everything before them -- uploads, vertex stage, set-up -- does not) */
static int guard_color_write(swgldev_ctx* c)
{
	if (!c->copy_inflight) return 0;
	CK(cudaStreamWaitEvent(c->stream, c->frame_ev[c->copy_inflight - 1], 0));
	c->copy_inflight = 0;
	return 0;
}

static int flush_clear(swgldev_ctx* c)
{
	ClearParams cp = c->pending_clear;
	if (!cp.flags) return 0;
	c->pending_clear.flags = 0;
	if (cp.x1 <= cp.x0 || cp.y1 <= cp.y0) return 0;
	dim3 block(128), grid(((uint32_t)(cp.x1 - cp.x0) + 511u) / 512u, (uint32_t)(cp.y1 - cp.y0));
	if (guard_color_write(c)) return -1;
	/* the assembled frame of a sort-first group lives in a peer's attachment (peer_color): the clear of this
	 * rank's rows goes there too.  A host mirror (shared or own) is brought up to date by the next read-back. */
	k_clear<<<grid, block, 0, c->stream>>>(c->color, c->depth, c->n_ranks > 1 ? c->peer_color : nullptr, c->W, cp,
	                                       c->rank, c->n_ranks, c->band_rows ? c->band_rows : 1u);
	c->mirror_synced = 0;
	c->n_launches++;
	CK(cudaGetLastError());
	return 0;
}

static int flush_owned_bands(swgldev_ctx* c);

int swgldev_sync(swgldev_ctx* c)
{
	if (IS_GROUP(c))
	{
		return group_each(c, [&](int, swgldev_ctx* m) { return swgldev_sync(m); });
	}
	cudaSetDevice(c->device);
	if (settle_last_draw(c)) return -1;
	if (flush_clear(c)) return -1;
	CK(cudaStreamSynchronize(c->stream));
	if (c->gather_pending)
	{   /* the stream waited for the member's last group upload: no gather is in flight any more */
		c->gather_pending = 0; c->gather_evicted = 0;
		for (int i = 0; i < 4; i++) c->gather_ring_dst[i] = 0;
	}
	if (c->copy_inflight) CK(cudaStreamSynchronize(c->copy));
	if (c->shared_mirror && !c->mirror_synced)
	{
		/* a handed-out device pointer (color_exposed) only means the copy has to be repeated every time:
		 * the attachment may change behind the library's back */
		if (flush_owned_bands(c)) return -1;
		c->mirror_synced = c->color_exposed ? 0 : 1;
	}
	return 0;
}

swgldev_ptr swgldev_build_mipmaps(swgldev_ctx* c, const swgldev_texture* base, int32_t* n_levels)
{
	if (IS_GROUP(c))
	{
		std::vector<swgldev_ptr> reps((size_t)c->n_group, 0);
		std::vector<int32_t> levels((size_t)c->n_group, 0);
		group_each(c, [&](int i, swgldev_ctx* m)
		{
			swgldev_texture b = *base;
			b.data = member_ptr(c, i, base->data);
			b.mips = member_ptr(c, i, base->mips);
			reps[(size_t)i] = swgldev_build_mipmaps(m, &b, &levels[(size_t)i]);
			return 0;
		});
		bool ok = true;
		for (int i = 0; i < c->n_group; i++) ok = ok && reps[(size_t)i];
		*n_levels = levels[0];
		if (!ok)
		{
			group_each(c, [&](int i, swgldev_ctx* m) { if (reps[(size_t)i]) swgldev_free(m, reps[(size_t)i]); return 0; });
			*n_levels = 0;
			return 0;
		}
		c->replicas[(uintptr_t)reps[0]] = reps;
		return reps[0];
	}
	cudaSetDevice(c->device);
	*n_levels = 0;
	if (!base->data || base->fpp < 1 || base->fpp > 4) return 0;
	/* The reference never clears a texture's MipMaps vector: a second glGenerateMipmap APPENDS the levels of the
	 * current image behind the ones that are there, and glTexImage2D leaves them in place (swgl.c:2094-2118,
	 * 2134-2166).  `base->mips` is that earlier chain (0 = none): its levels come first, unchanged. */
	uint32_t hdr[SWGL_MIP_HEADER_WORDS];
	memset(hdr, 0, sizeof(hdr));
	int n_prev = 0;
	size_t prev_floats = 0;
	if (base->mips && base->n_mips > 0)
	{
		if (settle_last_draw(c)) return 0;
		if (cudaStreamSynchronize(c->stream) != cudaSuccess
		    || cudaMemcpy(hdr, (const void*)(uintptr_t)base->mips, sizeof(hdr), cudaMemcpyDeviceToHost) != cudaSuccess)
		{
			set_err(c, "glGenerateMipmap: reading the earlier chain", cudaGetLastError());
			return 0;
		}
		n_prev = base->n_mips < SWGL_MIP_MAX_LEVELS ? base->n_mips : SWGL_MIP_MAX_LEVELS;
		for (int k = 0; k < n_prev; k++)
		{
			const size_t end_k = (size_t)hdr[k] + (size_t)hdr[SWGL_MIP_MAX_LEVELS + k] * hdr[2 * SWGL_MIP_MAX_LEVELS + k] * hdr[3 * SWGL_MIP_MAX_LEVELS + k];
			if (end_k > prev_floats) prev_floats = end_k;
		}
	}
	/* levels while CurWidth + CurHeight > 4, halving with integer division (swgl.c:2129-2135, 2167-2168) */
	int cw = base->width / 2, ch = base->height / 2, n = n_prev;
	size_t floats = prev_floats;
	while (cw + ch > 4 && n < SWGL_MIP_MAX_LEVELS)
	{
		if (floats > 0xffffffffull) { set_err(c, "glGenerateMipmap: chain too large", cudaSuccess); return 0; }
		hdr[n] = (uint32_t)floats;
		hdr[SWGL_MIP_MAX_LEVELS + n] = (uint32_t)(cw > 0 ? cw : 0);
		hdr[2 * SWGL_MIP_MAX_LEVELS + n] = (uint32_t)(ch > 0 ? ch : 0);
		hdr[3 * SWGL_MIP_MAX_LEVELS + n] = (uint32_t)base->fpp;
		floats += (size_t)(cw > 0 ? cw : 0) * (size_t)(ch > 0 ? ch : 0) * (size_t)base->fpp;
		n++;
		cw /= 2; ch /= 2;
	}
	if (n == n_prev) return 0;           /* nothing new (image too small, or the level table is full): the earlier chain stays */
	const swgldev_ptr chain = swgldev_alloc(c, sizeof(hdr) + floats * 4 + 4);
	if (!chain) return 0;
	float* data = (float*)(uintptr_t)(chain + sizeof(hdr));
	bool ok = cudaMemcpyAsync((void*)(uintptr_t)chain, hdr, sizeof(hdr), cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
	if (ok && prev_floats)
		ok = cudaMemcpyAsync(data, (const void*)(uintptr_t)(base->mips + sizeof(hdr)), prev_floats * 4, cudaMemcpyDeviceToDevice, c->stream) == cudaSuccess;
	if (!ok)
	{
		set_err(c, "glGenerateMipmap: upload of the level table", cudaGetLastError());
		swgldev_free(c, chain);
		return 0;
	}
	const void* prev = (const void*)(uintptr_t)base->data;
	int prev_u8 = base->is_float ? 0 : 1;
	for (int k = n_prev; k < n; k++)
	{
		const int lw = (int)hdr[SWGL_MIP_MAX_LEVELS + k], lh = (int)hdr[2 * SWGL_MIP_MAX_LEVELS + k];
		if (lw > 0 && lh > 0)
		{
			k_mipmap_box<<<dim3(((uint32_t)lw + 255u) / 256u, (uint32_t)lh), 256, 0, c->stream>>>(prev, prev_u8, c->lut255, base->fpp, lw, lh, data + hdr[k]);
			c->n_launches++;
		}
		prev = data + hdr[k]; prev_u8 = 0;
	}
	/* hdr[] is on this stack: the copy must have read it before returning */
	if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
	{
		set_err(c, "glGenerateMipmap: k_mipmap_box", cudaGetLastError());
		swgldev_free(c, chain);
		return 0;
	}
	*n_levels = n;
	return chain;
}

int swgldev_clear(swgldev_ctx* c, uint32_t flags, uint32_t color_word, int32_t x0, int32_t y0, int32_t x1, int32_t y1)
{
	if (IS_GROUP(c))
	{
		return group_each(c, [&](int, swgldev_ctx* m) { return swgldev_clear(m, flags, color_word, x0, y0, x1, y1); });
	}
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
	if (IS_GROUP(c))
	{
		group_each(c, [&](int, swgldev_ctx* m) { swgldev_fill(m, color_word, depth); return 0; });
		return;
	}
	cudaSetDevice(c->device);
	settle_last_draw(c);
	c->pending_clear.flags = 0;
	size_t n = (size_t)c->W * c->H;
	guard_color_write(c);
	k_fill_fb<<<1184, 256, 0, c->stream>>>(c->color, c->depth, n, color_word, depth);
	c->mirror_synced = 0;
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

static int launch_vertex(swgldev_ctx* c, DrawParams& P, uint32_t blocks)
{
	if (P.vs_kind == SWVS_JIT)
	{
		void* args[] = { (void*)&P };
		CK(cudaLaunchKernel((const void*)P.jit_vertex, dim3(blocks), dim3(256), args, 0, c->stream));
	}
	else if (P.vs_kind == SWVS_PASS) k_vertex<SWVS_PASS><<<blocks, 256, 0, c->stream>>>(P);
	else if (P.vs_kind == SWVS_MATRIX) k_vertex<SWVS_MATRIX><<<blocks, 256, 0, c->stream>>>(P);
	else k_vertex<SWVS_GENERIC><<<blocks, 256, 0, c->stream>>>(P);
	return 0;
}

/* Generic shaders: swap the on-device interpreter for kernels compiled from the program's IR
 * (swgl_jit.cpp).  A failure is reported once through the error string and the interpreter draws. */
static void apply_jit(swgldev_ctx* c, const swgldev_draw* d, DrawParams& P, bool raster_ok)
{
	const int want_v = P.vs_kind == SWVS_GENERIC, want_r = raster_ok && P.fs_kind == SWFS_GENERIC;
	if (!c->opt_jit || c->jit_failed || (!want_v && !want_r)) return;
	swgljit_kernels k;
	char msg[2048];
	if (swgljit_get(c->device, d, want_v, want_r, &k, msg, sizeof(msg)))
	{
		c->jit_failed = 1;
		char full[2300];
		snprintf(full, sizeof(full), "run-time shader compilation failed, the interpreter draws instead: %s", msg);
		set_err(c, full, cudaSuccess);
		return;
	}
	if (want_v && k.vertex) { P.vs_kind = SWVS_JIT; P.jit_vertex = k.vertex; }
	if (want_r && k.raster) { P.fs_kind = SWFS_JIT; P.jit_raster = k.raster; }
}

static int launch_draw(swgldev_ctx* c, DrawParams& P)
{
	P.cap_bands = (uint32_t)(c->cap_bands > 0xffffffffull ? 0xffffffffull : c->cap_bands);
	P.bin_cap = c->bin_cap;
	P.bands = c->bands; P.pairs = c->pairs;
	/* overflow pool of deep tiles: warp rasteriser only (the CTA cross-check kernels have no gather step) */
	P.max_hot = SWGL_MAX_HOT_TILES;
	P.ov_pool = c->ov_pool; P.ov_cap = (P.th_shift < SWGL_TILE_SHIFT && c->opt_overflow_pool) ? (uint32_t)c->cap_ov : 0u;
	P.hot_store = c->hot_store; P.hot_cap = (uint32_t)c->cap_hot;
	c->last_draw = P; c->last_draw_valid = 1;

	const bool timing = c->opt_stage_timing != 0;
#define STAGE(i) do { if (timing) cudaEventRecord(c->stage_ev[i], c->stream); } while (0)
	const uint32_t vb = (P.n_shade + 255u) / 256u;
	STAGE(0);
	if (launch_vertex(c, P, vb ? vb : 1)) return -1;
	STAGE(1);
	{
		/* programmatic dependent launch: the set-up CTAs become resident while the vertex kernel's last
		 * wave runs and wait (cudaGridDependencySynchronize) only before they read its output */
		cudaLaunchConfig_t cfg;
		memset(&cfg, 0, sizeof(cfg));
		uint32_t groups = (P.ntri + 127u) / 128u;
		const bool pipelined = c->opt_setup_pipelined == 1;
		if (pipelined && groups > 148u * SETUP_CTAS_PER_SM) groups = 148u * SETUP_CTAS_PER_SM;     /* one resident wave, the threads stride over the stream */
		cfg.gridDim = dim3(groups ? groups : 1u); cfg.blockDim = dim3(128); cfg.stream = c->stream;
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[0].val.programmaticStreamSerializationAllowed = timing ? 0 : 1;
		cfg.attrs = at; cfg.numAttrs = 1;
		if (P.setup_big)
		{
			/* a draw of big triangles: a warp per triangle (k_setup_big) */
			cfg.gridDim = dim3((P.ntri + 7u) / 8u); cfg.blockDim = dim3(256);
			CK(cudaLaunchKernelEx(&cfg, k_setup_big, P));
		}
		else if (pipelined) CK(cudaLaunchKernelEx(&cfg, k_setup_bin<true>, P));
		else CK(cudaLaunchKernelEx(&cfg, k_setup_bin<false>, P));
	}
	if (!P.inline_tall) k_bin_tall<<<148 * SWGL_BIN_TALL_CTAS_PER_SM, 256, 0, c->stream>>>(P);
	STAGE(2);
	/* overflow flags are final once set-up is done: snapshot them on the side stream so the next
	 * draw can be queued while this one is still rasterising */
	CK(cudaEventRecord(c->setup_event, c->stream));
	CK(cudaStreamWaitEvent(c->side, c->setup_event, 0));
	CK(cudaMemcpyAsync(c->h_ctr, c->ctr, 16, cudaMemcpyDeviceToHost, c->side));
	CK(cudaEventRecord(c->ctr_event, c->side));
	c->ctr_pending = 1;
	if (guard_color_write(c)) return -1;
	if (P.fs_kind == SWFS_JIT)
	{
		/* the program's own raster kernel (warp rasteriser only: apply_jit() checks the path) */
		void* args[] = { (void*)&P };
		c->last_raster_path = 3;
		CK(cudaLaunchKernel((const void*)P.jit_raster, dim3((P.tiles_x + WT_WARPS - 1) / WT_WARPS, P.owned_tile_rows ? P.owned_tile_rows : 1), dim3(WT_WARPS * 32), args, 0, c->stream));
	}
	else if (P.fs_kind == SWFS_VARYING) launch_raster<SWFS_VARYING>(c, P);
	else if (P.fs_kind == SWFS_TEXTURE) launch_raster<SWFS_TEXTURE>(c, P);
	else launch_raster<SWFS_GENERIC>(c, P);
	c->last_vs_kind = P.vs_kind; c->last_fs_kind = P.fs_kind;
	STAGE(3);
#undef STAGE
	if (c->opt_mip_lod) { k_last_level<<<1, 1, 0, c->stream>>>(P); c->n_launches++; }     /* what a later GL_POINTS draw samples with */
	c->n_launches += P.inline_tall ? 3 : 4;
	CK(cudaGetLastError());
	if (timing)
	{
		CK(cudaEventSynchronize(c->stage_ev[3]));
		for (int i = 0; i < 3; i++)
		{
			float ms = 0.0f;
			if (cudaEventElapsedTime(&ms, c->stage_ev[i], c->stage_ev[i + 1]) == cudaSuccess) c->stage_us[i] += 1000.0 * ms;
		}
		c->stage_draws++;
	}
	return 0;
}

/* Everything of DrawParams that does not depend on the primitive type.  Returns 1 when the draw
 * has to be skipped (error string set), 0 otherwise. */
static int fill_common_params(swgldev_ctx* c, const swgldev_draw* d, DrawParams& P)
{
	if (d->varying_floats > SWGL_MAX_VARYING_FLOATS || d->vs_words > SWGL_MAX_VAR_WORDS || d->fs_words > SWGL_MAX_VAR_WORDS)
	{
		set_err(c, "draw skipped: shader interface too large", cudaSuccess);
		return 1;
	}
	memset(&P, 0, sizeof(P));
	P.color = c->color; P.depth = c->depth; P.peer_color = c->peer_color; P.W = c->W; P.H = c->H;
	P.vx = d->vx; P.vy = d->vy; P.vw = d->vw; P.vh = d->vh;
	P.hw = (float)(d->vw / 2u); P.hh = (float)(d->vh / 2u);
	P.fvx = (float)d->vx; P.fvy = (float)d->vy;
	P.xlimit = (float)(uint32_t)((uint32_t)d->vx + d->vw);
	P.ylimit = (float)(uint32_t)((uint32_t)d->vy + d->vh);
	P.ytop = (int32_t)((int64_t)d->vh - 1 + 2 * (int64_t)d->vy - c->fold_shift);
	/* A viewport that ends below row 0: the reference's row limit VY + VH wraps around in unsigned arithmetic
	 * (swgl.c:3344, 3356), so its triangles are walked up to their last row whatever the viewport height, and every
	 * one of those rows is stored on row Height-1.  The virtual target of draw_folded() only has the viewport's own
	 * rows: the tile kernels get the limit that was meant, k_fold_row the wrapped one (draw_folded puts it back). */
	if (c->in_fold && (int64_t)d->vy + (int64_t)d->vh < 0) P.ylimit = (float)((int64_t)d->vy + (int64_t)d->vh);
	P.rank = c->rank; P.n_ranks = c->n_ranks; P.band_rows = c->band_rows ? c->band_rows : 1;
	P.vbo = (const uint8_t*)(uintptr_t)d->vbo; P.vbo_bytes = d->vbo_bytes;
	P.ibo = (const uint32_t*)(uintptr_t)d->ibo; P.ibo_count = d->ibo_bytes / 4u;
	P.first = d->first; P.count = d->count;
	P.nvf = d->varying_floats;
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
		P.tex[u].mips = (const uint32_t*)(uintptr_t)d->tex[u].mips; P.tex[u].n_mips = d->tex[u].mips ? d->tex[u].n_mips : 0;
		if (c->opt_mip_lod && P.tex[u].n_mips > 0) P.mip_lod = 1u;
	}
	/* the per-triangle LOD lives in the generic fragment path: a texture-shaped shader takes it too */
	if (P.mip_lod && P.fs_kind == SWFS_TEXTURE) P.fs_kind = SWFS_GENERIC;
	/* the built-in fragment shapes fetch a vec4 varying with 128-bit loads (swgl_host.c lays the records out for
	 * that); a caller of this layer that packs them differently gets the IR path */
	if (P.fs_kind == SWFS_VARYING && ((P.nvf | P.fs_slot) & 3u)) P.fs_kind = SWFS_GENERIC;
	if (d->vs_image) memcpy(P.vs_image, d->vs_image, 4u * d->vs_words);
	if (d->fs_image) memcpy(P.fs_image, d->fs_image, 4u * d->fs_words);
	P.count_fragments = (uint32_t)c->opt_count_fragments;
	P.diag = (uint32_t)c->opt_diag;
	if (P.vs_kind == SWVS_GENERIC)
	{
		P.vs_ops = upload_code(c, d->vs_code_id, d->vs_code); P.vs_nops = d->vs_code->n_ops;
		if (!P.vs_ops) { set_err(c, "vertex shader upload failed", cudaGetLastError()); return 1; }
	}
	if (P.fs_kind == SWFS_GENERIC)
	{
		P.fs_ops = upload_code(c, d->fs_code_id, d->fs_code); P.fs_nops = d->fs_code->n_ops;
		if (!P.fs_ops) { set_err(c, "fragment shader upload failed", cudaGetLastError()); return 1; }
	}
	P.tile_count = c->tile_count; P.ctr = c->ctr; P.lut255 = c->lut255; P.last_level = c->d_last_level;
	return 0;
}

static int draw_folded(swgldev_ctx* c, const swgldev_draw* d);
static int group_draw_folded(swgldev_ctx* c, const swgldev_draw* d);

int swgldev_draw_triangles(swgldev_ctx* c, const swgldev_draw* d)
{
	if (IS_GROUP(c))
	{
		/* a viewport that leaves the framebuffer rows: the fold needs the whole frame on one device */
		if ((d->vy < 0 || (uint64_t)d->vy + d->vh > c->H) && d->vw <= 0x7fffffffu && d->vh <= 0x7fffffffu && c->tiles_x <= 2047u
		    && (d->count + 2u) / 3u != 0 && d->vh != 0)
		{
			const int rc = group_draw_folded(c, d);
			if (rc != 1) return rc;
		}
		return group_each(c, [&](int i, swgldev_ctx* m)
		{
			swgldev_draw di = *d;
			di.vbo = member_ptr(c, i, d->vbo); di.ibo = member_ptr(c, i, d->ibo);
			for (int u = 0; u < SWGL_MAX_TEX_UNITS; u++) { di.tex[u].data = member_ptr(c, i, d->tex[u].data); di.tex[u].mips = member_ptr(c, i, d->tex[u].mips); }
			return swgldev_draw_triangles(m, &di);
		});
	}
	cudaSetDevice(c->device);
	if (settle_last_draw(c)) return -1;
	c->last_draw_valid = 0;

	/* The tile mapping needs storage row = VH-1+2*VY-y to be a bijection on the viewport rows,
	 * i.e. the viewport lies inside the framebuffer vertically (otherwise the reference clamps
	 * several raster rows onto row Height-1, swgl.c:3386). */
	const bool outside = d->vy < 0 || (uint64_t)d->vy + d->vh > c->H;
	if (outside && !c->in_fold && d->vw <= 0x7fffffffu && d->vh <= 0x7fffffffu && c->tiles_x <= 2047u)
	{
		const int rc = draw_folded(c, d);
		if (rc != 1) return rc;           /* 1: not possible here, refused below */
	}
	if ((outside && !c->in_fold) || d->vw > 0x7fffffffu || d->vh > 0x7fffffffu
	    || c->tiles_x > 2047u || (c->H + (1u << WT_H_SHIFT) - 1) / (1u << WT_H_SHIFT) > 1023u)
	{
		set_err(c, "draw skipped: viewport must lie inside the framebuffer rows (0 <= y, y+height <= Height)", cudaSuccess);
		c->draws_refused++;
		return flush_clear(c);
	}
	const uint32_t ntri = (d->count + 2u) / 3u;
	if (ntri == 0 || d->vh == 0) return flush_clear(c);
	if (ntri > (1u << 30))
	{   /* list entries are (primitive id << 1) | flag with primitive id = 2t + k */
		set_err(c, "draw skipped: more than 2^30 triangles in one draw", cudaSuccess);
		return flush_clear(c);
	}

	DrawParams P;
	if (fill_common_params(c, d, P)) return flush_clear(c);
	P.ntri = ntri;
	const int rpath = raster_path_for(c, ntri);
	P.th_shift = th_shift_of(c, rpath);
	/* a fragment shader outside the built-in shapes has its kernel compiled for 8-row tiles only */
	if (rpath == 3 && P.fs_kind == SWFS_GENERIC && c->opt_jit && !c->jit_failed) P.th_shift = WT_H_SHIFT;
	c->cur_th_shift = P.th_shift;
	P.tiles_x = c->tiles_x; P.tiles_y = (c->H + (1u << P.th_shift) - 1u) >> P.th_shift;
	P.owned_tile_rows = P.tiles_y;
	if (P.n_ranks > 1 && rpath == 3)
	{
		const uint32_t per = P.band_rows << (5u - P.th_shift), cycle = per * P.n_ranks;
		/* whole ownership cycles, plus what the last partial cycle leaves to this rank */
		const uint32_t full = P.tiles_y / cycle, rest = P.tiles_y % cycle;
		const uint32_t lo = P.rank * per;
		P.owned_tile_rows = full * per + (rest > lo ? (rest - lo < per ? rest - lo : per) : 0u);
	}
	apply_jit(c, d, P, rpath == 3 && P.th_shift == WT_H_SHIFT);
	P.lean_prims = (c->opt_lean_prims && rpath == 3) ? 1u : 0u;
	P.inline_tall = (rpath == 3 && small_triangle_draw(c, ntri)) ? 1u : 0u;
	/* draws of big triangles: a warp per triangle sets them up and inserts them (k_setup_big); the serial walk that is
	 * left for near-clipped ones inserts inline, so there is nothing for k_bin_tall */
	P.setup_big = (!P.inline_tall && c->opt_setup_big) ? 1u : 0u;
	if (P.setup_big) P.inline_tall = 1u;
	P.n_shade = d->ibo ? d->n_vertices : 3u * ntri;
	P.clip_vid_base = P.n_shade;
	/* the raster kernel addresses the varying of a built-in shape with 32-bit float offsets */
	if (((uint64_t)P.n_shade + 2ull * ntri + 1ull) * (uint64_t)(P.nvf ? P.nvf : 1u) + 16ull >= (1ull << 32))
	{
		set_err(c, "draw skipped: varying records of more than 2^32 floats", cudaSuccess);
		return flush_clear(c);
	}

	/* scratch */
	const size_t n_prims = 2ull * ntri;
	if (grow(c, &c->clip, &c->cap_clip, (size_t)P.n_shade)) return -1;
	if (grow(c, &c->clip_xy, &c->cap_clip_xy, (size_t)P.n_shade)) return -1;
	if (grow(c, &c->vary, &c->cap_vary, ((size_t)P.n_shade + n_prims) * (P.nvf ? P.nvf : 1) + 4)) return -1;
	if (grow(c, &c->prims, &c->cap_prims, n_prims)) return -1;
	if (grow(c, &c->bands, &c->cap_bands, (size_t)1 << 16)) return -1;
	{
		/* per-tile lists: K entries each, K grows (never shrinks) when a draw overflows it */
		const size_t ntiles = (size_t)c->tiles_x * ((c->H + (1u << c->cur_th_shift) - 1) >> c->cur_th_shift);   /* tiles of the current draw */
		if (c->bin_cap == 0) c->bin_cap = 256;
		if (grow(c, &c->pairs, &c->cap_pairs, ntiles * c->bin_cap)) return -1;
		if (ensure_overflow_pool(c, c->cap_ov ? c->cap_ov : ((size_t)1 << 18))) return -1;
	}
	P.clip = c->clip; P.clip_xy = c->clip_xy; P.vary = c->vary; P.prims = c->prims; P.prims2 = c->prims + ntri;

	/* fused clear: the raster kernel starts the covered pixels from the clear value */
	P.clear = c->pending_clear;
	c->pending_clear.flags = 0;

	/* write-through host mirror (single GPU, no peer target): see swgldev_ctx */
	c->draws_since_map++;
	/* a fused clear of the whole framebuffer makes every tile dirty: the mirror is complete afterwards */
	const bool full_clear = (P.clear.flags & 1u) && P.clear.x0 <= 0 && P.clear.y0 <= 0
	                        && P.clear.x1 >= (int32_t)c->W && P.clear.y1 >= (int32_t)c->H;
	if (c->shared_mirror)
	{
		/* shared frame mirror: this rank's bands of it stay in sync as long as every colour write goes through */
		if (!c->color_exposed && (c->mirror_synced || full_clear)) { c->mirror_synced = 1; P.peer_color = c->shared_mirror_dev; c->wt_draws++; }
		else { c->mirror_synced = 0; P.peer_color = nullptr; }
	}
	else if (c->opt_host_mirror && !c->in_fold && !P.peer_color && c->n_ranks == 1 && !c->color_exposed && (c->mirror_synced || full_clear)
	    && (c->opt_host_mirror == 2 || (c->wt_predict && c->draws_since_map <= 2)))
	{
		c->mirror_synced = 1;
		P.peer_color = c->h_color;
		c->wt_draws++;
	}
	else
		c->mirror_synced = 0;

	c->n_draws++;
	c->stats.draws = c->n_draws;
	c->stats.triangles_in = ntri;
	if (launch_draw(c, P)) return -1;
	return stamp_draw(c, d);
}


/* A draw whose viewport leaves the framebuffer rows (see k_fold_row).  Returns 0 / -1 when the draw has been
 * handled, 1 when this configuration cannot do it (sort-first ranks, assembled targets, a virtual framebuffer
 * beyond the tile-row limit): the caller then refuses the draw as before. */
static int draw_folded(swgldev_ctx* c, const swgldev_draw* d)
{
	if (c->n_ranks > 1 || c->peer_color || c->shared_mirror || c->n_group > 1 || c->mirror_borrowed) return 1;
	if (raster_path_for(c, 0) != 3) return 1;
	const int64_t s_lo = d->vy < 0 ? (int64_t)d->vy : 0;
	const int64_t s_hi = (int64_t)d->vy + (int64_t)d->vh > (int64_t)c->H ? (int64_t)d->vy + (int64_t)d->vh : (int64_t)c->H;
	const int64_t Hv = s_hi - s_lo;
	if (Hv <= 0 || (Hv + 7) / 8 > 1023 || d->vh == 0 || c->H < 1 || c->W == 0) return 1;
	if (flush_clear(c)) return -1;
	if (guard_color_write(c)) return -1;
	const size_t npx = (size_t)c->W * (size_t)Hv;
	const size_t ntiles = (size_t)c->tiles_x * (size_t)((Hv + (1 << WT_H_SHIFT_MIN) - 1) >> WT_H_SHIFT_MIN) + 1;
	uint32_t* vcol = nullptr; float* vdep = nullptr; uint32_t* vcount = nullptr;
	CK(cudaMallocAsync((void**)&vcol, npx * 4, c->stream));
	CK(cudaMallocAsync((void**)&vdep, npx * 4, c->stream));
	CK(cudaMallocAsync((void**)&vcount, ntiles * 4, c->stream));
	CK(cudaMemsetAsync(vcol, 0, npx * 4, c->stream));
	/* the rows of the virtual target that stand for the fold hold NaN depth: no fragment passes the test there (cur == 0
	 * and cur >= z are both false), so the tile kernels only count -- and shade -- what lands on real rows; every
	 * fragment is still TESTED there exactly once, which is the reference's count */
	CK(cudaMemsetAsync(vdep, 0xff, npx * 4, c->stream));
	CK(cudaMemsetAsync(vcount, 0, ntiles * 4, c->stream));
	/* real rows 0 .. H-2 keep their place in the virtual framebuffer; row H-1 and the rows outside are the fold */
	const size_t keep = (size_t)(c->H - 1u) * c->W * 4;
	const size_t off = (size_t)(-s_lo) * c->W;
	if (keep)
	{
		CK(cudaMemcpyAsync(vcol + off, c->color, keep, cudaMemcpyDeviceToDevice, c->stream));
		CK(cudaMemcpyAsync(vdep + off, c->depth, keep, cudaMemcpyDeviceToDevice, c->stream));
	}
	uint32_t* const rcol = c->color; float* const rdep = c->depth; uint32_t* const rcount = c->tile_count;
	const uint32_t rH = c->H, rty = c->tiles_y;
	c->color = vcol; c->depth = vdep; c->tile_count = vcount; c->H = (uint32_t)Hv; c->tiles_y = (uint32_t)((Hv + SWGL_TILE - 1) / SWGL_TILE);
	c->fold_shift = s_lo; c->in_fold = 1;
	int rc = swgldev_draw_triangles(c, d);
	if (!rc) rc = settle_last_draw(c);        /* a scratch overflow is resolved against the virtual target */
	DrawParams P = c->last_draw;
	const bool drew = c->last_draw_valid != 0;
	c->color = rcol; c->depth = rdep; c->tile_count = rcount; c->H = rH; c->tiles_y = rty;
	c->fold_shift = 0; c->in_fold = 0; c->last_draw_valid = 0; c->mirror_synced = 0;
	if (!rc && keep)
	{
		CK(cudaMemcpyAsync(c->color, vcol + off, keep, cudaMemcpyDeviceToDevice, c->stream));
		CK(cudaMemcpyAsync(c->depth, vdep + off, keep, cudaMemcpyDeviceToDevice, c->stream));
	}
	if (!rc && drew)
	{
		/* row H-1: every aliased raster row of every primitive, in submission order */
		P.color = c->color; P.depth = c->depth; P.peer_color = nullptr; P.H = c->H;
		P.ytop = (int32_t)((int64_t)d->vh - 1 + 2 * (int64_t)d->vy);
		P.ylimit = (float)(uint32_t)((uint32_t)d->vy + d->vh);
		P.clear.flags = 0;
		if (P.fs_kind == SWFS_JIT) P.fs_kind = SWFS_GENERIC;      /* the interpreter: its op list is uploaded with every generic draw */
		int x_lo = d->vx > 0 ? d->vx : 0;
		int64_t xh = (int64_t)d->vx + (int64_t)d->vw;
		int x_hi = xh < (int64_t)c->W ? (int)xh : (int)c->W;
		if (x_hi > x_lo)
		{
			const dim3 grid(((uint32_t)(x_hi - x_lo) + 127u) / 128u);
			/* fragments are counted as tested by whoever walks their row: the tile kernels up to the row limit they
			 * were given, this kernel beyond it (only a viewport that ends below row 0 has rows beyond) */
			const int64_t top = (int64_t)d->vy + (int64_t)d->vh;
			const int count_from = top < 0 ? (int)top : 0x7fffffff;
			if (P.fs_kind == SWFS_VARYING) k_fold_row<SWFS_VARYING><<<grid, 128, 0, c->stream>>>(P, x_lo, x_hi, count_from);
			else if (P.fs_kind == SWFS_TEXTURE) k_fold_row<SWFS_TEXTURE><<<grid, 128, 0, c->stream>>>(P, x_lo, x_hi, count_from);
			else k_fold_row<SWFS_GENERIC><<<grid, 128, 0, c->stream>>>(P, x_lo, x_hi, count_from);
			c->n_launches++;
			/* k_fold_row reads the element buffer, the textures and the draw's scratch: the draw's event has to lie
			 * behind it, or a re-specification of those buffers that follows the draw would only wait for the tile
			 * kernels (found by the API-sequence fuzz: an index buffer re-specified right after a folded draw) */
			if (stamp_draw(c, d)) rc = -1;
		}
		c->draws_folded++;
	}
	cudaFreeAsync(vcol, c->stream); cudaFreeAsync(vdep, c->stream); cudaFreeAsync(vcount, c->stream);
	CK(cudaGetLastError());
	return rc ? -1 : 0;
}

/* The same for a device group: the frame is spread over the members in bands, and the fold is one row that collects
 * fragments of every primitive.  The members' bands are brought together in the leader's attachments (peer copies),
 * the leader renders the draw alone like a single device does (draw_folded), and every member gets its bands back.
 * Correct, not fast -- like the single-device fold.  Returns 1 when it cannot be done (the draw is then refused). */
static int group_draw_folded(swgldev_ctx* c, const swgldev_draw* d)
{
	const int n = c->n_group;
	if (raster_path_for(c, 0) != 3 || c->peer_color) return 1;
	/* every member: earlier draws finished, its pending clear applied to its own rows */
	if (group_each(c, [&](int, swgldev_ctx* m)
	    {
		    if (settle_last_draw(m) || flush_clear(m)) return -1;
		    return cudaStreamSynchronize(m->stream) == cudaSuccess ? 0 : -1;
	    })) { set_err(c, "folded draw on a device group: a member could not be brought to rest", cudaSuccess); return -1; }
	cudaSetDevice(c->device);
	const size_t band_px = (size_t)32u * (c->band_rows ? c->band_rows : 1u) * c->W, total = (size_t)c->W * c->H;
	auto move_bands = [&](bool to_leader) -> int
	{
		for (int i = 1; i < n; i++)
		{
			swgldev_ctx* m = c->group[i];
			for (size_t b = (size_t)i; b * band_px < total; b += (size_t)n)
			{
				const size_t off = b * band_px, cnt = (off + band_px < total ? band_px : total - off) * 4;
				if (to_leader)
				{
					CK(cudaMemcpyPeerAsync(c->color + off, c->device, m->color + off, m->device, cnt, c->stream));
					CK(cudaMemcpyPeerAsync(c->depth + off, c->device, m->depth + off, m->device, cnt, c->stream));
				}
				else
				{
					CK(cudaMemcpyPeerAsync(m->color + off, m->device, c->color + off, c->device, cnt, c->stream));
					CK(cudaMemcpyPeerAsync(m->depth + off, m->device, c->depth + off, c->device, cnt, c->stream));
				}
			}
		}
		return 0;
	};
	if (move_bands(true)) return -1;
	/* the leader alone, as a single device that owns every row and keeps no assembled mirror */
	const uint32_t rank = c->rank, n_ranks = c->n_ranks;
	uint32_t* const sm = c->shared_mirror; uint32_t* const smd = c->shared_mirror_dev;
	const int borrowed = c->mirror_borrowed;
	c->rank = 0; c->n_ranks = 1; c->shared_mirror = nullptr; c->shared_mirror_dev = nullptr; c->mirror_borrowed = 0; c->n_group = 1;
	int rc = draw_folded(c, d);
	if (!rc) rc = settle_last_draw(c);
	c->rank = rank; c->n_ranks = n_ranks; c->shared_mirror = sm; c->shared_mirror_dev = smd; c->mirror_borrowed = borrowed; c->n_group = n;
	if (rc == 1) return 1;
	if (!rc) rc = move_bands(false);
	if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = -1;
	/* the members' bands of the assembled mirror are out of date: their next read-back copies them */
	for (int i = 0; i < n; i++) c->group[i]->mirror_synced = 0;
	/* the draw's counters are the leader's: the other members still hold those of their last own draw */
	group_each(c, [&](int i, swgldev_ctx* m)
	{
		if (!i) return 0;
		cudaMemsetAsync(m->ctr, 0, sizeof(Counters), m->stream);
		/* ... and the level of detail the draw's last triangle left (k_last_level ran on the leader only) */
		if (c->opt_mip_lod) cudaMemcpyPeerAsync(m->d_last_level, m->device, c->d_last_level, c->device, sizeof(float), m->stream);
		return 0;
	});
	cudaSetDevice(c->device);
	if (rc) set_err(c, "folded draw on a device group failed", cudaGetLastError());
	return rc ? -1 : 0;
}

int swgldev_precompile(swgldev_ctx* c, const swgldev_draw* d, char* msg, size_t msg_len)
{
	if (msg_len) msg[0] = 0;
	const int want_v = d->vs_kind == SWVS_GENERIC, want_r = d->fs_kind == SWFS_GENERIC;
	if (!want_v && !want_r) return 0;
	if (c) cudaSetDevice(c->device);
	swgljit_kernels k;
	return swgljit_get(c ? c->device : -1, d, want_v, want_r, &k, msg, msg_len);
}

int swgldev_draw_points(swgldev_ctx* c, const swgldev_draw* d)
{
	if (IS_GROUP(c))
	{
		return group_each(c, [&](int i, swgldev_ctx* m)
		{
			swgldev_draw di = *d;
			di.vbo = member_ptr(c, i, d->vbo); di.ibo = member_ptr(c, i, d->ibo);
			for (int u = 0; u < SWGL_MAX_TEX_UNITS; u++) { di.tex[u].data = member_ptr(c, i, d->tex[u].data); di.tex[u].mips = member_ptr(c, i, d->tex[u].mips); }
			return swgldev_draw_points(m, &di);
		});
	}
	cudaSetDevice(c->device);
	if (settle_last_draw(c)) return -1;
	c->last_draw_valid = 0;
	if (flush_clear(c)) return -1;           /* points overwrite pixels: the clear must land first */
	if (d->count == 0) return 0;

	DrawParams P;
	if (fill_common_params(c, d, P)) return 0;
	P.n_shade = d->ibo ? d->n_vertices : d->count;
	P.clip_vid_base = P.n_shade;
	P.th_shift = SWGL_TILE_SHIFT; P.tiles_x = c->tiles_x; P.tiles_y = c->tiles_y;
	if (grow(c, &c->clip, &c->cap_clip, (size_t)P.n_shade)) return -1;
	if (grow(c, &c->clip_xy, &c->cap_clip_xy, (size_t)P.n_shade)) return -1;
	if (grow(c, &c->vary, &c->cap_vary, (size_t)P.n_shade * (P.nvf ? P.nvf : 1) + 4)) return -1;
	if (!c->winner)
	{
		const size_t bytes = (size_t)c->W * c->H * 4;
		CK(cudaMalloc((void**)&c->winner, bytes ? bytes : 4));
		CK(cudaMemsetAsync(c->winner, 0, bytes, c->stream));
	}
	P.clip = c->clip; P.clip_xy = c->clip_xy; P.vary = c->vary; P.winner = c->winner;

	const uint32_t vb = (P.n_shade + 255u) / 256u, pb = (P.count + 255u) / 256u;
	apply_jit(c, d, P, false);                /* the vertex stage; k_points_write keeps the interpreter for generic shaders */
	if (launch_vertex(c, P, vb ? vb : 1)) return -1;
	if (guard_color_write(c)) return -1;
	k_points_claim<<<pb, 256, 0, c->stream>>>(P);
	if (P.fs_kind == SWFS_VARYING) k_points_write<SWFS_VARYING><<<pb, 256, 0, c->stream>>>(P);
	else if (P.fs_kind == SWFS_TEXTURE) k_points_write<SWFS_TEXTURE><<<pb, 256, 0, c->stream>>>(P);
	else k_points_write<SWFS_GENERIC><<<pb, 256, 0, c->stream>>>(P);
	c->n_launches += 3;
	c->n_draws++;
	c->draws_since_map++;
	c->mirror_synced = 0;
	c->stats.draws = c->n_draws;
	CK(cudaGetLastError());
	return stamp_draw(c, d);
}

/* rows [r0, r1) of the colour attachment this rank owns (storage rows), band by band */
static int flush_owned_bands(swgldev_ctx* c)
{
	const uint32_t band_px = 32u * (c->band_rows ? c->band_rows : 1u);
	for (uint32_t b = 0; (size_t)b * band_px < c->H; b++)
	{
		if (c->n_ranks > 1 && b % c->n_ranks != c->rank) continue;
		const size_t r0 = (size_t)b * band_px, r1 = (r0 + band_px < c->H) ? r0 + band_px : c->H;
		CK(cudaMemcpyAsync(c->shared_mirror + r0 * c->W, c->color + r0 * c->W, (r1 - r0) * c->W * 4, cudaMemcpyDeviceToHost, c->stream));
	}
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

uint32_t* swgldev_map_color(swgldev_ctx* c)
{
	if (IS_GROUP(c))
	{
		/* every member brings its bands of the leader's pinned mirror up to date (written through by its raster
		 * kernels, or copied by its swgldev_sync) */
		group_each(c, [&](int, swgldev_ctx* m) { swgldev_map_color(m); return 0; });
		return c->h_mirror[0];
	}
	if (c->shared_mirror)
	{
		/* swgldev_sync has brought this rank's bands up to date; the other ranks' are theirs to finish
		 * (the application's barrier, as with the peer colour target) */
		swgldev_sync(c);
		c->draws_since_map = 0;
		return c->shared_mirror;
	}
	if (swgldev_sync(c)) { c->mirror_synced = 0; return c->h_color; }
	if (!c->mirror_synced)
	{
		cudaMemcpyAsync(c->h_color, c->color, (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->stream);
		cudaStreamSynchronize(c->stream);
	}
	/* an application that reads back after every draw or two gets the write-through mirror next time */
	c->wt_predict = c->draws_since_map <= 2;
	c->draws_since_map = 0;
	c->mirror_synced = (c->n_ranks == 1 && !c->peer_color && !c->color_exposed) ? 1 : 0;
	return c->h_color;
}

/* ---- frame pipelining ---- */
uint64_t swgldev_frame_submit(swgldev_ctx* c)
{
	if (IS_GROUP(c)) { set_err(c, "swglFrameSubmit: not available with several devices (swglSetDeviceCount)", cudaSuccess); return 0; }
	cudaSetDevice(c->device);
	if (c->shared_mirror) { set_err(c, "swglFrameSubmit: not available with a shared frame mirror", cudaSuccess); return 0; }
	if (settle_last_draw(c) || flush_clear(c)) return 0;
	const size_t bytes = (size_t)c->W * c->H * 4;
	if (!c->h_mirror[1])
	{
		if (cudaMallocHost((void**)&c->h_mirror[1], bytes ? bytes : 4) != cudaSuccess) { set_err(c, "cudaMallocHost (second frame mirror)", cudaGetLastError()); return 0; }
	}
	const uint32_t slot = (uint32_t)(c->frame_serial & 1u);
	cudaError_t e = cudaSuccess;
	if (c->mirror_synced)
		e = cudaEventRecord(c->frame_ev[slot], c->stream);          /* already written through */
	else
	{
		/* copy engine on its own stream: the next frame's uploads, vertex stage and set-up overlap it */
		if (c->copy_inflight) e = cudaStreamWaitEvent(c->stream, c->frame_ev[c->copy_inflight - 1], 0);
		if (e == cudaSuccess) e = cudaEventRecord(c->frame_done, c->stream);
		if (e == cudaSuccess) e = cudaStreamWaitEvent(c->copy, c->frame_done, 0);
		if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_mirror[slot], c->color, bytes, cudaMemcpyDeviceToHost, c->copy);
		if (e == cudaSuccess) e = cudaEventRecord(c->frame_ev[slot], c->copy);
		c->copy_inflight = (int)slot + 1;
	}
	if (e != cudaSuccess) { set_err(c, "swglFrameSubmit", e); return 0; }
	c->frame_serial++;
	/* the next frame goes to the other mirror.  No write-through while frames are pipelined: SM
	 * stores to host memory starve the next frame's upload of PCIe read requests (measured). */
	c->h_color = c->h_mirror[c->frame_serial & 1u];
	c->mirror_synced = 0;
	c->wt_predict = 0;
	c->draws_since_map = 0;
	return c->frame_serial;     /* ticket */
}

const uint32_t* swgldev_frame_wait(swgldev_ctx* c, uint64_t ticket)
{
	cudaSetDevice(c->device);
	if (ticket == 0 || ticket > c->frame_serial || c->frame_serial - ticket >= 2u)
	{
		set_err(c, "swglFrameWait: ticket is not one of the last two submitted frames", cudaSuccess);
		return nullptr;
	}
	const uint32_t slot = (uint32_t)((ticket - 1u) & 1u);
	if (cudaEventSynchronize(c->frame_ev[slot]) != cudaSuccess) { set_err(c, "cudaEventSynchronize (frame)", cudaGetLastError()); return nullptr; }
	return c->h_mirror[slot];
}

/* glGetFramePtr consumers that want bytes in R, G, B, A order: swizzled on the device, then copied */
int swgldev_read_rgba8(swgldev_ctx* c, void* dst)
{
	if (IS_GROUP(c))
	{   /* the assembled frame only exists in host memory: swizzle it there */
		const uint32_t* src = swgldev_map_color(c);
		uint8_t* o = (uint8_t*)dst;
		for (size_t i = 0, n = (size_t)c->W * c->H; i < n; i++) { const uint32_t w = src[i]; o[4 * i] = (uint8_t)(w >> 24); o[4 * i + 1] = (uint8_t)(w >> 16); o[4 * i + 2] = (uint8_t)(w >> 8); o[4 * i + 3] = (uint8_t)w; }
		return 0;
	}
	if (swgldev_sync(c)) return -1;
	const size_t n = (size_t)c->W * c->H;
	if (!n) return 0;
	if (!c->rgba_staging) CK(cudaMalloc((void**)&c->rgba_staging, n * 4));
	k_pack_rgba8<<<1184, 256, 0, c->stream>>>(c->color, c->rgba_staging, n);
	c->n_launches++;
	CK(cudaMemcpyAsync(dst, c->rgba_staging, n * 4, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

/* rows of the depth attachment this member owns, into `dst` (the leader's pinned depth mirror) */
static int copy_owned_depth_bands(swgldev_ctx* m, float* dst)
{
	const uint32_t band_px = 32u * (m->band_rows ? m->band_rows : 1u);
	cudaSetDevice(m->device);
	for (uint32_t b = 0; (size_t)b * band_px < m->H; b++)
	{
		if (m->n_ranks > 1 && b % m->n_ranks != m->rank) continue;
		const size_t r0 = (size_t)b * band_px, r1 = (r0 + band_px < m->H) ? r0 + band_px : m->H;
		if (cudaMemcpyAsync(dst + r0 * m->W, m->depth + r0 * m->W, (r1 - r0) * m->W * 4, cudaMemcpyDeviceToHost, m->stream) != cudaSuccess) return -1;
	}
	return cudaStreamSynchronize(m->stream) == cudaSuccess ? 0 : -1;
}

float* swgldev_map_depth(swgldev_ctx* c)
{
	if (IS_GROUP(c))
	{
		swgldev_sync(c);
		if (group_each(c, [&](int, swgldev_ctx* m) { return copy_owned_depth_bands(m, c->h_depth); }))
			set_err(c, "depth read-back of a device group member failed", cudaGetLastError());
		return c->h_depth;
	}
	if (swgldev_sync(c)) return c->h_depth;
	cudaMemcpyAsync(c->h_depth, c->depth, (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, c->stream);
	cudaStreamSynchronize(c->stream);
	return c->h_depth;
}

swgldev_ptr swgldev_color_devptr(swgldev_ctx* c)
{
	/* the caller may now write the attachment behind the library's back: no write-through mirror */
	c->color_exposed = 1; c->mirror_synced = 0;
	return (swgldev_ptr)(uintptr_t)c->color;
}
swgldev_ptr swgldev_depth_devptr(swgldev_ctx* c) { return (swgldev_ptr)(uintptr_t)c->depth; }

void swgldev_get_stats(swgldev_ctx* c, swgldev_stats* out)
{
	if (IS_GROUP(c))
	{
		/* fragments and list entries add up over the members (a pixel belongs to one of them); a primitive
		 * that spans bands of several members is counted by each */
		swgldev_stats sum, each[SWGL_MAX_GROUP];
		memset(&sum, 0, sizeof(sum));
		group_each(c, [&](int i, swgldev_ctx* m) { swgldev_get_stats(m, &each[i]); return 0; });
		for (int i = 0; i < c->n_group; i++)
		{
			const swgldev_stats& one = each[i];
			sum.draws = one.draws; sum.triangles_in = one.triangles_in;
			sum.prims_out += one.prims_out; sum.tested += one.tested; sum.shaded += one.shaded; sum.tile_pairs += one.tile_pairs; sum.bands += one.bands;
		}
		*out = sum;
		return;
	}
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
	if (IS_GROUP(c)) { set_err(c, "swglSetStripe: the device group shards the frame itself (swglSetDeviceCount)", cudaSuccess); return; }
	swgldev_sync(c);
	c->rank = rank; c->n_ranks = n_ranks ? n_ranks : 1; c->band_rows = band_tile_rows ? band_tile_rows : 1;
	if (c->shared_mirror) c->mirror_synced = 0;     /* other bands are this rank's now */
}

void swgldev_set_peer_color(swgldev_ctx* c, swgldev_ptr peer_color)
{
	if (IS_GROUP(c)) { set_err(c, "swglSetPeerColorTarget: not available with several devices (swglSetDeviceCount)", cudaSuccess); return; }
	swgldev_sync(c);
	c->peer_color = (uint32_t*)(uintptr_t)peer_color;
}

int swgldev_set_shared_mirror(swgldev_ctx* c, void* host_ptr, uint64_t bytes)
{
	if (c->n_group > 1 || c->mirror_borrowed) { set_err(c, "swglSetSharedFrameMirror: the device group assembles the frame in its own mirror (swglSetDeviceCount)", cudaSuccess); return -1; }
	cudaSetDevice(c->device);
	swgldev_sync(c);
	if (c->shared_mirror)
	{
		cudaHostUnregister(c->shared_mirror);
		c->shared_mirror = nullptr; c->shared_mirror_dev = nullptr; c->shared_mirror_bytes = 0;
		c->mirror_synced = 0;
	}
	if (!host_ptr) return 0;
	if (bytes < (uint64_t)c->W * c->H * 4) { set_err(c, "swglSetSharedFrameMirror: the memory is smaller than the colour attachment", cudaSuccess); return -1; }
	cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
	if (e != cudaSuccess) { cudaGetLastError(); set_err(c, "cudaHostRegister (shared frame mirror)", e); return -1; }   /* do not leave the error for the next launch check */
	void* d = nullptr;
	e = cudaHostGetDevicePointer(&d, host_ptr, 0);
	if (e != cudaSuccess) { cudaGetLastError(); cudaHostUnregister(host_ptr); set_err(c, "cudaHostGetDevicePointer (shared frame mirror)", e); return -1; }
	c->shared_mirror = (uint32_t*)host_ptr; c->shared_mirror_dev = (uint32_t*)d; c->shared_mirror_bytes = bytes;
	c->mirror_synced = 0;      /* the next swgldev_sync copies this rank's bands unless a cleared frame has been written through */
	return 0;
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
	if (IS_GROUP(c))
	{
		group_each(c, [&](int, swgldev_ctx* m) { swgldev_set_option(m, name, value); return 0; });
		return;
	}
	swgldev_sync(c);
	if (!strcmp(name, "fuse_clear")) c->opt_fuse_clear = (int)value;
	else if (!strcmp(name, "count_fragments")) c->opt_count_fragments = (int)value;
	else if (!strcmp(name, "raster_path")) c->opt_raster_path = (int)value;
	else if (!strcmp(name, "diag")) c->opt_diag = (int)value;
	else if (!strcmp(name, "lean_prims")) c->opt_lean_prims = (int)value;
	else if (!strcmp(name, "mip_lod")) c->opt_mip_lod = value ? 1 : 0;
	else if (!strcmp(name, "jit")) c->opt_jit = value ? 1 : 0;
	else if (!strcmp(name, "tile_rows")) c->opt_tile_rows = (int)value;
	else if (!strcmp(name, "overflow_pool")) c->opt_overflow_pool = value ? 1 : 0;
	else if (!strcmp(name, "setup_pipelined")) c->opt_setup_pipelined = (int)value;
	else if (!strcmp(name, "setup_big")) c->opt_setup_big = value ? 1 : 0;
	else if (!strcmp(name, "host_mirror")) { c->opt_host_mirror = (int)value; c->mirror_synced = 0; }
	else if (!strcmp(name, "bin_limit_bytes") && value > 0) c->opt_bin_limit = (size_t)value;
	else if (!strcmp(name, "bin_cap") && value > 0)
	{
		/* test hook: force a (small) list capacity for the next draws */
		c->bin_cap = (uint32_t)value;
	}
	else if (!strcmp(name, "selftest_division") && value > 0)
	{
		/* value = number of operand pairs; result in get_option("selftest_division_mismatches") */
		unsigned long long* d = nullptr;
		unsigned long long h = ~0ull;
		if (cudaMalloc((void**)&d, 8) == cudaSuccess)
		{
			cudaMemsetAsync(d, 0, 8, c->stream);
			k_selftest_division<<<148 * 8, 256, 0, c->stream>>>((uint64_t)value, 0x5357474cull + (uint64_t)value, d);
			c->n_launches++;
			if (cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) h = ~0ull;
			cudaFree(d);
		}
		c->selftest_mismatches = (int64_t)h;
	}
	else if (!strcmp(name, "stage_timing"))
	{
		c->opt_stage_timing = (int)value;
		for (int i = 0; i < 8; i++) c->stage_us[i] = 0.0;
		c->stage_draws = 0;
	}
}

int64_t swgldev_get_option(swgldev_ctx* c, const char* name)
{
	if (!strncmp(name, "jit_", 4))
	{
		/* process-wide (the compiled programs outlive a context); `c` may be NULL */
		swgljit_stats js;
		swgljit_get_stats(&js);
		if (!strcmp(name, "jit_compiles")) return (int64_t)js.compiles;
		if (!strcmp(name, "jit_cache_hits")) return (int64_t)js.cache_hits;
		if (!strcmp(name, "jit_compile_us_total")) return (int64_t)(js.compile_ms_total * 1e3);
		if (!strcmp(name, "jit_cubin_bytes")) return (int64_t)js.cubin_bytes;
		return -1;
	}
	if (!c) return -1;
	if (!strcmp(name, "fuse_clear")) return c->opt_fuse_clear;
	if (!strcmp(name, "count_fragments")) return c->opt_count_fragments;
	if (!strcmp(name, "raster_path")) return c->opt_raster_path;
	if (!strcmp(name, "host_mirror")) return c->opt_host_mirror;
	if (!strcmp(name, "mip_lod")) return c->opt_mip_lod;
	if (!strcmp(name, "jit")) return c->opt_jit;
	if (!strcmp(name, "tile_rows")) return c->opt_tile_rows;
	if (!strcmp(name, "last_tile_rows")) return 1 << c->cur_th_shift;
	if (!strcmp(name, "last_vs_kind")) return c->last_vs_kind;
	if (!strcmp(name, "last_fs_kind")) return c->last_fs_kind;
	if (!strcmp(name, "mirror_synced")) return c->mirror_synced;
	if (!strcmp(name, "wt_draws")) return (int64_t)c->wt_draws;
	if (!strcmp(name, "last_raster_path")) return c->last_raster_path;
	if (!strcmp(name, "tile_size")) return SWGL_TILE;
	if (!strcmp(name, "kernel_launches")) return (int64_t)c->n_launches;
	if (!strcmp(name, "stage_draws")) return (int64_t)c->stage_draws;
	if (!strncmp(name, "stage_ns_", 9)) { int i = atoi(name + 9); return (i >= 0 && i < 3) ? (int64_t)(c->stage_us[i] * 1000.0) : -1; }
	if (!strcmp(name, "bin_cap")) return c->bin_cap;
	if (!strcmp(name, "overflow_pool_entries")) return (int64_t)c->cap_ov;
	if (!strcmp(name, "pairs_bytes")) return (int64_t)(c->cap_pairs * 4);
	if (!strcmp(name, "selftest_division_mismatches")) return c->selftest_mismatches;
	if (!strcmp(name, "device")) return c->device;
	if (!strcmp(name, "device_count")) return c->n_group;
	if (!strcmp(name, "draws_refused")) return (int64_t)c->draws_refused;
	if (!strcmp(name, "draws_folded")) return (int64_t)c->draws_folded;
	if (!strcmp(name, "kernel_launches_all_devices"))
	{
		int64_t n = 0;
		for (int i = 0; i < c->n_group; i++) n += (int64_t)(c->n_group > 1 ? c->group[i] : c)->n_launches;
		return n;
	}
	if (!strcmp(name, "sizeof_draw_params")) return (int64_t)sizeof(DrawParams);
	return -1;
}

} /* extern "C" */
