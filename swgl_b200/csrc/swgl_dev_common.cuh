/*
 * swgl_dev_common.cuh -- device code shared by the ahead-of-time build (swgl_dev.cu) and the run-time
 * compiled kernels of arbitrary shaders (swgl_jit.cpp hands this file, the two headers before it and
 * swgl_raster_warp.cuh to NVRTC together with the code generated from a program's IR):
 *
 *   k_vertex<VS>       attribute fetch + vertex shader + varying capture + divide/viewport snap
 *                                                              swgl.c:3618-3666, 3683-3692
 *   prim_ref/load_prim the primitive behind a tile-list entry
 *   run_fragment<FS>   varying interpolation + fragment shader   swgl.c:3394-3408
 *   walk_to_row        span-walk state on entering a tile row     swgl.c:3350-3356, 3466-3471
 *   clamp_color, blend_pack_lut                                   swgl.c:3428-3462
 *
 * SWVS_JIT / SWFS_JIT (only under SWGL_JIT, i.e. inside an NVRTC translation unit) call jit_vertex() /
 * jit_fragment(), the straight-line __device__ functions generated from the shader pair.
 */
#ifndef SWGL_DEV_COMMON_CUH
#define SWGL_DEV_COMMON_CUH

#include "swgl_dev_math.cuh"

struct FragIn;
#ifdef SWGL_JIT
/* generated from the program's IR (swgl_jit.cpp), defined at the end of the translation unit */
__device__ __forceinline__ void jit_vertex(const DrawParams& P, long long vid, float4& pos, float* vout);
__device__ __forceinline__ float4 jit_fragment(const DrawParams& P, const FragIn& f);
#endif

/* programmatic dependent launch (cudaTriggerProgrammaticLaunchCompletion / cudaGridDependencySynchronize):
 * spelled as the PTX they compile to, NVRTC has no device-runtime header */
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float4 to_screen(const float4& p, const DrawParams& P)
{
	/* swgl.c:3685-3691: int <- x / w * (VW/2) + (VW/2) + VX, stored back as float */
	int X = cvt_x86((fdiv(p.x, p.w) * P.hw + P.hw) + P.fvx);
	int Y = cvt_x86((fdiv(p.y, p.w) * P.hh + P.hh) + P.fvy);
	return make_float4((float)X, (float)Y, p.z, p.w);
}

/* ---- vertex stage (swgl.c:3618-3666) ---- */
template <int VS>
__global__ void __launch_bounds__(256) k_vertex(const __grid_constant__ DrawParams P)
{
	uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	pdl_launch_dependents();   /* the set-up kernel may start loading its indices */
	if (v == 0 && blockIdx.x == 0)
	{
		/* per-draw counters: this kernel is the first of the draw */
		P.ctr->band_cursor = 0; P.ctr->max_list = 0; P.ctr->overflow = 0; P.ctr->prims_out = 0; P.ctr->pair_total = 0ull;
		P.ctr->ov_cursor = 0; P.ctr->hot_tiles = 0; P.ctr->hot_cursor = 0;
	}
	if (v < SWGL_CTR_SLOTS && blockIdx.x == 0) { P.ctr->tested[v] = 0ull; P.ctr->shaded[v] = 0ull; }
	if (v >= P.n_shade) return;
	/* glDrawArrays: stream vertex first + v;  glDrawElements: unique vertex v */
	long long vid = P.ibo ? (long long)v : (long long)P.first + (long long)v;
	float4 pos = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	float* vout = P.vary + (size_t)v * P.nvf;

#ifdef SWGL_JIT
	if (VS == SWVS_JIT) jit_vertex(P, vid, pos, vout);
	else
#endif
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
		ir_execute(P.vs_ops, P.vs_nops, V, P, 0.0f);   /* texture() in a vertex shader reads the base level */
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
	/* divide + viewport snap here, once per vertex (swgl.c:3685-3691 does it per triangle corner);
	 * the clip-space x, y are kept for triangles that cross the near plane */
	P.clip[v] = to_screen(pos, P);
	P.clip_xy[v] = make_float2(pos.x, pos.y);
}

/* primitive 2t+k (k = 1: the second triangle the near clipper makes of input triangle t, rare) */
__device__ __forceinline__ Prim* prim_at(const DrawParams& P, uint32_t pid)
{
	return ((pid & 1u) ? P.prims2 : P.prims) + (pid >> 1);
}

/* The three snapped vertices of input triangle t and their varying records: stream positions 3t,
 * 3t+1, 3t+2 (a trailing partial triangle is still drawn, swgl.c:3611), through the element buffer
 * for glDrawElements; an index past the shaded range reads as a zero clip-space vertex. */
template <bool WAIT_FOR_VERTEX_KERNEL = false>
__device__ __forceinline__ void tri_vertices(const DrawParams& P, uint32_t t, float4& p0, float4& p1, float4& p2,
                                             uint32_t& s0, uint32_t& s1, uint32_t& s2)
{
	s0 = 3u * t; s1 = s0 + 1u; s2 = s0 + 2u;
	if (P.ibo)
	{
		const unsigned long long at = (unsigned long long)(long long)P.first + s0;
		s0 = (at < P.ibo_count) ? __ldg(P.ibo + at) : 0xffffffffu;
		s1 = (at + 1 < P.ibo_count) ? __ldg(P.ibo + at + 1) : 0xffffffffu;
		s2 = (at + 2 < P.ibo_count) ? __ldg(P.ibo + at + 2) : 0xffffffffu;
	}
	/* k_setup_bin is launched while the vertex kernel drains (programmatic dependent launch): the
	 * indices above do not depend on it, everything below does */
	if (WAIT_FOR_VERTEX_KERNEL) pdl_wait();
	const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	p0 = (s0 < P.n_shade) ? P.clip[s0] : to_screen(zero, P);
	p1 = (s1 < P.n_shade) ? P.clip[s1] : to_screen(zero, P);
	p2 = (s2 < P.n_shade) ? P.clip[s2] : to_screen(zero, P);
}

/* A tile-list entry is (primitive id << 1) | has_record.  Short, unclipped primitives of a draw that
 * goes to the warp rasteriser have no record: the consumer gathers the three vertices through the
 * element buffer again (the per-vertex data is shared by the neighbouring triangles and stays in
 * cache) instead of the set-up kernel writing 64 bytes per triangle. */
/* Same, as three vertex addresses: the vertex loads and everything behind them are shared by both
 * kinds of entry.  A record-less entry always has its three indices inside the shaded range (the
 * set-up kernel writes a record otherwise). */
struct PrimRef { const float4* a; const float4* b; const float4* c; uint32_t vid0, vid1, vid2, band; };
__device__ __forceinline__ PrimRef prim_ref(const DrawParams& P, uint32_t entry)
{
	PrimRef r;
	if (entry & 1u)
	{
		const Prim* q = prim_at(P, entry >> 1);
		const uint4 m = *(const uint4*)q->vid;
		r.a = q->v; r.b = q->v + 1; r.c = q->v + 2;
		r.vid0 = m.x; r.vid1 = m.y; r.vid2 = m.z; r.band = m.w;
		return r;
	}
	const uint32_t s = 3u * (entry >> 2);
	r.vid0 = s; r.vid1 = s + 1u; r.vid2 = s + 2u;
	if (P.ibo)
	{
		const uint32_t* ix = P.ibo + ((unsigned long long)(long long)P.first + s);
		r.vid0 = __ldg(ix); r.vid1 = __ldg(ix + 1); r.vid2 = __ldg(ix + 2);
	}
	r.a = P.clip + r.vid0; r.b = P.clip + r.vid1; r.c = P.clip + r.vid2;
	r.band = 0xffffffffu;
	return r;
}

__device__ __forceinline__ Prim load_prim(const DrawParams& P, uint32_t entry)
{
	if (entry & 1u) return *prim_at(P, entry >> 1);
	Prim r;
	tri_vertices(P, entry >> 2, r.v[0], r.v[1], r.v[2], r.vid[0], r.vid[1], r.vid[2]);
	r.band = 0xffffffffu;
	return r;
}

/* MipMapLevel of the primitive behind a list entry (swgl.c:3316: computed from the snapped vertices in
 * submission order, before the y sort); only draws with the mip_lod option call it */
__device__ __noinline__ float prim_lod(const DrawParams& P, uint32_t entry)
{
	const Prim q = load_prim(P, entry);
	return mip_level(q.v[0].x, q.v[0].y, q.v[1].x, q.v[1].y, q.v[2].x, q.v[2].y);
}

/* sort-first: does this rank own tile row `tr`? */
__device__ __forceinline__ bool owns_tile_row(const DrawParams& P, uint32_t tr)
{
	/* ownership is decided per band of 32 framebuffer rows whatever the tile height */
	return P.n_ranks <= 1 || (((((tr << P.th_shift) >> 5) / P.band_rows) % P.n_ranks) == P.rank);
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
	float lod;                    /* generic shape, mip_lod draws: MipMapLevel of the primitive (swgl.c:3316) */
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
#ifdef SWGL_JIT
	if (FS == SWFS_JIT) return jit_fragment(P, f);
#endif
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
	ir_execute(P.fs_ops, P.fs_nops, V, P, f.lod);
	float o[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
	for (uint32_t k = 0; k < P.out_floats; k++) o[k] = __uint_as_float(V[P.out_word + k]);
	return make_float4(o[0], o[1], o[2], o[3]);
}

/* Walk state (x0, x1, s1, switched) of a primitive on entering row y_in of tile row `ty`
 * (swgl.c:3350-3356, 3466-3471): from the band entry for tall primitives, by replaying the
 * additions from the first row for short ones (at most two tile heights of float additions). */
__device__ __forceinline__ void walk_to_row(const DrawParams& P, const TriWalk& w, uint32_t band, uint32_t ty, int y_in,
                                            float& x0, float& x1, float& s1, bool& switched)
{
	if (y_in == w.ys) { x0 = w.c0x; x1 = w.c0x; s1 = w.s1; switched = false; return; }
	if (band != 0xffffffffu)
	{
		const uint32_t tr_hi = (uint32_t)(P.ytop - w.ys) >> P.th_shift;
		const BandEntry be = P.bands[band + (tr_hi - ty)];
		x0 = be.x0; x1 = be.x1;
		switched = (float)y_in >= w.c1y;
		s1 = switched ? w.s2 : w.s1;
		return;
	}
	x0 = w.c0x; x1 = w.c0x; s1 = w.s1; switched = false;
	for (int y = w.ys; y < y_in; y++)
	{
		if (!switched && (float)y + 1.0f >= w.c1y) { switched = true; s1 = w.s2; x1 = w.c1x; }
		x0 += w.s0; x1 += s1;
	}
}

/* blend with the destination unpacked through the byte/255.0f table */
__device__ __forceinline__ float4 clamp_color(float4 o)
{
	/* swgl.c:3428-3431, ternary MIN/MAX: NaN -> 0 */
	o.x = RMIN(RMAX(o.x, 0.0f), 1.0f);
	o.y = RMIN(RMAX(o.y, 0.0f), 1.0f);
	o.z = RMIN(RMAX(o.z, 0.0f), 1.0f);
	o.w = RMIN(RMAX(o.w, 0.0f), 1.0f);
	return o;
}

/* r, g, b, a already clamped; the destination unpacked arithmetically (byte_over_255).  The warp rasteriser's
 * choice: the byte / 255 table cost four shared-memory reads per blend with random, bank-conflicting addresses,
 * and that kernel's busiest unit is the shared-memory data pipe (ncu: l1tex 65 %), not the FMA pipe (27 %). */
__device__ __forceinline__ uint32_t blend_pack_arith(float r, float g, float b, float a, uint32_t cur)
{
	const float cr = byte_over_255(cur >> 24), cg = byte_over_255((cur >> 16) & 0xFFu);
	const float cb = byte_over_255((cur >> 8) & 0xFFu), ca = byte_over_255(cur & 0xFFu);
	r = cr + a * (r - cr);
	g = cg + a * (g - cg);
	b = cb + a * (b - cb);
	a = ca + a * (a - ca);
	uint32_t word = 0;
	word |= (uint32_t)(int)(r * 255.0f) << 24;
	word |= (uint32_t)(int)(g * 255.0f) << 16;
	word |= (uint32_t)(int)(b * 255.0f) << 8;
	word |= (uint32_t)(int)(a * 255.0f);
	return word;
}

/* r, g, b, a already clamped */
__device__ __forceinline__ uint32_t blend_pack_lut(float r, float g, float b, float a, uint32_t cur, const float* lut)
{
	const float cr = lut[(cur >> 24) & 0xFF], cg = lut[(cur >> 16) & 0xFF];
	const float cb = lut[(cur >> 8) & 0xFF], ca = lut[cur & 0xFF];
	r = cr + a * (r - cr);
	g = cg + a * (g - cg);
	b = cb + a * (b - cb);
	a = ca + a * (a - ca);
	uint32_t word = 0;
	word |= (uint32_t)(int)(r * 255.0f) << 24;
	word |= (uint32_t)(int)(g * 255.0f) << 16;
	word |= (uint32_t)(int)(b * 255.0f) << 8;
	word |= (uint32_t)(int)(a * 255.0f);
	return word;
}


#endif
