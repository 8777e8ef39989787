/*
 * oracle/swgl_oracle.c -- TEST INFRASTRUCTURE (checker), never linked into the product.
 *
 * CPU restatement of the draw-call hot path of waternine9/swgl.  Every function cites the
 * reference lines (swgl.c) it restates.  All arithmetic is IEEE binary32, evaluated left
 * to right exactly as the reference writes it; build with -ffp-contract=off (oracle/Makefile)
 * so no multiply-add is fused.  float->int conversions are plain C casts: on x86-64 they
 * compile to cvttss2si, the same instruction the compiled reference uses.
 *
 * Parity: PINNED against the compiled reference, see swgl_oracle.h.
 */
#include "swgl_oracle.h"

#include <string.h>
#include <time.h>

/* swgl.c:15-16 -- ternary macros; with a NaN operand they yield the second operand. */
#define RMIN(x, y) (((x) < (y)) ? (x) : (y))
#define RMAX(x, y) (((x) > (y)) ? (x) : (y))

typedef struct { float x, y, z, w; } vec4;

typedef struct
{
	vec4  pos;      /* clip-space gl_Position, later (X, Y, z_clip, w_clip) */
	float var[4];   /* the one captured varying (float/vec2/vec3/vec4) */
} overtex;

typedef struct
{
	float attr_pos[4];  /* layout variable fed by attribute 0: persists across vertices */
	float attr_var[4];  /* layout variable fed by attribute 1 */
	float D[16];        /* mat4 uniform storage after glUniformMatrix4fv */
} vs_state;

/* glUniformMatrix4fv (swgl.c:3895-3928): `transpose = !transpose`; GL_FALSE therefore stores
 * the transposed array, D[r*4+c] = value[c*4+r]. */
static void store_matrix(const swglo_shader* s, float* D)
{
	int transpose = !s->matrix_transpose;
	if (!transpose) memcpy(D, s->matrix_value, sizeof(float) * 16);
	else
		for (int r = 0; r < 4; r++)
			for (int c = 0; c < 4; c++)
				D[r * 4 + c] = s->matrix_value[c * 4 + r];
}

/* Attribute fetch (swgl.c:3618-3639) + vertex shader + varying capture (swgl.c:3641-3666).
 * mat4 variable load reads m30 from Data[10] (swgl.c:2231-2235); MatMulMat4Vec adds left to
 * right (swgl.c:758-768). */
static void run_vertex(const swglo_shader* s, vs_state* st, const uint8_t* vbo, size_t vbo_bytes,
                       uint64_t vertex, overtex* out)
{
	size_t base = (size_t)vertex * s->stride;
	int np = s->pos_size > 4 ? 4 : s->pos_size;
	int nv = s->var_size > 4 ? 4 : s->var_size;
	if (np > 0 && base + s->pos_offset + (size_t)np * 4 <= vbo_bytes)
		memcpy(st->attr_pos, vbo + base + s->pos_offset, (size_t)np * 4);
	if (nv > 0 && base + s->var_offset + (size_t)nv * 4 <= vbo_bytes)
		memcpy(st->attr_var, vbo + base + s->var_offset, (size_t)nv * 4);

	float px = st->attr_pos[0], py = st->attr_pos[1], pz = st->attr_pos[2], pw = st->attr_pos[3];
	if (s->use_matrix)
	{
		const float* D = st->D;
		float m30 = D[10], m31 = D[13], m32 = D[14], m33 = D[15]; /* the load quirk */
		out->pos.x = D[0] * px + D[1] * py + D[2] * pz + D[3] * pw;
		out->pos.y = D[4] * px + D[5] * py + D[6] * pz + D[7] * pw;
		out->pos.z = D[8] * px + D[9] * py + D[10] * pz + D[11] * pw;
		out->pos.w = m30 * px + m31 * py + m32 * pz + m33 * pw;
	}
	else
	{
		out->pos.x = px; out->pos.y = py; out->pos.z = pz; out->pos.w = pw;
	}
	for (int k = 0; k < 4; k++) out->var[k] = (k < s->var_comps) ? st->attr_var[k] : 0.0f;
}

/* IntersectNearPlane (swgl.c:455-466) + InterpolateExValue (swgl.c:468-497). */
static overtex intersect_near(const overtex* a, const overtex* b, int comps)
{
	overtex r;
	float t = (a->pos.z + a->pos.w) / (a->pos.w - b->pos.w + a->pos.z - b->pos.z);
	r.pos.x = a->pos.x + t * (b->pos.x - a->pos.x);
	r.pos.y = a->pos.y + t * (b->pos.y - a->pos.y);
	r.pos.z = a->pos.z + t * (b->pos.z - a->pos.z);
	r.pos.w = a->pos.w + t * (b->pos.w - a->pos.w);
	for (int k = 0; k < 4; k++)
		r.var[k] = (k < comps) ? a->var[k] + t * (b->var[k] - a->var[k]) : 0.0f;
	return r;
}

/* ClipTriangleAgainstNearPlane (swgl.c:499-697): near plane only, inside iff z >= -w. */
static int clip_near(const overtex in[3], overtex out[2][3], int comps)
{
	const overtex* ins[3]; int n_in = 0;
	const overtex* outs[3]; int n_out = 0;
	for (int i = 0; i < 3; i++)
	{
		if (in[i].pos.z >= -in[i].pos.w) ins[n_in++] = &in[i];
		else outs[n_out++] = &in[i];
	}
	if (n_in == 0) return 0;
	if (n_in == 3)
	{
		out[0][0] = in[0]; out[0][1] = in[1]; out[0][2] = in[2];
		return 1;
	}
	if (n_in == 1)
	{
		out[0][0] = *ins[0];
		out[0][1] = intersect_near(ins[0], outs[0], comps);
		out[0][2] = intersect_near(ins[0], outs[1], comps);
		return 1;
	}
	out[0][0] = *ins[0];
	out[0][1] = *ins[1];
	out[0][2] = intersect_near(ins[0], outs[0], comps);
	out[1][0] = *ins[1];
	out[1][1] = out[0][2];
	out[1][2] = intersect_near(ins[1], outs[0], comps);
	return 2;
}

/* texture() without mip maps (swgl.c:2483-2565). */
static void sample_nearest(const swglo_shader* s, float u, float v, float rgba[4])
{
	int tx = u * s->tex_w;
	int ty = v * s->tex_h;
	if (s->wrap_s_repeat) tx %= s->tex_w;
	tx = RMIN(RMAX(tx, 0), s->tex_w - 1);
	if (s->wrap_t_repeat) ty %= s->tex_h;
	ty = RMIN(RMAX(ty, 0), s->tex_h - 1);
	const float* p = s->tex + (size_t)s->tex_fpp * ((size_t)tx + (size_t)ty * s->tex_w);
	rgba[0] = rgba[1] = rgba[2] = rgba[3] = 0.0f; /* `{ GLSL_VEC4 }` zero-fills (swgl.c:2561) */
	if (s->tex_fpp >= 1) rgba[0] = p[0];
	if (s->tex_fpp >= 2) rgba[1] = p[1];
	if (s->tex_fpp >= 3) rgba[2] = p[2];
	if (s->tex_fpp == 4) rgba[3] = p[3];
}

/* MipMapLevel (swgl.c:3314): a global that every DrawTriangle call sets and texture() reads whatever the primitive --
 * a GL_POINTS draw samples with what the last triangle left. */
static float g_mip_level = 0.0f;

/* rsqrt (swgl.c:3238-3254) with the pun through a 32-bit integer: the defined variant (see swgl_oracle.h). */
static float rsqrt_defined(float number)
{
	int32_t i;
	float x2, y;
	const float threehalfs = 1.5F;
	x2 = number * 0.5F;
	y = number;
	memcpy(&i, &y, 4);
	i = 0x5f3759df - (i >> 1);
	memcpy(&y, &i, 4);
	y = y * (threehalfs - (x2 * y * y));
	y = y * (threehalfs - (x2 * y * y));
	y = y * (threehalfs - (x2 * y * y));
	return y;
}

/* DistBetweenPointAndLine (swgl.c:3299-3312) and MipMapLevel = 40 / distance (swgl.c:3316). */
static float mip_level(float x1, float y1, float x2, float y2, float x3, float y3)
{
	float m = (y2 - y1) / RMAX(x2 - x1, 1.0f);
	float c = y1 - m * x1;
	float distance = (m * x3 - y3 + c);
	if (distance < 0.0f) distance *= -1.0f;
	distance *= rsqrt_defined(m * m + 1);
	return 40.0f / distance;
}

static void texel_of(const float* data, int w, int h, int fpp, int rep_s, int rep_t, float u, float v, float rgba[4])
{
	int tx = u * w;
	int ty = v * h;
	if (rep_s) tx %= w;
	tx = RMIN(RMAX(tx, 0), w - 1);
	if (rep_t) ty %= h;
	ty = RMIN(RMAX(ty, 0), h - 1);
	const float* p = data + (size_t)fpp * ((size_t)tx + (size_t)ty * w);
	rgba[0] = rgba[1] = rgba[2] = rgba[3] = 0.0f;
	if (fpp >= 1) rgba[0] = p[0];
	if (fpp >= 2) rgba[1] = p[1];
	if (fpp >= 3) rgba[2] = p[2];
	if (fpp == 4) rgba[3] = p[3];
}

/* texture() with mip levels (swgl.c:2516-2595): level L > 0 reads MipMaps[MIN(L, n-1)] and MipMaps[MIN(L-1, n-1)]
 * (the float is truncated to the vector index) and mixes them with T = 1 - (L - (int)L). */
static void sample_lod(const swglo_shader* s, float u, float v, float level, float rgba[4])
{
	if (!(s->n_mips > 0 && level > 0.0f)) { sample_nearest(s, u, v, rgba); return; }
	const int k0 = (int)RMIN(level, s->n_mips - 1);
	const int k1 = (int)RMIN(level - 1, s->n_mips - 1);
	float hi[4];
	texel_of(s->mip_data + s->mip_off[k0], s->mip_w[k0], s->mip_h[k0], s->tex_fpp, s->wrap_s_repeat, s->wrap_t_repeat, u, v, rgba);
	texel_of(s->mip_data + s->mip_off[k1], s->mip_w[k1], s->mip_h[k1], s->tex_fpp, s->wrap_s_repeat, s->wrap_t_repeat, u, v, hi);
	float T = level - (int)level;
	T = 1.0f - T;
	if (s->tex_fpp >= 1) rgba[0] = rgba[0] + T * (hi[0] - rgba[0]);
	if (s->tex_fpp >= 2) rgba[1] = rgba[1] + T * (hi[1] - rgba[1]);
	if (s->tex_fpp >= 3) rgba[2] = rgba[2] + T * (hi[2] - rgba[2]);
	if (s->tex_fpp == 4) rgba[3] = rgba[3] + T * (hi[3] - rgba[3]);
}

int swglo_build_mipmaps(const float* base, int32_t w, int32_t h, int32_t fpp,
                        float* out, int32_t* off, int32_t* lw, int32_t* lh, size_t* total_floats)
{
	int cw = w / 2, ch = h / 2, n = 0;
	size_t floats = 0;
	const float* prev = base;
	while (cw + ch > 4) /* swgl.c:2134 */
	{
		if (out)
		{
			float* cur = out + floats;
			off[n] = (int32_t)floats; lw[n] = cw; lh[n] = ch;
			for (int y = 0; y < ch; y++)
				for (int x = 0; x < cw; x++)
				{
					float* cp = cur + fpp * (x + y * cw);
					for (int k = 0; k < fpp; k++) cp[k] = 0.0f;
					for (int sy = 0; sy < 2; sy++)
						for (int sx = 0; sx < 2; sx++)
						{
							/* the row stride of the previous level is taken as 2 * CurWidth (swgl.c:2151) */
							const float* pp = prev + fpp * ((x * 2 + sx) + (y * 2 + sy) * cw * 2);
							for (int k = 0; k < fpp; k++) cp[k] += pp[k];
						}
					for (int k = 0; k < fpp; k++) cp[k] /= 4.0f;
				}
			prev = cur;
		}
		floats += (size_t)cw * ch * fpp;
		n++;
		cw /= 2; ch /= 2;
	}
	if (total_floats) *total_floats = floats;
	return n;
}

/* DrawTriangle (swgl.c:3314-3473). `o` = the three vertices in submission order with
 * pos = ((float)X, (float)Y, z_clip, w_clip). */
static void draw_triangle(const swglo_target* t, const swglo_shader* s, const overtex o[3],
                          swglo_stats* stats)
{
	const uint32_t W = t->width, H = t->height;
	const int32_t VX = t->vx, VY = t->vy;
	const uint32_t VW = t->vw, VH = t->vh;

	/* swgl.c:3316: the level of detail of the whole triangle, from the vertices as submitted */
	const float level = s->lod ? mip_level(o[0].pos.x, o[0].pos.y, o[1].pos.x, o[1].pos.y, o[2].pos.x, o[2].pos.y) : 0.0f;
	if (s->lod) g_mip_level = level;

	/* sort a copy by y with the reference's three compare-swaps (swgl.c:3323-3342) */
	vec4 c[3] = { o[0].pos, o[1].pos, o[2].pos };
	if (c[0].y > c[2].y) { vec4 tmp = c[0]; c[0] = c[2]; c[2] = tmp; }
	if (c[0].y > c[1].y) { vec4 tmp = c[0]; c[0] = c[1]; c[1] = tmp; }
	if (c[1].y > c[2].y) { vec4 tmp = c[1]; c[1] = c[2]; c[2] = tmp; }

	if (c[0].y >= VY + VH) return; /* int + uint32 -> uint32 -> float (swgl.c:3344) */

	float s0 = (c[2].x - c[0].x) / RMAX(c[2].y - c[0].y, 1.0f); /* swgl.c:3346-3348 */
	float s1 = (c[1].x - c[0].x) / RMAX(c[1].y - c[0].y, 1.0f);
	float s2 = (c[2].x - c[1].x) / RMAX(c[2].y - c[1].y, 1.0f);

	float y = RMAX(c[0].y, VY); /* swgl.c:3350: x is NOT advanced when the top is clipped */
	float x0 = c[0].x;
	float x1 = x0;
	uint8_t switched = 0;

	const vec4 a = o[0].pos, b = o[1].pos, cc = o[2].pos;

	for (; y < RMIN(c[2].y, VY + VH); y++, x0 += s0, x1 += s1) /* swgl.c:3356 */
	{
		for (int x = RMAX(RMIN(x0, x1), VX); x < RMIN(RMAX(x0, x1), VX + VW); x++) /* 3358 */
		{
			if (x < 0) continue;
			if ((uint32_t)x >= W) break;
			if (stats) stats->tested++;

			/* Barycentric (swgl.c:3256-3268) on the unsorted, snapped vertices */
			float px = x, py = y;
			float v0x = b.x - a.x, v0y = b.y - a.y;
			float v1x = cc.x - a.x, v1y = cc.y - a.y;
			float v2x = px - a.x, v2y = py - a.y;
			float d00 = v0x * v0x + v0y * v0y;
			float d01 = v0x * v1x + v0y * v1y;
			float d11 = v1x * v1x + v1y * v1y;
			float d20 = v2x * v0x + v2y * v0y;
			float d21 = v2x * v1x + v2y * v1y;
			float denom = d00 * d11 - d01 * d01;
			float bv = (d11 * d20 - d01 * d21) / denom;
			float bw = (d00 * d21 - d01 * d20) / denom;
			float bu = 1.0f - bv - bw;

			/* perspective correction (swgl.c:3367-3380) */
			float uc = bu / a.w, vc = bv / b.w, wc = bw / cc.w;
			float sum = uc + vc + wc;
			uc /= sum; vc /= sum; wc /= sum;

			float z = (a.z * uc + b.z * vc + cc.z * wc); /* swgl.c:3382 */

			/* swgl.c:3386: unsigned arithmetic; MAX(0, unsigned) is a no-op */
			uint32_t row = (VH - (uint32_t)((int)y - VY + 1)) + (uint32_t)VY;
			row = RMIN(H - 1, row);
			uint32_t idx = (uint32_t)x + row * W;

			float* curz = &t->depth[idx];
			if (*curz == 0.0f || *curz >= z) /* swgl.c:3387 */
			{
				*curz = z;
				if (stats) stats->shaded++;

				/* varying interpolation (swgl.c:3394-3406, 3270-3297) */
				float vr[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
				for (int k = 0; k < s->var_comps; k++)
					vr[k] = o[0].var[k] * uc + o[1].var[k] * vc + o[2].var[k] * wc;

				float out[4];
				if (s->fs_mode == 1) { if (s->lod) sample_lod(s, vr[0], vr[1], level, out); else sample_nearest(s, vr[0], vr[1], out); }
				else { out[0] = vr[0]; out[1] = vr[1]; out[2] = vr[2]; out[3] = vr[3]; }

				float r = RMIN(RMAX(out[0], 0.0f), 1.0f); /* swgl.c:3428-3431 */
				float g = RMIN(RMAX(out[1], 0.0f), 1.0f);
				float bl = RMIN(RMAX(out[2], 0.0f), 1.0f);
				float al = RMIN(RMAX(out[3], 0.0f), 1.0f);

				uint32_t cur = t->color[idx];
				float cr = ((cur >> 24) & 0xFF) / 255.0f; /* swgl.c:3434-3437 */
				float cg = ((cur >> 16) & 0xFF) / 255.0f;
				float cb = ((cur >> 8) & 0xFF) / 255.0f;
				float ca = (cur & 0xFF) / 255.0f;

				r = cr + al * (r - cr); /* swgl.c:3439-3442 */
				g = cg + al * (g - cg);
				bl = cb + al * (bl - cb);
				al = ca + al * (al - ca);

				uint32_t word = 0; /* swgl.c:3455-3459 */
				word |= (uint32_t)((int)(r * 255)) << 24;
				word |= (uint32_t)((int)(g * 255)) << 16;
				word |= (uint32_t)((int)(bl * 255)) << 8;
				word |= (uint32_t)((int)(al * 255));
				t->color[idx] = word;
			}
		}
		if (y + 1 >= c[1].y && !switched) /* swgl.c:3466-3471 */
		{
			switched = 1;
			s1 = s2;
			x1 = c[1].x;
		}
	}
}

static void draw_common(const swglo_target* t, const swglo_shader* s,
                        const uint8_t* vbo, size_t vbo_bytes, const uint32_t* indices,
                        int32_t first, uint32_t count, swglo_stats* stats)
{
	vs_state st;
	memset(&st, 0, sizeof(st));
	store_matrix(s, st.D);

	/* `for (int i = first; i < first + count; i += 3)` (swgl.c:3611): the comparison is
	 * unsigned, and a trailing partial triangle is still drawn. */
	for (int i = first; (uint32_t)i < (uint32_t)first + count; i += 3)
	{
		overtex tri[3];
		for (int j = 0; j < 3; j++)
		{
			uint64_t v = indices ? indices[(uint32_t)(i + j)] : (uint64_t)(uint32_t)(i + j);
			run_vertex(s, &st, vbo, vbo_bytes, v, &tri[j]);
		}
		if (stats) stats->triangles_in++;

		overtex clipped[2][3];
		int n = clip_near(tri, clipped, s->var_comps);
		for (int k = 0; k < n; k++)
		{
			overtex scr[3];
			for (int j = 0; j < 3; j++)
			{
				/* swgl.c:3685-3691: VW/2 is unsigned integer division, then int <- float */
				const vec4 p = clipped[k][j].pos;
				int X = p.x / p.w * (t->vw / 2) + (t->vw / 2) + t->vx;
				int Y = p.y / p.w * (t->vh / 2) + (t->vh / 2) + t->vy;
				scr[j] = clipped[k][j];
				scr[j].pos.x = X;
				scr[j].pos.y = Y;
			}
			if (stats) stats->prims_out++;
			draw_triangle(t, s, scr, stats);
		}
	}
}

/* glDrawArrays(GL_POINTS, first, count) (swgl.c:3496-3608): one pixel per vertex.  X is scaled with ViewportHeight / 2
 * (3531), the varyings are the vertex's own, depth is written without a test and without the row flip of the triangle
 * path (3582-3583), colour without a blend (3599-3604). */
void swglo_draw_points(const swglo_target* t, const swglo_shader* s,
                       const uint8_t* vbo, size_t vbo_bytes, int32_t first, uint32_t count)
{
	vs_state st;
	memset(&st, 0, sizeof(st));
	store_matrix(s, st.D);
	for (int i = first; (uint32_t)i < (uint32_t)first + count; i++)
	{
		overtex v;
		run_vertex(s, &st, vbo, vbo_bytes, (uint64_t)(uint32_t)i, &v);
		int X = v.pos.x / v.pos.w * (t->vh / 2) + (t->vw / 2) + t->vx;
		int Y = v.pos.y / v.pos.w * (t->vh / 2) + (t->vh / 2) + t->vy;
		if (X < 0 || (uint32_t)X >= t->width) continue;
		if (Y < 0 || (uint32_t)Y >= t->height) continue;
		float out[4];
		if (s->fs_mode == 1) { if (s->lod) sample_lod(s, v.var[0], v.var[1], g_mip_level, out); else sample_nearest(s, v.var[0], v.var[1], out); }
		else { out[0] = v.var[0]; out[1] = v.var[1]; out[2] = v.var[2]; out[3] = v.var[3]; }
		float r = RMIN(RMAX(out[0], 0.0f), 1.0f);
		float g = RMIN(RMAX(out[1], 0.0f), 1.0f);
		float bl = RMIN(RMAX(out[2], 0.0f), 1.0f);
		float al = RMIN(RMAX(out[3], 0.0f), 1.0f);
		const uint32_t idx = (uint32_t)X + (uint32_t)Y * t->width;
		t->depth[idx] = v.pos.z;
		uint32_t word = 0;
		word |= (uint32_t)((int)(r * 255)) << 24;
		word |= (uint32_t)((int)(g * 255)) << 16;
		word |= (uint32_t)((int)(bl * 255)) << 8;
		word |= (uint32_t)((int)(al * 255));
		t->color[idx] = word;
	}
}

void swglo_draw_arrays(const swglo_target* t, const swglo_shader* s,
                       const uint8_t* vbo, size_t vbo_bytes,
                       int32_t first, uint32_t count, swglo_stats* stats)
{
	draw_common(t, s, vbo, vbo_bytes, NULL, first, count, stats);
}

void swglo_draw_elements(const swglo_target* t, const swglo_shader* s,
                         const uint8_t* vbo, size_t vbo_bytes,
                         const uint32_t* indices, uint32_t count, swglo_stats* stats)
{
	draw_common(t, s, vbo, vbo_bytes, indices, 0, count, stats);
}

/* glClearColor (swgl.c:3175-3181) + glClear (swgl.c:3183-3214): viewport rect, no Y flip. */
void swglo_clear(const swglo_target* t, uint32_t flags, float r, float g, float b, float a)
{
	r = RMIN(RMAX(r, 0.0f), 1.0f);
	g = RMIN(RMAX(g, 0.0f), 1.0f);
	b = RMIN(RMAX(b, 0.0f), 1.0f);
	a = RMIN(RMAX(a, 0.0f), 1.0f);
	uint32_t word = 0;
	word |= (uint32_t)(r * 255) << 24;
	word |= (uint32_t)(g * 255) << 16;
	word |= (uint32_t)(b * 255) << 8;
	word |= (uint32_t)(a * 255);
	/* MIN(ViewportY + ViewportHeight, Height) is an unsigned MIN assigned to an int bound */
	int y0 = RMAX(t->vy, 0), y1 = (int)RMIN((uint32_t)t->vy + t->vh, t->height);
	int x0 = RMAX(t->vx, 0), x1 = (int)RMIN((uint32_t)t->vx + t->vw, t->width);
	for (int y = y0; y < y1; y++)
		for (int x = x0; x < x1; x++)
		{
			if (flags & 1u) t->color[(size_t)y * t->width + x] = word;
			if (flags & 2u) t->depth[(size_t)y * t->width + x] = 0.0f;
		}
}

void swglo_texels_from_u8(const uint8_t* in, float* out, size_t n)
{
	for (size_t i = 0; i < n; i++) out[i] = in[i] / 255.0f; /* swgl.c:2116 */
}

uint64_t swglo_fnv1a64(const uint32_t* words, size_t n)
{
	uint64_t h = 1469598103934665603ull;
	for (size_t i = 0; i < n; i++) h = (h ^ words[i]) * 1099511628211ull;
	return h;
}

double swglo_timed_frame(const swglo_target* t, const swglo_shader* s,
                         const uint8_t* vbo, size_t vbo_bytes, const uint32_t* indices,
                         int32_t first, uint32_t count, float cr, float cg, float cb, float ca,
                         int reps, swglo_stats* stats)
{
	double best = 1e300;
	for (int r = 0; r < reps; r++)
	{
		struct timespec t0, t1;
		swglo_stats local;
		memset(&local, 0, sizeof(local));
		clock_gettime(CLOCK_MONOTONIC, &t0);
		swglo_clear(t, 3u, cr, cg, cb, ca);
		draw_common(t, s, vbo, vbo_bytes, indices, first, count, &local);
		clock_gettime(CLOCK_MONOTONIC, &t1);
		double dt = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
		if (dt < best) best = dt;
		if (stats) *stats = local;
	}
	return best;
}
