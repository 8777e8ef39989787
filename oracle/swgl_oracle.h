/*
 * oracle/swgl_oracle.h -- TEST INFRASTRUCTURE (checker), never linked into the product.
 *
 * Plain-C restatement of the reference's draw-call hot path (waternine9/swgl, swgl.c),
 * written from SURVEY.md appendix A with the reference line each step follows cited in
 * swgl_oracle.c.  Parity status: PINNED -- tests/test_oracle.py checks this restatement
 * bit for bit (colour words and depth bit patterns) against the compiled, unmodified
 * reference (oracle/_ref/libswgl_ref.so) and against golden hashes in tests/golden/.
 *
 * It knows only what the hot path needs: a flat vertex stream, a "shader description"
 * covering the shader shapes of the BASELINE configs, one texture (with its mip levels and the defined
 * level of detail, pinned against oracle/_ref/libswgl_ref_lod.so), GL_TRIANGLES and GL_POINTS, one framebuffer.
 */
#ifndef SWGL_ORACLE_H
#define SWGL_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
	uint32_t width, height;   /* framebuffer (swgl.c:3156-3166) */
	int32_t  vx, vy;          /* viewport   (swgl.c:3151-3154) */
	uint32_t vw, vh;
	uint32_t* color;          /* [height*width], word = R<<24|G<<16|B<<8|A, row 0 = top */
	float*    depth;          /* [height*width], 0.0f = empty */
} swglo_target;

typedef struct
{
	/* vertex stage */
	int32_t  use_matrix;      /* 0: gl_Position = aPos;  1: gl_Position = uM * aPos */
	float    matrix_value[16];/* the array handed to glUniformMatrix4fv(loc,1,transpose,value) */
	int32_t  matrix_transpose;/* the GLboolean handed to glUniformMatrix4fv */
	/* attribute layout (GL_FLOAT only, swgl.c:3624) */
	uint32_t stride;
	uint32_t pos_offset;  int32_t pos_size;   /* floats copied into the vec4 position input */
	uint32_t var_offset;  int32_t var_size;   /* floats copied into the varying source */
	int32_t  var_comps;       /* declared type of the varying: 1,2,3,4 floats (0 = none) */
	/* fragment stage */
	int32_t  fs_mode;         /* 0: FragColor = varying (vec4); 1: FragColor = texture(uTex, varying.xy) */
	const float* tex;         /* float texels as stored by glTexImage2D (swgl.c:2107-2118) */
	int32_t  tex_w, tex_h, tex_fpp;
	int32_t  wrap_s_repeat, wrap_t_repeat;
	/* glGenerateMipmap levels (swgl.c:2122-2173) and the per-triangle level of detail, the DEFINED variant: the
	 * reference's rsqrt() puns a 4-byte float through an 8-byte long (swgl.c:3240-3246, undefined behaviour); with
	 * lod != 0 the restatement uses the same code with a 32-bit pun, which is what oracle/_ref/libswgl_ref_lod.so
	 * (ref_shim.c, -DSWGLREF_DEFINED_RSQRT) is built with and what the library's "mip_lod" option selects.
	 * n_mips levels: level k is mip_w[k] x mip_h[k] texels at mip_data + mip_off[k] (floats).  0 = no chain. */
	int32_t  lod;
	int32_t  n_mips;
	const float*   mip_data;
	const int32_t* mip_off;
	const int32_t* mip_w;
	const int32_t* mip_h;
} swglo_shader;

typedef struct
{
	uint64_t triangles_in;    /* input triangles */
	uint64_t prims_out;       /* triangles after near clip that reach DrawTriangle */
	uint64_t tested;          /* Barycentric evaluations  (swgl.c:3365) */
	uint64_t shaded;          /* depth-test passes = FS runs (swgl.c:3408) */
} swglo_stats;

/* glClear (swgl.c:3183-3214) with glClearColor's clamp (swgl.c:3175-3181). */
void swglo_clear(const swglo_target* t, uint32_t flags, float r, float g, float b, float a);

/* glDrawArrays(GL_TRIANGLES, first, count) (swgl.c:3609-3710) over a raw vertex buffer. */
void swglo_draw_arrays(const swglo_target* t, const swglo_shader* s,
                       const uint8_t* vbo, size_t vbo_bytes,
                       int32_t first, uint32_t count, swglo_stats* stats);

/* glDrawArrays(GL_POINTS, first, count) (swgl.c:3496-3608).  Under `lod` the points sample with the level the last
 * swglo_draw_* triangle left (the reference's global MipMapLevel). */
void swglo_draw_points(const swglo_target* t, const swglo_shader* s,
                       const uint8_t* vbo, size_t vbo_bytes, int32_t first, uint32_t count);

/* Same, fetching vertex i of the stream as vbo[indices[i]] -- the *definition* of the
 * glDrawElements extension: glDrawArrays over the de-indexed stream. */
void swglo_draw_elements(const swglo_target* t, const swglo_shader* s,
                         const uint8_t* vbo, size_t vbo_bytes,
                         const uint32_t* indices, uint32_t count, swglo_stats* stats);

/* glGenerateMipmap (swgl.c:2122-2173): 2x2 box levels of a float image while CurWidth + CurHeight > 4.  Call with
 * out == NULL to get the number of levels and floats; with buffers to fill them (offsets in floats).  Returns the
 * number of levels. */
int swglo_build_mipmaps(const float* base, int32_t w, int32_t h, int32_t fpp,
                        float* out, int32_t* off, int32_t* lw, int32_t* lh, size_t* total_floats);

/* glTexImage2D's byte->float conversion (swgl.c:2116). out has n floats. */
void swglo_texels_from_u8(const uint8_t* in, float* out, size_t n);

/* 64-bit FNV-1a variant of SURVEY.md appendix C over 32-bit words. */
uint64_t swglo_fnv1a64(const uint32_t* words, size_t n);

/* Time `reps` clear+draw frames, return seconds of the best frame. */
double swglo_timed_frame(const swglo_target* t, const swglo_shader* s,
                         const uint8_t* vbo, size_t vbo_bytes, const uint32_t* indices,
                         int32_t first, uint32_t count, float cr, float cg, float cb, float ca,
                         int reps, swglo_stats* stats);

#ifdef __cplusplus
}
#endif

#endif /* SWGL_ORACLE_H */
