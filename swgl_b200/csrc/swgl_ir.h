/*
 * swgl_ir.h -- straight-line shader IR shared by the host front-end (swgl_glsl.c, plain C)
 * and the device back-end (swgl_dev.cu).
 *
 * The reference executes GLSL-subset shaders with a tree-walking interpreter
 * (swgl.c:2177-2868).  The language has no control flow, every
 * variable has a declared type and every operator's result type is a pure function of its
 * operand types, so a shader is compiled ONCE, at glCompileShader, into a typed, linear op
 * list.  Assignments whose types do not match are dropped at compile time -- exactly the
 * silent no-op of AssignToExVal (swgl.c:1898-1901).  Device code either runs a recognised
 * shader shape through a specialised __device__ functor or runs this op list through the
 * generic __device__ evaluator.
 */
#ifndef SWGL_IR_H
#define SWGL_IR_H

#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* value types (reference: glslType, swgl.c:271-283) */
enum
{
	SWT_FLOAT = 0,
	SWT_VEC2,
	SWT_VEC3,
	SWT_VEC4,
	SWT_INT,
	SWT_MAT2,
	SWT_MAT3,
	SWT_MAT4,
	SWT_SAMPLER2D,
	SWT_UNKNOWN
};

#ifndef __CUDACC_RTC__   /* host helpers */
/* storage words of a variable of each type (VerifyVar, swgl.c:1880-1892) */
static inline int swt_words(int t)
{
	switch (t)
	{
	case SWT_FLOAT: return 1;
	case SWT_VEC2: return 2;
	case SWT_VEC3: return 3;
	case SWT_VEC4: return 4;
	case SWT_INT: return 1;
	case SWT_SAMPLER2D: return 1;
	case SWT_MAT2: return 4;
	case SWT_MAT3: return 9;
	case SWT_MAT4: return 16;
	default: return 0;
	}
}
static inline int swt_is_mat(int t) { return t == SWT_MAT2 || t == SWT_MAT3 || t == SWT_MAT4; }
static inline int swt_mat_dim(int t) { return t == SWT_MAT2 ? 2 : t == SWT_MAT3 ? 3 : t == SWT_MAT4 ? 4 : 0; }
#endif

/*
 * Register model.  A vector temp carries what a glslExValue carries besides its matrices
 * (swgl.c:381-395): four floats and one int.  A matrix temp carries 16 floats, row-major,
 * dim x dim in the top-left.  Temps are allocated stack-like by the front-end.
 *
 * Variables live in a per-invocation "variable file" of 32-bit words; `a`/`dst` of LDV/STV
 * are word offsets into it.
 */
enum
{
	SWOP_NOP = 0,
	SWOP_LDV,    /* T[dst] <- n words at V[a] (xyzw, rest 0, i = 0)                  swgl.c:2179-2202 */
	SWOP_LDI,    /* T[dst].i <- V[a] (int / sampler), xyzw = 0                        swgl.c:2203-2212 */
	SWOP_LDM,    /* M[dst] <- matrix variable at V[a], dim n, WITH the load quirks    swgl.c:2213-2238 */
	SWOP_CONF,   /* T[dst] <- (imm as float, 0,0,0), i = 0                           swgl.c:2240-2246 */
	SWOP_CONI,   /* T[dst] <- i = imm, xyzw = 0                                       swgl.c:2247-2251 */
	SWOP_ZERO,   /* T[dst] <- all zero: the `{ GLSL_UNKNOWN }` value                  swgl.c:2279 */
	SWOP_STV,    /* V[dst .. dst+n) <- T[a].xyzw                                      swgl.c:1903-1927 */
	SWOP_STI,    /* V[dst] <- T[a].i                                                  swgl.c:1929-1932 */
	SWOP_STM,    /* V[dst ..) <- M[a], dim n, row-major                               swgl.c:1934-1973 */
	SWOP_ADD,    /* T[dst] <- T[a] + T[b]  (xyzw and i)                               swgl.c:2283-2287 */
	SWOP_SUB,    /*                                                                   swgl.c:2347-2351 */
	SWOP_MUL,    /*                                                                   swgl.c:2412-2416 */
	SWOP_DIV,    /* i only when T[b].i != 0                                           swgl.c:2475-2479 */
	SWOP_ADDM,   /* M[dst] <- M[a] + M[b], dim n (mat4: column-3 typo kept)           swgl.c:2289-2332 */
	SWOP_SUBM,   /*                                                                   swgl.c:2353-2396 */
	SWOP_MULMM,  /* M[dst] <- M[a] * M[b], dim n                                      swgl.c:699-756 */
	SWOP_MULMV,  /* T[dst] <- M[a] * T[b], dim n                                      swgl.c:758-789 */
	SWOP_TEX,    /* T[dst] <- texture(unit T[a].i, uv T[b].xy)                        swgl.c:2483-2598 */
	SWOP_SIN,    /* xyzw each                                                         swgl.c:2617-2634, 90-97 */
	SWOP_COS,    /*                                                                   swgl.c:2599-2616, 99-102 */
	SWOP_TAN,    /*                                                                   swgl.c:2635-2652, 104-107 */
	SWOP_MIN,    /* xyzw each, ternary-macro NaN behaviour                            swgl.c:2653-2673 */
	SWOP_MAX,    /*                                                                   swgl.c:2674-2694 */
	SWOP_SWZ,    /* T[dst] <- n comps picked from T[a]; imm = 2 bits per pick         swgl.c:2695-2765 */
	SWOP_CONS,   /* T[dst] <- n floats from T[a],T[b],T[c],T[d] (.x, or (float).i when the
	                arg is int-typed: bit k of imm2); c = imm & 0xffff, d = imm >> 16  swgl.c:2766-2829 */
	SWOP_ICONS,  /* T[dst].i <- (imm2 & 1) ? T[a].i : (int)T[a].x                     swgl.c:2830-2839 */
	SWOP_MOVM2T, /* T[dst] <- zero (a matrix value read where a vector is expected keeps xyzw = 0) */
	SWOP__COUNT
};

typedef struct
{
	uint8_t  op;
	uint8_t  n;
	uint16_t dst;
	uint16_t a;
	uint16_t b;
	uint32_t imm;
	uint32_t imm2;
} swgl_ir_op;

#define SWGL_MAX_OPS        192
#define SWGL_MAX_VAR_WORDS  128   /* variable file, words */
#define SWGL_MAX_TEMPS      24
#define SWGL_MAX_MTEMPS     6
#define SWGL_MAX_VARYING_FLOATS 16
#define SWGL_MAX_FETCH      16
#define SWGL_MAX_TEX_UNITS  8
/* glGenerateMipmap chains (swgldev_texture.mips): a table of SWGL_MIP_MAX_LEVELS entries x 4 fields, field-major
 * -- float offset of the level behind the table, width, height, floats per texel it was built with -- then the levels */
#define SWGL_MIP_MAX_LEVELS    32
#define SWGL_MIP_HEADER_WORDS  (4 * SWGL_MIP_MAX_LEVELS)

typedef struct swgl_ir_code
{
	uint32_t   n_ops;
	uint32_t   n_words;     /* size of the variable file */
	uint32_t   n_temps;
	uint32_t   n_mtemps;
	swgl_ir_op ops[SWGL_MAX_OPS];
} swgl_ir_code;

/* ---- recognised shader shapes (specialised device functors) ---- */
enum
{
	SWVS_GENERIC = 0, /* run the op list */
	SWVS_PASS,        /* gl_Position = <vec4 attribute>;  every linked varying = attribute copy */
	SWVS_MATRIX,      /* gl_Position = <mat4 uniform> * <vec4 attribute>; varyings = attribute copies */
	SWVS_JIT          /* device layer only: the op list compiled to a __device__ function at run time (swgl_jit.cpp) */
};
enum
{
	SWFS_GENERIC = 0, /* run the op list */
	SWFS_VARYING,     /* out = <vec4 varying> */
	SWFS_TEXTURE,     /* out = texture(<sampler uniform>, <varying>[.swizzle to vec2]) */
	SWFS_JIT          /* device layer only: the op list compiled to a __device__ function at run time (swgl_jit.cpp) */
};

#ifdef __cplusplus
}
#endif

#endif /* SWGL_IR_H */
