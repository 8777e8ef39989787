/*
 * swgl_glsl.h -- host-side front-end for the reference's GLSL subset.
 *
 * glCompileShader turns a source string into a swgl_shader: a variable table plus a typed
 * straight-line op list (swgl_ir.h).  The accepted language and its quirks are the
 * reference tokenizer's (swgl.c:791-1867): newline/tab deletion without
 * a separating blank, strict left-to-right folding of binary chains with no precedence,
 * `-` followed by a digit read as a sign, swizzle detection that stops at the first blank,
 * `layout (location = N) T name;` without an `in` keyword, and so on.
 */
#ifndef SWGL_GLSL_H
#define SWGL_GLSL_H

#include <stddef.h>
#include "swgl_ir.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SWGL_MAX_VARS 96
#define SWGL_NAME_LEN 64

typedef struct
{
	char     name[SWGL_NAME_LEN];
	int32_t  type;          /* SWT_* */
	uint8_t  is_uniform, is_layout, is_in, is_out, is_local;
	int32_t  location;      /* layout(location = N) */
	uint32_t word;          /* offset in the variable file */
	int32_t  copy_of;       /* main() assigns this variable exactly once, from variable copy_of; else -1 */
} swgl_var;

typedef struct
{
	int32_t  ok;                    /* 1: executable IR was produced */
	char     error[192];
	uint64_t id;                    /* unique per compilation (device-side cache key) */
	int32_t  n_vars;                /* globals first, in declaration order (gl_Position is 0) */
	int32_t  n_globals;
	swgl_var vars[SWGL_MAX_VARS];
	swgl_ir_code code;
	/* current contents of the variable file: uniforms written by glUniform*, rest zero.
	 * Like the reference's glslVariable storage it belongs to the shader object and is
	 * shared by every program the shader is attached to. */
	uint32_t image[SWGL_MAX_VAR_WORDS];
	/* shape recognition (SWVS_x / SWFS_x candidates; finalised at draw time) */
	int32_t  has_main;
	int32_t  n_statements;          /* executed statements in main */
	int32_t  simple_copies;         /* 1: every statement is `var = var` or the position forms below */
	int32_t  pos_kind;              /* 0 none, 1: gl_Position = attr, 2: gl_Position = mat4 * attr */
	int32_t  pos_attr_var, pos_mat_var;
	int32_t  fs_kind;               /* SWFS_* candidate for a fragment shader */
	int32_t  fs_in_var, fs_sampler_var;
	int32_t  fs_swz[2];             /* texture coordinate picks */
} swgl_shader;

/* Compile `source` (NUL terminated).  Always returns a shader object; check ->ok. */
swgl_shader* swgl_glsl_compile(const char* source);
void         swgl_glsl_free(swgl_shader* s);

int  swgl_glsl_find_var(const swgl_shader* s, const char* name);   /* first global with that name, or -1 */
/* Human-readable dump (tests, swglDebugShaderIR).  Returns bytes needed. */
size_t swgl_glsl_dump(const swgl_shader* s, char* buf, size_t len);

/* The reference's literal parsers, restated (swgl.c:18-88): literal VALUES must match. */
double swgl_glsl_atof(const char* str);
int    swgl_glsl_atoi(const char* str);

#ifdef __cplusplus
}
#endif

#endif /* SWGL_GLSL_H */
