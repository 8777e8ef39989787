/*
 * include/swgl.h -- drop-in replacement for the reference header (waternine9/swgl, swgl.h).
 *
 * Every declaration below keeps the reference's name, argument order, types and enum
 * *positions* (swgl.h:29-34 typedefs, swgl.h:40-81 positional enum, swgl.h:87-158
 * prototypes), so host code written against the reference compiles and links unchanged
 * against libswgl_b200.so.  The implementation behind it is the B200 draw-call path:
 * buffers, textures and the framebuffer live in device memory and glClear/glDrawArrays
 * run as CUDA kernels.
 *
 * Differences from the reference header, all ABI-neutral:
 *   - GL_COLOR_BUFFER_BIT / GL_DEPTH_BUFFER_BIT were *defined* as `const uint32_t` objects
 *     in the header (swgl.h:19-20), which breaks linking of two C translation units.  They
 *     are macros here with the same values (1 and 2).
 *   - New enumerators (GL_ELEMENT_ARRAY_BUFFER, GL_UNSIGNED_INT) are APPENDED after
 *     GL_TEXTURE7 so every existing enumerator keeps its value.
 *   - glDrawElements is a new entry point (the reference has no indexed draw).  It is
 *     defined as "glDrawArrays over the de-indexed vertex stream".
 * Further extensions (depth readback, sync, statistics, multi-GPU) are in swgl_b200.h.
 *
 * Limits the reference does not have (a draw outside them is skipped and swglGetLastError() says why):
 * framebuffers up to 65 504 x 8 184 pixels (2 047 tile columns, 1 023 tile rows), 2^30 triangles per draw,
 * 16 varying floats per vertex.  A viewport that leaves the framebuffer rows is drawn like the reference
 * draws it (rows outside fold onto the last row, swgl.c:3386) on a single device and on a device group
 * (swglSetDeviceCount); sort-first ranks of separate processes (swglSetStripe) skip such draws.
 */
#ifndef SOFTWARE_GL_H
#define SOFTWARE_GL_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* CONSTANTS (reference: swgl.h:19-23) */
#define GL_COLOR_BUFFER_BIT 0x1u
#define GL_DEPTH_BUFFER_BIT 0x2u

#define GL_TRUE 1
#define GL_FALSE 0

/* TYPES (reference: swgl.h:29-34) */
typedef uint32_t GLuint;
typedef uint32_t GLsizei; /* unsigned, as in the reference */
typedef int32_t GLint;
typedef char GLchar;
typedef uint8_t GLboolean;
typedef float GLfloat;

/* ENUMS (reference: swgl.h:40-81) -- positional values, NOT Khronos values */
typedef enum
{
	GL_VERTEX_SHADER,          /* 0 */
	GL_FRAGMENT_SHADER,        /* 1 */
	GL_COMPILE_STATUS,         /* 2 */
	GL_LINK_STATUS,            /* 3 */
	GL_ARRAY_BUFFER,           /* 4 */

	GL_STATIC_DRAW,            /* 5 */
	GL_STREAM_DRAW,            /* 6 */
	GL_DYNAMIC_DRAW,           /* 7 */

	GL_FLOAT,                  /* 8 */
	GL_INT,                    /* 9 */
	GL_UNSIGNED_BYTE,          /* 10 */

	GL_DEPTH_COMPONENT,        /* 11 */
	GL_DEPTH_STENCIL,          /* 12 */
	GL_RED,                    /* 13 */
	GL_RG,                     /* 14 */
	GL_RGB,                    /* 15 */
	GL_RGBA,                   /* 16 */

	GL_TRIANGLES,              /* 17 */
	GL_POINTS,                 /* 18 */
	GL_LINES,                  /* 19 */
	GL_REPEAT,                 /* 20 */
	GL_CLAMP,                  /* 21 */

	GL_TEXTURE_2D,             /* 22 */
	GL_TEXTURE_WRAP_S,         /* 23 */
	GL_TEXTURE_WRAP_T,         /* 24 */

	GL_TEXTURE0,               /* 25 */
	GL_TEXTURE1,
	GL_TEXTURE2,
	GL_TEXTURE3,
	GL_TEXTURE4,
	GL_TEXTURE5,
	GL_TEXTURE6,
	GL_TEXTURE7,               /* 32 */

	/* ---- extensions: appended, never inserted ---- */
	GL_ELEMENT_ARRAY_BUFFER,   /* 33 */
	GL_UNSIGNED_INT            /* 34 */
} GLenum;

/* NON-OPENGL HELPERS (reference: swgl.h:87-88) */
void glInit(GLsizei width, GLsizei height);
uint32_t* glGetFramePtr(void);

/* SHADERS (reference: swgl.h:94-102) */
GLuint glCreateShader(GLenum type);
void glShaderSource(GLuint shader, const GLchar* string);
void glCompileShader(GLuint shader);
void glDeleteShader(GLuint shader);

GLuint glCreateProgram(void);
void glAttachShader(GLuint program, GLuint shader);
void glLinkProgram(GLuint program);
void glUseProgram(GLuint program);

/* VERTEX ARRAYS (reference: swgl.h:108-111) -- one object per call, id written to *arrays */
GLuint glGenVertexArrays(GLsizei n, GLuint* arrays);
void glBindVertexArray(GLuint array);
void glVertexAttribPointer(GLuint index, GLint size, GLenum type, GLboolean normalized, GLsizei stride, const void* pointer);
void glEnableVertexAttribArray(GLuint index);

/* BUFFERS (reference: swgl.h:118-120) */
GLuint glGenBuffers(GLsizei n, GLuint* buffers);
void glBindBuffer(GLenum type, GLuint buffer);
void glBufferData(GLenum target, GLsizei size, const void* data, GLenum usage);

/* DRAW (reference: swgl.h:126-130) */
void glClearColor(GLfloat red, GLfloat green, GLfloat blue, GLfloat alpha);
void glClear(GLuint flags);
void glViewport(GLint x, GLint y, GLsizei width, GLsizei height);

void glDrawArrays(GLenum mode, GLint first, GLsizei count);

/* EXTENSION: indexed draw.  type must be GL_UNSIGNED_INT; `indices` is a byte offset into the
 * GL_ELEMENT_ARRAY_BUFFER bound to the current vertex array.  Result == glDrawArrays over
 * the stream vertex[k] = VBO[index[k]]. */
void glDrawElements(GLenum mode, GLsizei count, GLenum type, const void* indices);

/* TEXTURES (reference: swgl.h:136-141) */
void glGenTextures(GLsizei n, GLuint* textures);
void glActiveTexture(GLenum target);
void glBindTexture(GLenum target, GLuint texture);
void glTexParameteri(GLenum target, GLenum type, GLenum mode);
void glTexImage2D(GLenum target, GLint level, GLint internalformat, GLsizei width, GLsizei height, GLint border, GLenum format, GLenum type, const void* data);
void glGenerateMipmap(GLenum target);

/* UNIFORMS (reference: swgl.h:147-158) */
GLint glGetUniformLocation(GLuint program, const GLchar* name);

void glUniform1f(GLint location, GLfloat v0);
void glUniform2f(GLint location, GLfloat v0, GLfloat v1);
void glUniform3f(GLint location, GLfloat v0, GLfloat v1, GLfloat v2);
void glUniform4f(GLint location, GLfloat v0, GLfloat v1, GLfloat v2, GLfloat v3);

void glUniform1i(GLint location, GLint v0);

void glUniformMatrix2fv(GLint location, GLsizei count, GLboolean transpose, const GLfloat* value);
void glUniformMatrix3fv(GLint location, GLsizei count, GLboolean transpose, const GLfloat* value);
void glUniformMatrix4fv(GLint location, GLsizei count, GLboolean transpose, const GLfloat* value);

#ifdef __cplusplus
}
#endif

#endif /* SOFTWARE_GL_H */
