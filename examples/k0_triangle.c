/* A plain C application written against swgl.h, linked with libswgl_b200.so instead of compiling
 * swgl.c: SURVEY.md appendix C, known-answer scene K0 (64x48, one triangle, pass-through shaders).
 * Prints the number of drawn pixels and the FNV-1a hash of the colour buffer; the reference gives
 * 1008 and 5e2ecfac3685e7ef.
 *
 *   gcc -std=gnu11 examples/k0_triangle.c examples/k0_hash.c -Iinclude -Lswgl_b200 -lswgl_b200 \
 *       -Wl,-rpath,$PWD/swgl_b200 -o k0 && ./k0
 */
#include <stdio.h>

#include "swgl.h"

unsigned long long k0_fnv1a64(const uint32_t* words, unsigned long long n);   /* k0_hash.c, a second C unit that includes swgl.h */

static const char* VS =
    "layout (location = 0) vec4 aPos;\n"
    "layout (location = 1) vec4 aCol;\n"
    "out vec4 vCol;\n"
    "void main()\n{\n"
    "gl_Position = aPos;\n"
    "vCol = aCol;\n"
    "}\n";
static const char* FS =
    "in vec4 vCol;\n"
    "out vec4 FragColor;\n"
    "void main()\n{\n"
    "FragColor = vCol;\n"
    "}\n";

int main(void)
{
    const GLsizei W = 64, H = 48;
    static const float verts[3][8] = {
        { -0.8f, -0.8f, 0.5f, 1.0f,   1.0f, 0.0f, 0.0f, 1.0f },
        {  0.8f, -0.8f, 0.5f, 1.0f,   0.0f, 1.0f, 0.0f, 1.0f },
        {  0.0f,  0.8f, 0.5f, 1.0f,   0.0f, 0.0f, 1.0f, 1.0f },
    };
    glInit(W, H);
    GLuint vs = glCreateShader(GL_VERTEX_SHADER);
    glShaderSource(vs, VS);
    glCompileShader(vs);
    GLuint fs = glCreateShader(GL_FRAGMENT_SHADER);
    glShaderSource(fs, FS);
    glCompileShader(fs);
    GLuint prog = glCreateProgram();
    glAttachShader(prog, vs);
    glAttachShader(prog, fs);
    glLinkProgram(prog);
    glUseProgram(prog);

    GLuint vao = 0, vbo = 0;
    glGenVertexArrays(1, &vao);
    glBindVertexArray(vao);
    glGenBuffers(1, &vbo);
    glBindBuffer(GL_ARRAY_BUFFER, vbo);
    glBufferData(GL_ARRAY_BUFFER, sizeof(verts), verts, GL_STATIC_DRAW);
    glVertexAttribPointer(0, 4, GL_FLOAT, 0, 32, (const void*)0);
    glVertexAttribPointer(1, 4, GL_FLOAT, 0, 32, (const void*)16);

    glViewport(0, 0, W, H);
    glClearColor(0.0f, 0.0f, 0.0f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glDrawArrays(GL_TRIANGLES, 0, 3);

    const uint32_t* frame = glGetFramePtr();
    if (!frame) { fprintf(stderr, "no frame (no CUDA device?)\n"); return 2; }
    unsigned drawn = 0;
    for (unsigned i = 0; i < (unsigned)(W * H); i++) drawn += frame[i] != 0x000000FFu;
    printf("%u %016llx\n", drawn, k0_fnv1a64(frame, (unsigned long long)W * H));
    return 0;
}
