/*
 * examples/c4_multi_gpu.c -- a plain C application (no Python, no CUDA headers) that renders BASELINE
 * config 4 (4K, 1,002,528 small depth-tested triangles, indexed) on N GPUs of one box through the
 * reference's own GL-style API.  The only multi-GPU call is swglSetDeviceCount(N) before glInit.
 *
 *     gcc -std=gnu11 -O2 -ffp-contract=off examples/c4_multi_gpu.c -I include -L swgl_b200 -lswgl_b200 \
 *         -Wl,-rpath,$PWD/swgl_b200 -o c4_multi_gpu
 *     ./c4_multi_gpu <devices> [grid width height frames]
 *
 * prints:  <devices> <colour FNV-1a64> <covered pixels> <ms per end-to-end frame>
 *          and on stderr where the host thread spent that time (vertex upload, index upload, clear + draw, glGetFramePtr)
 *
 * The scene is swgl_b200/scenes.py's grid_mesh() restated in C (32-bit LCG s = s*1664525 + 1013904223,
 * rnd = (s >> 8) / 2^24, seed 12345; per vertex: w, jx, jy, z, r, g, b), so the colour hash of the default
 * arguments is the known-answer value of C4 rendered by the unmodified reference
 * (tests/golden/fullsize_kats.json: 56d0d4e1a9cd44f1) whatever the number of devices.
 *
 * An end-to-end frame here is what bench.py's `e2e` times: the vertex and index arrays re-specified from
 * page-locked host memory, glClear + glDrawElements, glGetFramePtr.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "swgl.h"
#include "swgl_b200.h"

static uint32_t lcg_state;
static float rnd(void)
{
	lcg_state = lcg_state * 1664525u + 1013904223u;
	return (float)(lcg_state >> 8) / 16777216.0f;
}

static double now_ms(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return 1e3 * (double)t.tv_sec + 1e-6 * (double)t.tv_nsec;
}

int main(int argc, char** argv)
{
	const int devices = argc > 1 ? atoi(argv[1]) : 1;
	const int grid = argc > 2 ? atoi(argv[2]) : 708;
	const int width = argc > 3 ? atoi(argv[3]) : 3840;
	const int height = argc > 4 ? atoi(argv[4]) : 2160;
	const int frames = argc > 5 ? atoi(argv[5]) : 20;
	const int g1 = grid + 1;
	const size_t nv = (size_t)g1 * g1, ni = (size_t)grid * grid * 6;
	const size_t vbytes = nv * 8 * sizeof(float), ibytes = ni * sizeof(uint32_t);

	swglSetDeviceCount(devices);
	glInit(width, height);
	if (swglGetLastError()[0]) { fprintf(stderr, "glInit failed\n"); return 2; }

	/* the application's arrays, in page-locked write-combined memory (swglHostAlloc): read by the copy engines only */
	float* verts = (float*)swglHostAlloc(vbytes, 1);
	uint32_t* idx = (uint32_t*)swglHostAlloc(ibytes, 1);
	if (!verts || !idx) { fprintf(stderr, "swglHostAlloc failed\n"); return 2; }
	lcg_state = 12345u;
	const float G = (float)grid;
	for (int j = 0; j < g1; j++)
		for (int i = 0; i < g1; i++)
		{
			float r[7];
			for (int k = 0; k < 7; k++) r[k] = rnd();
			const float w = 1.0f + (1.5f - 1.0f) * r[0];
			const float x = -0.98f + 1.96f * (float)i / G + (r[1] - 0.5f) * 0.6f / G;
			const float y = -0.98f + 1.96f * (float)j / G + (r[2] - 0.5f) * 0.6f / G;
			const float z = 0.1f + (0.9f - 0.1f) * r[3];
			float* v = verts + ((size_t)j * g1 + i) * 8;
			v[0] = x * w; v[1] = y * w; v[2] = z; v[3] = w;
			v[4] = r[4]; v[5] = r[5]; v[6] = r[6]; v[7] = 1.0f;
		}
	size_t o = 0;
	for (int qj = 0; qj < grid; qj++)
		for (int qi = 0; qi < grid; qi++)
		{
			const uint32_t a = (uint32_t)(qj * g1 + qi), b = a + 1u, c = a + (uint32_t)g1, d = c + 1u;
			idx[o++] = a; idx[o++] = b; idx[o++] = c; idx[o++] = b; idx[o++] = d; idx[o++] = c;
		}

	const GLuint vs = glCreateShader(GL_VERTEX_SHADER);
	glShaderSource(vs, "layout (location = 0) vec4 aPos;\nlayout (location = 1) vec4 aCol;\nout vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\n}\n");
	glCompileShader(vs);
	const GLuint fs = glCreateShader(GL_FRAGMENT_SHADER);
	glShaderSource(fs, "in vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = vCol;\n}\n");
	glCompileShader(fs);
	const GLuint prog = glCreateProgram();
	glAttachShader(prog, vs);
	glAttachShader(prog, fs);
	glLinkProgram(prog);
	glUseProgram(prog);

	GLuint vao = 0, vbo = 0, ebo = 0;
	glGenVertexArrays(1, &vao);
	glBindVertexArray(vao);
	glGenBuffers(1, &vbo);
	glBindBuffer(GL_ARRAY_BUFFER, vbo);
	glBufferData(GL_ARRAY_BUFFER, (GLsizei)vbytes, verts, GL_STATIC_DRAW);
	glGenBuffers(1, &ebo);
	glBindBuffer(GL_ELEMENT_ARRAY_BUFFER, ebo);
	glBufferData(GL_ELEMENT_ARRAY_BUFFER, (GLsizei)ibytes, idx, GL_STATIC_DRAW);
	glVertexAttribPointer(0, 4, GL_FLOAT, GL_FALSE, 32, (const void*)0);
	glVertexAttribPointer(1, 4, GL_FLOAT, GL_FALSE, 32, (const void*)16);
	glViewport(0, 0, (GLsizei)width, (GLsizei)height);
	glClearColor(0.0f, 0.0f, 0.0f, 1.0f);

	const uint32_t* frame = NULL;
	double t0 = 0.0, phase[4] = { 0.0, 0.0, 0.0, 0.0 };
	for (int f = 0; f < frames + 3; f++)
	{
		if (f == 3) { t0 = now_ms(); phase[0] = phase[1] = phase[2] = phase[3] = 0.0; }   /* three warm-up frames */
		const double a = now_ms();
		swglBufferRespecify(GL_ARRAY_BUFFER, (GLsizei)vbytes, verts);
		const double b = now_ms();
		swglBufferRespecify(GL_ELEMENT_ARRAY_BUFFER, (GLsizei)ibytes, idx);
		const double c = now_ms();
		glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
		glDrawElements(GL_TRIANGLES, (GLsizei)ni, GL_UNSIGNED_INT, (const void*)0);
		const double d = now_ms();
		frame = glGetFramePtr();
		const double e = now_ms();
		phase[0] += b - a; phase[1] += c - b; phase[2] += d - c; phase[3] += e - d;
	}
	const double ms = (now_ms() - t0) / (double)(frames > 0 ? frames : 1);
	fprintf(stderr, "devices %d: vertex upload %.3f ms, index upload %.3f ms, clear + draw calls %.3f ms, glGetFramePtr %.3f ms\n", devices,
	        phase[0] / (frames > 0 ? frames : 1), phase[1] / (frames > 0 ? frames : 1), phase[2] / (frames > 0 ? frames : 1), phase[3] / (frames > 0 ? frames : 1));
	const char* err = swglGetLastError();
	if (err[0]) { fprintf(stderr, "error: %s\n", err); return 3; }

	const float* depth = swglGetDepthPtr();
	unsigned long long covered = 0;
	for (size_t p = 0; p < (size_t)width * height; p++) { uint32_t bits; memcpy(&bits, depth + p, 4); covered += bits != 0u; }
	printf("%d %016llx %llu %.4f\n", devices, (unsigned long long)swglHashWords(frame, (uint64_t)width * height), covered, ms);
	swglHostFree(verts);
	swglHostFree(idx);
	return 0;
}
