/* tests/frontend_fuzz/harness.c -- test infrastructure: the GLSL-subset front end (swgl_b200/csrc/swgl_glsl.c) alone,
 * built with -fsanitize=address,undefined by tests/test_frontend_sanitizers.py.  Reads records (4-byte little-endian
 * length + bytes) from stdin, compiles each, dumps the IR twice (short buffer, exact buffer), frees everything. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "swgl_glsl.h"

int main(void)
{
	unsigned n = 0, ok = 0;
	for (;;)
	{
		unsigned len;
		if (fread(&len, 4, 1, stdin) != 1) break;
		char* s = (char*)malloc((size_t)len + 1);
		if (!s || fread(s, 1, len, stdin) != len) { free(s); break; }
		s[len] = 0;
		swgl_shader* sh = swgl_glsl_compile(s);
		if (sh)
		{
			char buf[256];
			swgl_glsl_dump(sh, buf, sizeof(buf));
			const size_t need = swgl_glsl_dump(sh, NULL, 0);
			char* big = (char*)malloc(need + 1);
			if (big) swgl_glsl_dump(sh, big, need + 1);
			free(big);
			swgl_glsl_find_var(sh, "tint");
			swgl_glsl_find_var(sh, "gl_Position");
			ok++;
			swgl_glsl_free(sh);
		}
		free(s);
		n++;
	}
	printf("compiled %u records, %u returned a shader\n", n, ok);
	return 0;
}
