/*
 * swgl_jit.h -- run-time compiled shader kernels (swgl_jit.cpp), internal to libswgl_b200.so.
 */
#ifndef SWGL_JIT_H
#define SWGL_JIT_H

#include <stddef.h>
#include <stdint.h>

#include "swgl_dev.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
	void* vertex;      /* cudaKernel_t of k_vertex<SWVS_JIT>, NULL when not requested */
	void* raster;      /* cudaKernel_t of k_raster_warp<SWFS_JIT> */
} swgljit_kernels;

typedef struct
{
	uint64_t compiles, cache_hits;
	double   compile_ms_total;
	uint64_t cubin_bytes;          /* of the last program compiled */
} swgljit_stats;

/* Kernels for the draw's shader pair and interface, compiled on first use and cached for the life of
 * the process.  device < 0: generate and compile only (no device needed), nothing is loaded.
 * Returns 0, or -1 with a message in `err`. */
int swgljit_get(int device, const swgldev_draw* d, int want_vertex, int want_raster, swgljit_kernels* out, char* err, size_t errlen);
const char* swgljit_last_source(void);
void swgljit_get_stats(swgljit_stats* s);

#ifdef __cplusplus
}
#endif

#endif
