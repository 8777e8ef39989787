/*
 * include/swgl_b200.h -- extension entry points of libswgl_b200.so (none exist in the
 * reference; they cover what the reference exposes only as file-scope globals, plus device
 * control).  The reference-compatible API is in swgl.h.
 */
#ifndef SWGL_B200_H
#define SWGL_B200_H

#include "swgl.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
	uint64_t draws;            /* draw calls executed since glInit */
	uint64_t triangles_in;     /* last draw: input triangles */
	uint64_t prims_out;        /* last draw: triangles after the near clip */
	uint64_t tested;           /* last draw: fragments evaluated (reference: Barycentric calls) */
	uint64_t shaded;           /* last draw: fragments that passed the depth test (FS runs) */
	uint64_t tile_pairs;       /* last draw: (tile, primitive) bin entries */
	uint64_t bands;            /* last draw: band-entry records */
} swglStats;

/* The reference keeps depth in GlobalFramebuffer->DepthAttachment (swgl.c:3165) with no
 * accessor; this returns the pinned host mirror, refreshed like glGetFramePtr. */
float* swglGetDepthPtr(void);

/* Block until all queued device work is done (the reference is synchronous). */
void swglFinish(void);

/* "" when no error is pending.  CUDA failures, unsupported-viewport and shader-compile
 * diagnostics land here; the reference-compatible entry points keep their silent returns. */
const char* swglGetLastError(void);

/* Shaders outside the built-in shapes (pass-through / matrix vertex shader, varying / texture fragment
 * shader) are turned into kernels at run time, at the first draw that uses them (about a second, once
 * per program and vertex layout per process).  This compiles the kernels of the program in use with the
 * bound vertex array now.  Without a device the code is generated and compiled but not loaded.
 * Returns 0 on success (or nothing to compile), -1 on failure (swglGetLastError). */
int swglPrecompileProgram(void);

/* Fragment / primitive counters of the last draw (synchronises). */
void swglGetStats(swglStats* out);

/* Select the CUDA device used by the NEXT glInit (default: $LOCAL_RANK, else 0). */
void swglSetDevice(int ordinal);

/* Number of CUDA devices the NEXT glInit drives from the calling thread (default 1): ordinals d .. d+n-1 with d
 * the swglSetDevice ordinal; with the default start (0) and fewer devices than are visible, every (visible / n)-th
 * device in PCI bus order instead (neighbouring GPUs of a board share their path to host memory).  The frame is sharded sort-first by tile-row bands, buffers and textures
 * are replicated (swglBufferRespecify splits the transfer over the devices' PCIe links and completes it over
 * NVLink), and glGetFramePtr returns one pinned frame every device has written its bands into.  The
 * reference-compatible entry points need no other change.  Not available in this mode: swglFrameSubmit,
 * swglSetStripe and the peer / shared-mirror targets (the group does its own assembly); swglGetColorDevicePtr
 * returns device d's attachment, which only holds that device's bands. */
void swglSetDeviceCount(int count);

/* cudaStream_t (as void*) on which glClear / glDraw* queue their kernels. */
void* swglGetStream(void);

/* Device addresses of the colour (uint32 [H][W]) and depth (float [H][W]) attachments. */
uint64_t swglGetColorDevicePtr(void);
uint64_t swglGetDepthDevicePtr(void);

/* Fill the WHOLE framebuffer (glInit leaves it uninitialised, swgl.c:3720-3723; glClear
 * only touches the viewport rectangle). */
void swglFillFramebuffer(uint32_t color_word, float depth);

/* Replace the contents of the buffer bound to `target` (GL_ARRAY_BUFFER or
 * GL_ELEMENT_ARRAY_BUFFER).  The reference's glBufferData ignores re-specification
 * (swgl.c:3140); streaming geometry therefore needs this extension.  Size may differ. */
void swglBufferRespecify(GLenum target, GLsizei size, const void* data);

/* Sort-first sharding: this process rasterises only tile-row bands b with b % n_ranks == rank
 * (band = band_tile_rows rows of 32-pixel tiles).  Geometry is replicated. */
void swglSetStripe(GLuint rank, GLuint n_ranks, GLuint band_tile_rows);

/* Peer colour target (device address valid in this process, e.g. from cudaIpcOpenMemHandle):
 * raster write-back also stores each finished tile row into it.  0 disables. */
void swglSetPeerColorTarget(uint64_t device_ptr);

/* Shared frame mirror: the other way to assemble a sort-first frame, for applications that want it in HOST
 * memory (glGetFramePtr).  `host_ptr` is W*H*4 bytes (page aligned, `bytes` a multiple of the page size) that
 * every rank has mapped -- POSIX shared memory, for instance; each rank calls this once.  From then on the raster
 * kernels of a rank store its finished tiles straight into that memory over the rank's own PCIe link, so N links
 * carry the frame in parallel instead of rank 0's alone; swglFinish / glGetFramePtr bring the calling rank's
 * bands up to date (by copies, when a frame did not start from a whole-framebuffer clear), and after the
 * application's barrier glGetFramePtr on any rank returns the assembled frame: `host_ptr` itself.
 * NULL switches it off.  Returns 0 on success. */
int swglSetSharedFrameMirror(void* host_ptr, uint64_t bytes);

/* CUDA IPC plumbing for the peer colour target (one process per GPU): rank 0 exports its colour
 * attachment as a 64-byte handle, the other ranks open it and pass the address to
 * swglSetPeerColorTarget.  Return 0 / an address on success. */
int      swglIpcExportColor(void* handle64);
uint64_t swglIpcOpen(const void* handle64);
void     swglIpcClose(uint64_t device_ptr);

/* The step after the path (glGetFramePtr consumers, swgl.c:3738): frame pipelining.  swglFrameSubmit stamps
 * the frame drawn so far and returns a ticket (> 0) at once; swglFrameWait blocks until that frame lies in
 * one of two pinned mirrors and returns it (row 0 = top, the reference's R<<24|G<<16|B<<8|A words).  The
 * pointer stays valid until the next-but-one swglFrameSubmit.  With swglBufferRespecify on a second set of
 * buffers, frame N+1's upload crosses PCIe while frame N is rasterised and written to its mirror. */
uint64_t swglFrameSubmit(void);
const uint32_t* swglFrameWait(uint64_t ticket);
/* The colour attachment as bytes R, G, B, A (swizzled on the device) into W*H*4 bytes of host memory. */
int swglReadPixelsRGBA8(void* dst);
/* The current frame as a binary PPM (P6); 0 on success. */
int swglWritePPM(const char* path);
/* FNV-1a over 32-bit words (h = (h ^ word) * 1099511628211 from 1469598103934665603): the image hash of the
 * committed known-answer frames (tests/golden), for applications and bench.py to check a frame they read back. */
uint64_t swglHashWords(const void* words, uint64_t n_words);

/* Geometry that reaches the device in pieces (sort-first ranks: every rank uploads 1/N of the arrays over its own
 * PCIe link and an all-gather over NVLink, queued on swglGetStream(), replicates them).
 * swglBufferSubData copies `size` bytes to byte `offset` of the bound buffer (already specified with that size or
 * more; the copy completes before the call returns and waits only for draws that read the buffer);
 * swglGetBufferDevicePtr is the device address of the bound buffer's storage;
 * swglBufferDeviceWritten tells the library that work queued on its stream has rewritten the bound buffer (element
 * data: the largest index is recomputed behind that work). */
void     swglBufferSubData(GLenum target, uint64_t offset, GLsizei size, const void* data);
uint64_t swglGetBufferDevicePtr(GLenum target);
void     swglBufferDeviceWritten(GLenum target);

/* Page-locked staging memory for the application's vertex / index arrays (the source of glBufferData and
 * swglBufferRespecify).  write_combined != 0 asks for write-combined pages: the CPU writes them once,
 * front to back, and never reads them, and host-to-device copies do not have to snoop the CPU caches --
 * on the boxes measured, freshly written cacheable pinned pages upload at a fraction of the PCIe rate for
 * the first tens of copies (bench.py's end-to-end step varied 1.27 .. 2.2 ms), write-combined ones at
 * the full rate from the first.  NULL on failure.  Works without glInit. */
void* swglHostAlloc(uint64_t bytes, int write_combined);
void  swglHostFree(void* p);

/* Tuning / test hooks: "raster_path" (0 default = 3, 1 pixel-owner CTA, 2 fragment-parallel CTA, 3 warp per 32x8 tile),
 * "host_mirror" (1 adaptive: when glGetFramePtr follows every draw or two, the raster kernels also store finished
 * tiles into the pinned frame mirror and glGetFramePtr only waits; 0 always copy; 2 whenever the mirror is in sync),
 * "fuse_clear" (0/1), "count_fragments" (0/1), "stage_timing" (0/1: per-kernel CUDA-event timing, synchronous);
 * "bin_cap" (per-tile list capacity, test hook), "bin_limit_bytes", "lean_prims" (1 default: short narrow unclipped
 * primitives have no record, the rasteriser gathers them through the element buffer; 0 writes a record for every one),
 * "selftest_division" (value = number of operand pairs: runs the device self-test of the shared-reciprocal division
 * against `/`, mismatches in "selftest_division_mismatches");
 * "jit" (1 default: shaders outside the built-in shapes are compiled to kernels at run time with NVRTC; 0: the
 * on-device interpreter, kept as the cross-check);
 * "mip_lod" (0 default: glGenerateMipmap builds the chain but the base level is sampled, which is what the compiled
 * reference does -- its level of detail goes through an rsqrt() with undefined behaviour, swgl.c:3240-3246; 1: the
 * chain is sampled with the per-triangle level the same code gives with a 32-bit pun, bit-identical to the reference
 * built that way, oracle/ref_shim.c -- including what the reference does around it: levels outlive the image they were
 * built from, a second glGenerateMipmap appends its levels behind the first, GL_POINTS sample with the level the last
 * triangle left);
 * "setup_big" (1 default: draws of big triangles -- more than 64 pixels per triangle on average -- are set up by a warp per
 * triangle; 0: a thread per triangle plus a second kernel for the tall ones), "setup_pipelined" (0 default; 1: the
 * software-pipelined form of the set-up kernel, measured slower);
 * "tile_rows" (0 default: 8 rows per tile, 4 when the context owns less than one resident wave of 8-row tiles; 8 / 4 /
 * 2 pin it); "overflow_pool" (1 default: tile lists longer than bin_cap continue in a shared overflow pool; 0: bin_cap
 * grows for every tile instead);
 * read-only: "wt_draws", "mirror_synced", "kernel_launches", "stage_ns_0".."stage_ns_2" (vertex, setup+bin, raster),
 * "stage_draws", "tile_size", "last_tile_rows", "device", "device_count", "last_vs_kind" / "last_fs_kind" (shape of the last
 * draw as launched: 0 interpreter, 1 / 2 built-in shapes, 3 run-time compiled), "jit_compiles", "jit_cache_hits",
 * "jit_compile_us_total", "draws_folded" (draws whose viewport leaves the framebuffer rows: the reference folds the raster
 * rows outside onto row Height-1, swgl.c:3386; drawn identically here, through a slower two-part path), "draws_refused"
 * (such draws in configurations that cannot fold -- sort-first ranks of separate processes, peer / shared-mirror targets, a virtual
 * framebuffer beyond 8184 rows: nothing is drawn and swglGetLastError says so), "overflow_pool_entries", "pairs_bytes",
 * "bin_cap". */
void swglSetOption(const char* name, int64_t value);
int64_t swglGetOption(const char* name);

/* Text dump of a compiled shader's IR and of the shape it was recognised as. Returns the
 * number of bytes that the full dump needs (excluding the terminator). */
size_t swglDebugShaderIR(GLuint shader, char* buf, size_t buf_len);
/* 1 if glCompileShader produced executable IR, 0 if the source is outside the subset. */
int swglGetShaderCompiled(GLuint shader);

#ifdef __cplusplus
}
#endif

#endif /* SWGL_B200_H */
