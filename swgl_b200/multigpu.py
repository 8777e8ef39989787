"""Sort-first multi-GPU plumbing: one process per GPU, torch.distributed for the control plane.

The frame is split by tile-row bands dealt round-robin to the ranks (``swglSetStripe``);
geometry is replicated, so there is no exchange until the image is assembled on rank 0.
Three ways to assemble it:

* ``PeerColorTarget`` -- rank 0 exports its colour attachment through CUDA IPC, every other rank
  maps it and the raster kernel's 128-bit write-back stores each finished strip straight into
  rank 0's framebuffer over NVLink (the collective is fused into the kernel epilogue; only a
  barrier remains);
* ``SharedFrameMirror`` -- for frames that are wanted in host memory: every rank maps one shared
  host segment and its raster kernels write finished tiles there over the rank's own PCIe link;
* ``gather_color`` -- plain NCCL: every rank contributes its band rows, rank 0 receives them
  (the baseline the fused variant is measured against).

``owner_of_tile_row`` is the pure partition function; CPU tests exercise it with gloo.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

TILE = 32


def owner_of_tile_row(tile_row: int, n_ranks: int, band_tile_rows: int) -> int:
    """Rank that rasterises ``tile_row`` (mirrors owns_tile_row() in swgl_dev.cu)."""
    return (tile_row // max(1, band_tile_rows)) % max(1, n_ranks)


def rows_of_rank(height: int, rank: int, n_ranks: int, band_tile_rows: int):
    """Framebuffer row ranges [(r0, r1), ...] owned by ``rank`` (storage rows, row 0 = top)."""
    tiles_y = (height + TILE - 1) // TILE
    out = []
    for tr in range(tiles_y):
        if owner_of_tile_row(tr, n_ranks, band_tile_rows) != rank:
            continue
        r0, r1 = tr * TILE, min(height, (tr + 1) * TILE)
        if out and out[-1][1] == r0:
            out[-1] = (out[-1][0], r1)
        else:
            out.append((r0, r1))
    return out


def assemble(stripes, height: int, width: int, n_ranks: int, band_tile_rows: int) -> np.ndarray:
    """Host-side assembly of per-rank full-size images into one (what the collective computes)."""
    out = np.zeros((height, width), dtype=stripes[0].dtype)
    for r in range(n_ranks):
        for r0, r1 in rows_of_rank(height, r, n_ranks, band_tile_rows):
            out[r0:r1] = stripes[r][r0:r1]
    return out


class _DevArray:
    """Minimal __cuda_array_interface__ wrapper so torch can alias library-owned device memory."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def color_tensor(api, height: int, width: int, device):
    """torch view (int32 [H, W]) of the library's colour attachment -- no copy."""
    import torch

    return torch.as_tensor(_DevArray(api.swglGetColorDevicePtr(), (height, width), "<i4"), device=device)


class PeerColorTarget:
    """Map rank 0's colour attachment into every other rank and make it their store target."""

    def __init__(self, api, dist, rank: int, world: int):
        self.api, self.rank, self.ptr = api, rank, 0
        handle = (C.c_ubyte * 64)()
        if rank == 0:
            rc = api.swglIpcExportColor(handle)
            if rc != 0:
                raise RuntimeError("cudaIpcGetMemHandle failed: " + api.swglGetLastError().decode())
        box = [bytes(handle)]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            buf = (C.c_ubyte * 64).from_buffer_copy(box[0])
            self.ptr = api.swglIpcOpen(buf)
            if not self.ptr:
                raise RuntimeError("cudaIpcOpenMemHandle failed: " + api.swglGetLastError().decode())
            api.swglSetPeerColorTarget(self.ptr)
        dist.barrier()

    def close(self):
        if self.ptr:
            self.api.swglSetPeerColorTarget(0)
            self.api.swglIpcClose(self.ptr)
            self.ptr = 0


class HostArray:
    """A numpy array copied into page-locked, write-combined host memory of the library (swglHostAlloc)."""

    def __init__(self, api, a):
        a = np.ascontiguousarray(a)
        self.api = api
        self.nbytes = int(a.nbytes)
        self.ptr = api.swglHostAlloc(self.nbytes, 1)
        if not self.ptr:
            raise RuntimeError("swglHostAlloc failed")
        C.memmove(self.ptr, a.ctypes.data, self.nbytes)

    def free(self):
        if self.ptr:
            self.api.swglHostFree(self.ptr)
            self.ptr = None


class ShardedUpload:
    """Replicated geometry without replicated PCIe traffic: every rank copies 1/N of an array into the
    bound buffer over its own PCIe link (``swglBufferSubData``) and an in-place NCCL all-gather over
    NVLink, queued on the library's stream, hands every rank the rest."""

    def __init__(self, api, dist, rank: int, world: int, device):
        import torch

        self.api, self.dist, self.rank, self.world, self.device = api, dist, rank, world, device
        self.stream = torch.cuda.ExternalStream(api.swglGetStream(), device=device)
        self._views = {}

    @staticmethod
    def divisible(nbytes: int, world: int) -> bool:
        return nbytes % (4 * world) == 0       # equal chunks of whole 32-bit words

    def _view(self, ptr: int, nbytes: int):
        import torch

        key = (ptr, nbytes)
        if key not in self._views:
            self._views[key] = torch.as_tensor(_DevArray(ptr, (nbytes // 4,), "<i4"), device=self.device)
        return self._views[key]

    def upload(self, target: int, host: HostArray):
        import torch

        n, w, r = host.nbytes, self.world, self.rank
        if not self.divisible(n, w):
            raise ValueError("array size is not a multiple of 4 bytes x ranks")
        chunk = n // w
        self.api.swglBufferSubData(target, r * chunk, chunk, C.c_void_p(host.ptr + r * chunk))
        full = self._view(self.api.swglGetBufferDevicePtr(target), n)
        mine = full[r * chunk // 4:(r + 1) * chunk // 4]
        with torch.cuda.stream(self.stream):
            self.dist.all_gather_into_tensor(full, mine)
        self.api.swglBufferDeviceWritten(target)


class SharedFrameMirror:
    """Assemble the frame in HOST memory: one POSIX shared-memory segment mapped by every rank
    (``swglSetSharedFrameMirror``).  Each rank's raster kernels store its finished tiles there over the
    rank's own PCIe link, so the frame reaches the host over N links in parallel and rank 0's
    ``glGetFramePtr`` has nothing left to copy.  ``dist`` may be None (single process)."""

    PAGE = 4096

    def __init__(self, api, dist, rank: int, world: int, width: int, height: int):
        import mmap
        import os

        self.api, self.rank, self.active = api, rank, False
        self.nbytes = width * height * 4
        size = (self.nbytes + self.PAGE - 1) // self.PAGE * self.PAGE
        box = [None]
        if rank == 0:
            path = f"/dev/shm/swgl_b200_frame_{os.getpid()}"
            try:
                fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o600)
                try:
                    os.posix_fallocate(fd, 0, size)     # a shared-memory file system that is too small says so here, not by SIGBUS later
                finally:
                    os.close(fd)
                box[0] = path
            except OSError:
                try:
                    os.unlink(path)
                except OSError:
                    pass
        if dist is not None:
            dist.broadcast_object_list(box, src=0)
        self.path = box[0]
        self.map, self.addr, ok = None, 0, False
        try:
            if self.path is None:
                raise OSError("no shared segment")
            fd = os.open(self.path, os.O_RDWR)
            try:
                self.map = mmap.mmap(fd, size)
            finally:
                os.close(fd)
            self.addr = C.addressof(C.c_char.from_buffer(self.map))
            ok = api.swglSetSharedFrameMirror(C.c_void_p(self.addr), size) == 0
        except (OSError, ValueError, TypeError):
            ok = False          # every rank still reaches the vote below
        if dist is not None:
            votes = [None] * world
            dist.all_gather_object(votes, bool(ok))
            all_ok = all(votes)
        else:
            all_ok = ok
        if not all_ok:          # some rank could not register the segment: nobody uses it
            if ok:
                api.swglSetSharedFrameMirror(None, 0)
            self._unmap()
            raise RuntimeError("swglSetSharedFrameMirror failed: " + api.swglGetLastError().decode())
        self.active = True
        if dist is not None:
            dist.barrier()

    def frame(self, height: int, width: int) -> np.ndarray:
        """The assembled frame ([H, W] uint32 view of the shared segment)."""
        return np.frombuffer(self.map, dtype=np.uint32, count=height * width).reshape(height, width)

    def _unmap(self):
        import os

        try:
            if self.map is not None:
                self.map.close()
        except (BufferError, ValueError):
            pass                # a numpy view is still alive: the mapping goes with the process
        if self.rank == 0 and self.path is not None:
            try:
                os.unlink(self.path)
            except OSError:
                pass

    def close(self):
        if self.active:
            self.api.swglSetSharedFrameMirror(None, 0)
            self.active = False
            self._unmap()


def gather_rows(img, dist, rank: int, world: int, height: int, band_tile_rows: int):
    """Assemble the band rows of every rank into ``img`` on rank 0 (grouped send/recv).

    ``img`` is a [H, W] tensor on any device torch.distributed can move (CUDA with NCCL, CPU
    with gloo); rank 0's own rows are already in place."""
    ops = []
    if rank == 0:
        for r in range(1, world):
            for r0, r1 in rows_of_rank(height, r, world, band_tile_rows):
                ops.append(dist.P2POp(dist.irecv, img[r0:r1], r))
    else:
        for r0, r1 in rows_of_rank(height, rank, world, band_tile_rows):
            ops.append(dist.P2POp(dist.isend, img[r0:r1], 0))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return img


def gather_color(api, dist, rank: int, world: int, height: int, width: int, band_tile_rows: int, device):
    """NCCL assembly on rank 0, in place in the library's colour attachment."""
    import torch

    img = color_tensor(api, height, width, device)
    api.swglFinish()
    gather_rows(img, dist, rank, world, height, band_tile_rows)
    torch.cuda.synchronize()
    return img
