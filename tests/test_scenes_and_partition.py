"""CPU: scene generators, algorithmic-byte accounting, and the sort-first partition + gather
logic over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest

from swgl_b200 import multigpu, scenes as S


def test_lcg_stream_matches_scalar_recurrence():
    s, out = 12345, []
    for _ in range(1000):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        out.append(np.float32(s >> 8) / np.float32(16777216.0))
    assert np.array_equal(S.lcg_stream(12345, 1000), np.array(out, np.float32))


def test_algorithmic_bytes_match_baseline_md():
    """V*stride + I*4 + texture + W*H*8."""
    assert S.config(1).algorithmic_bytes() == 2_553_600
    assert S.config(2).algorithmic_bytes() == 19_413_024
    assert S.config(3).algorithmic_bytes() == 23_202_328
    c4 = S.config(4)
    assert c4.n_triangles == 1_002_528 and len(c4.vertices) == 502_681
    assert c4.algorithmic_bytes() == 94_471_328


def test_partition_covers_every_row_once():
    for height in (480, 1080, 2160, 4320):
        for n in (1, 2, 4, 8):
            for band in (1, 2, 4, 17):
                seen = np.zeros(height, np.int32)
                for r in range(n):
                    for r0, r1 in multigpu.rows_of_rank(height, r, n, band):
                        seen[r0:r1] += 1
                assert (seen == 1).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, height, width, band, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    full = rng.integers(0, 2**31 - 1, size=(height, width), dtype=np.int32)   # the "frame"
    mine = np.zeros_like(full)
    for r0, r1 in multigpu.rows_of_rank(height, rank, world, band):
        mine[r0:r1] = full[r0:r1]       # each rank only has its own bands
    img = torch.from_numpy(mine)
    multigpu.gather_rows(img, dist, rank, world, height, band)
    if rank == 0:
        q.put(bool(np.array_equal(img.numpy(), full)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("band", [1, 3])
def test_gather_rows_world2_gloo(band):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 200, 64, band, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok


class _FakeMirrorApi:
    """Stands in for the library on a box without a GPU: records the registered segment."""

    def __init__(self, fail=False):
        self.addr, self.fail = None, fail

    def swglSetSharedFrameMirror(self, ptr, nbytes):
        if ptr is not None and self.fail:
            return -1
        self.addr = ptr
        return 0

    def swglGetLastError(self):
        return b"refused (test)"


def _mirror_worker(rank, world, port, height, width, band, fail_rank, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    api = _FakeMirrorApi(fail=(rank == fail_rank))
    try:
        m = multigpu.SharedFrameMirror(api, dist, rank, world, width, height)
    except RuntimeError:
        if rank == 0:
            q.put("refused")        # one rank could not register: every rank backs out together
        dist.barrier()
        dist.destroy_process_group()
        return
    frame = m.frame(height, width)
    for r0, r1 in multigpu.rows_of_rank(height, rank, world, band):
        frame[r0:r1] = rank + 1      # what the rank's raster kernels do through the mapping
    dist.barrier()
    if rank == 0:
        want = np.zeros((height, width), np.uint32)
        for r in range(world):
            for r0, r1 in multigpu.rows_of_rank(height, r, world, band):
                want[r0:r1] = r + 1
        q.put("ok" if np.array_equal(frame, want) else "mismatch")
    dist.barrier()
    del frame
    m.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank,expect", [(-1, "ok"), (1, "refused")])
def test_shared_frame_mirror_world2_gloo(fail_rank, expect):
    """The host-side half of swglSetSharedFrameMirror: one POSIX shared-memory segment, mapped by both
    ranks, each writing its bands; a rank that cannot register it makes both back out."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mirror_worker, args=(r, 2, port, 200, 64, 1, fail_rank, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert got == expect


def test_sharded_upload_chunking():
    assert multigpu.ShardedUpload.divisible(16_085_792, 8) and multigpu.ShardedUpload.divisible(12_030_336, 8)
    assert not multigpu.ShardedUpload.divisible(16_085_792 + 4, 8)
