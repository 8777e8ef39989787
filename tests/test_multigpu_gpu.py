"""GPU (needs >= 2 devices, skipped otherwise): two processes, NCCL control plane, each rank
rasterises its tile-row bands; the three assembly paths (peer stores over CUDA IPC, NCCL send/recv,
write-through into a shared host segment) must reproduce the single-GPU frame bit for bit."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, q):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import swgl_b200 as sw
    from swgl_b200 import gl as G, multigpu, scenes as S

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    api = sw.load()
    scene = S.config(2)
    api.swglSetDevice(rank)
    api.glInit(scene.width, scene.height)
    st = G.setup_scene(api, scene, indexed=True, init=False)
    band = 2
    api.swglFillFramebuffer(0x01020304, C.c_float(0.0))
    api.swglSetStripe(rank, world, band)
    peer = multigpu.PeerColorTarget(api, dist, rank, world) if mode == "peer" else None
    if mode == "host":
        peer = multigpu.SharedFrameMirror(api, dist, rank, world, scene.width, scene.height)
        # the geometry again, this time 1/N per rank over PCIe + all-gather over NVLink: scrub the buffers first
        hv, hi = multigpu.HostArray(api, scene.vertices), multigpu.HostArray(api, scene.indices)
        junk = multigpu.HostArray(api, np.zeros(scene.vertices.size, np.float32))
        api.swglBufferRespecify(G.GL_ARRAY_BUFFER, junk.nbytes, C.c_void_p(junk.ptr))
        up = multigpu.ShardedUpload(api, dist, rank, world, torch.device("cuda", rank))
        assert up.divisible(hv.nbytes, world) and up.divisible(hi.nbytes, world)
        up.upload(G.GL_ARRAY_BUFFER, hv)
        up.upload(G.GL_ELEMENT_ARRAY_BUFFER, hi)
    api.glClear(3)
    api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
    api.swglFinish()
    dist.barrier()
    if mode == "nccl":
        multigpu.gather_color(api, dist, rank, world, scene.height, scene.width, band, torch.device("cuda", rank))
    if rank == 0:
        multi = G.frame_color(api, scene.width, scene.height).copy()
        if mode == "host":
            assert np.array_equal(multi, peer.frame(scene.height, scene.width))
            peer.close()
            peer = None
        api.swglSetStripe(0, 1, 1)
        api.glClear(3)
        api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
        single = G.frame_color(api, scene.width, scene.height)
        q.put((bool(np.array_equal(multi, single)), api.swglGetLastError().decode()))
    dist.barrier()
    if peer is not None:
        peer.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "nccl", "host"])
def test_two_ranks_assemble_the_single_gpu_frame(mode):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    equal, err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
    assert err == ""
    assert equal
