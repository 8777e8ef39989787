"""Helpers shared by the GPU parity tests: drive libswgl_b200.so through its C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

import swgl_b200
from swgl_b200 import gl as G


def gpu_render(api, scene, *, indexed=True, clear=True, fill=(0, 0.0), first=0, count=None,
               options=None, stripe=None, draws=None, mipmaps=False, points=None):
    """Render ``scene`` with the CUDA library; returns (color, depth, stats dict, error str).
    mipmaps: glGenerateMipmap on the scene's texture (sampled when options has mip_lod = 1);
    points = (first, count): glDrawArrays(GL_POINTS, ...) over the same vertex stream after the triangles."""
    api.glInit(scene.width, scene.height)
    err = api.swglGetLastError()
    assert not err, f"glInit failed: {err!r}"
    if options:
        for k, v in options.items():
            api.swglSetOption(k.encode(), int(v))
    if stripe:
        api.swglSetStripe(*stripe)
    st = G.setup_scene(api, scene, indexed=indexed, init=False)
    if mipmaps:
        api.glGenerateMipmap(G.GL_TEXTURE_2D)
    api.swglFillFramebuffer(fill[0], C.c_float(fill[1]))
    if clear:
        api.glClear(3)
    n = st["n_draw"] if count is None else count
    for (f, k) in (draws or [(first, n)]):
        if st["indexed"]:
            api.glDrawElements(G.GL_TRIANGLES, k, G.GL_UNSIGNED_INT, C.c_void_p(4 * f))
        else:
            api.glDrawArrays(G.GL_TRIANGLES, f, k)
    if points is not None:
        api.glDrawArrays(G.GL_POINTS, points[0], points[1])
    H, W = scene.height, scene.width
    col = G.frame_color(api, W, H)
    dep = np.ctypeslib.as_array(api.swglGetDepthPtr(), shape=(H, W)).copy()
    s = swgl_b200.swglStats()
    api.swglGetStats(C.byref(s))
    return col, dep, s.as_dict(), api.swglGetLastError().decode()


def assert_bit_exact(cmp: dict, what: str = ""):
    assert cmp["coverage_mismatch"] == 0, f"{what}: coverage differs: {cmp}"
    assert cmp["depth_mismatch"] == 0, f"{what}: depth bits differ: {cmp}"
    # north_star tolerance for colour is 1/255 per channel; the kernels are written to be exact
    assert cmp["max_channel_delta"] <= 1, f"{what}: colour off by more than 1/255: {cmp}"
    assert cmp["color_mismatch"] == 0, f"{what}: colour words differ (within 1/255 but not exact): {cmp}"
