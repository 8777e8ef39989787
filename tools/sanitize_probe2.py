"""compute-sanitizer target for the round-2 kernels: tile heights 8 / 4 / 2, the overflow pool of a deep tile,
k_mipmap_box + LOD sampling, a run-time compiled program, a folded viewport (k_fold_row), and a device group of
emulated members (k_group_gather, owned-band clear).  Small scenes: the sanitizer slows kernels ~50x."""
import copy
import ctypes as C
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

import swgl_b200 as sw
from swgl_b200 import gl as G, scenes as S
from util import gpu_render

api = sw.load()
for rows in (8, 4, 2):
    sc = S.random_triangles(200, 160, 120, seed=3 + rows, near_cross=True, alpha=None, centre_range=1.3)
    print(sc.name, rows, gpu_render(api, sc, indexed=False, options={"raster_path": 3, "tile_rows": rows})[2]["tested"])
# deep tile -> overflow pool
mesh = S.grid_mesh(12, 160, 120, alpha=0.5)
rng = np.random.default_rng(3)
deep = 2000
spot = np.empty((deep * 3, 8), np.float32)
spot[:, 0:2] = np.array([0.3, -0.2], np.float32) + rng.uniform(-0.03, 0.03, (deep * 3, 2)).astype(np.float32)
spot[:, 2] = rng.uniform(0.1, 0.9, deep * 3); spot[:, 3] = 1.0
spot[:, 4:7] = rng.uniform(0, 1, (deep * 3, 3)); spot[:, 7] = 0.5
sc = copy.copy(mesh); sc.vertices = np.concatenate([mesh.deindexed(), spot]).astype(np.float32); sc.indices = None
print("deep", gpu_render(api, sc, indexed=False)[2]["tested"], api.swglGetOption(b"bin_cap"))
# generic (run-time compiled) shader + mip chain with the defined LOD
sc = S.grid_mesh(16, 160, 120, textured=True)
sc.fs = "in vec2 vUV;\nuniform sampler2D uTex;\nout vec4 FragColor;\nvoid main()\n{\nvec4 t = texture(uTex,vUV).zyxw;\nFragColor = t * vec4(0.9, 0.8, 0.7, 1.0);\n}\n"
sc.texture = S.lcg_texture(64)
api.glInit(sc.width, sc.height)
api.swglSetOption(b"mip_lod", 1)
st = G.setup_scene(api, sc, indexed=True, init=False)
api.glGenerateMipmap(G.GL_TEXTURE_2D)
api.glClear(3)
api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
api.glGetFramePtr()
print("jit+mip", api.swglGetOption(b"last_fs_kind"), api.swglGetLastError())
# folded viewport
sc = S.random_triangles(150, 160, 120, seed=9, near_cross=True, alpha=None)
api.glInit(sc.width, sc.height)
st = G.setup_scene(api, sc, indexed=False, init=False)
api.glViewport(-5, -20, 170, 170)
api.glClear(3)
api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])
api.glGetFramePtr()
print("folded", api.swglGetOption(b"draws_folded"), api.swglGetLastError())
# device group of emulated members, unfused clear, respecify (sharded upload + gather)
os.environ["SWGL_B200_GROUP_EMULATE"] = "1"
sc = S.grid_mesh(24, 200, 160)
api.swglSetDeviceCount(3)
api.glInit(sc.width, sc.height)
api.swglSetOption(b"fuse_clear", 0)
st = G.setup_scene(api, sc, indexed=True, init=False)
v = np.ascontiguousarray(sc.vertices); i = np.ascontiguousarray(sc.indices)
for _ in range(2):
    api.swglBufferRespecify(G.GL_ARRAY_BUFFER, v.nbytes, v.ctypes.data_as(C.c_void_p))
    api.swglBufferRespecify(G.GL_ELEMENT_ARRAY_BUFFER, i.nbytes, i.ctypes.data_as(C.c_void_p))
    api.glClear(3)
    api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
    api.glGetFramePtr()
api.swglGetDepthPtr()
print("group", api.swglGetOption(b"device_count"), api.swglGetLastError())
api.swglSetDeviceCount(1)
api.glInit(16, 16)
