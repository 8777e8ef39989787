import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import swgl_b200 as sw
from swgl_b200 import scenes as S
from util import gpu_render
api = sw.load()
for sc, opts in [(S.random_triangles(300, 160, 120, seed=3, near_cross=True, alpha=None, centre_range=1.3), {"raster_path": 3}),
                 (S.grid_mesh(40, 333, 211, alpha=0.5, use_matrix=True), {}),
                 (S.random_triangles(200, 160, 120, seed=4, extent=0.9, alpha=0.5), {"raster_path": 3}),
                 (S.grid_mesh(30, 200, 150, textured=True), {"raster_path": 3}),
                 (S.random_triangles(300, 160, 120, seed=5, near_cross=True), {"raster_path": 2}),
                 (S.random_triangles(300, 160, 120, seed=6, near_cross=True), {"raster_path": 1})]:
    col, dep, st, err = gpu_render(api, sc, indexed=sc.indices is not None, options=opts)
    print(sc.name, st["tested"], err)
