"""C3 geometry (1080p, 100k textured triangles) with a fragment shader outside the built-in shapes:
run-time compiled kernel vs the on-device interpreter vs the built-in texture shape (development / profiles)."""
import json
import sys

sys.path.insert(0, ".")
import swgl_b200 as sw
from swgl_b200 import scenes as S
from tools.perf_probe import probe, api

GENERIC_FS = ("in vec2 vUV;\nuniform sampler2D uTex;\nout vec4 FragColor;\nvoid main()\n{\n"
              "vec4 t = texture(uTex,vUV).zyxw;\nFragColor = t * vec4(0.9, 0.8, 0.7, 1.0);\n}\n")

out = {}
sc = S.config(3)
out["built_in_texture_shape_us"] = probe(sc, reps=20) * 1e6
sc2 = S.config(3)
sc2.fs = GENERIC_FS
sc2.name += "_generic_fs"
out["generic_fs_compiled_us"] = probe(sc2, reps=20) * 1e6
out["fs_kind_compiled"] = api.swglGetOption(b"last_fs_kind")
out["generic_fs_interpreter_us"] = probe(sc2, reps=20, options={"jit": 0}) * 1e6
out["fs_kind_interpreter"] = api.swglGetOption(b"last_fs_kind")
out["jit_compiles"] = api.swglGetOption(b"jit_compiles")
out["jit_compile_ms_total"] = api.swglGetOption(b"jit_compile_us_total") / 1e3
out["compiled_over_built_in"] = out["generic_fs_compiled_us"] / out["built_in_texture_shape_us"]
out["interpreter_over_compiled"] = out["generic_fs_interpreter_us"] / out["generic_fs_compiled_us"]
print(json.dumps(out))
