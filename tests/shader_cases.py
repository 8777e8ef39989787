"""Shader pairs outside the built-in shapes, shared by the GPU parity tests (tests/test_shaders_gpu.py) and
the CPU-side check that every one of them compiles to register-resident kernels (tests/test_jit.py).
Every case is inside the reproducible subset (vec4 `out`, no comparison operators, no reads
of never-written variables)."""

VS_HEAD = "layout (location = 0) vec4 aPos;\nlayout (location = 1) vec4 aCol;\n"

CASES = {
    # (vs, fs, uniforms)
    "arith_no_precedence": (
        VS_HEAD + "out vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\n}\n",
        "in vec4 vCol;\nuniform vec4 tint;\nuniform float k;\nout vec4 FragColor;\nvoid main()\n{\n"
        "vec4 t = vCol + tint * vec4(k, 0.5, 0.25, 1.0);\nFragColor = t - (vCol*tint).wzyx / vec4(2.0, 2.0, 2.0, 1.0);\n}\n",
        {"tint": ("4f", [0.2, 0.4, 0.6, 0.8]), "k": ("1f", [0.7])},
    ),
    "swizzle_constructors_minmax": (
        VS_HEAD + "out vec4 vCol;\nout vec2 vUV;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\nvUV = aCol.zy;\n}\n",
        "in vec2 vUV;\nin vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\n"
        "float a = max(vUV.x, 0.3);\nfloat b = min(vCol.w, vUV.y);\nFragColor = vec4(a, b, vCol.z, 1);\n}\n",
        {},
    ),
    "trig_parabola": (
        VS_HEAD + "out vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\n}\n",
        "in vec4 vCol;\nuniform float freq;\nout vec4 FragColor;\nvoid main()\n{\n"
        "vec4 s = sin(vCol * vec4(freq, freq, freq, freq));\nvec4 c = cos(vCol);\nFragColor = vec4(s.x, c.y, tan(vCol.z), 1.0);\n}\n",
        {"freq": ("1f", [9.5])},
    ),
    "matrix_chain_quirks": (
        VS_HEAD + "uniform mat4 A;\nuniform mat4 B;\nout vec4 vCol;\nvoid main()\n{\nmat4 M = A * B + A;\n"
        "gl_Position = M * aPos;\nvCol = aCol;\n}\n",
        "in vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = vCol;\n}\n",
        {"A": ("m4", [0.5, 0, 0, 0, 0, 0.5, 0, 0, 0.02, -0.01, 0.0, 0, 0, 0, 0.125, 0.5]),
         "B": ("m4", [1, 0.1, 0, 0, -0.1, 1, 0, 0, 0, 0, 1, 0, 0.05, 0, 0, 1])},
    ),
    "mat3_mat2_and_int": (
        VS_HEAD + "uniform mat3 N;\nuniform mat2 R;\nout vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\n"
        "vec3 n = N * aCol.xyz;\nvec2 r = R * aCol.xw;\nint two = 1 + 1;\nvCol = vec4(n.x, r.y, float(two), 1);\n}\n",
        "in vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = vCol * vec4(1.0, 1.0, 0.25, 1.0);\n}\n",
        {"N": ("m3", [0.5, 0.1, 0, 0.2, 0.6, 0.1, 0, 0.3, 0.7]), "R": ("m2", [0.8, -0.6, 0.6, 0.8])},
    ),
    "type_mismatch_is_noop": (
        VS_HEAD + "out vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\nvCol = aCol * 2.0;\n}\n",
        "in vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = vCol;\nFragColor = vCol.xyz;\n}\n",
        {},
    ),
    "texture_math": (
        VS_HEAD + "out vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\n}\n",
        "in vec4 vCol;\nuniform sampler2D uTex;\nout vec4 FragColor;\nvoid main()\n{\n"
        "vec4 t = texture(uTex,vCol.xy).zyxw;\nFragColor = t * vCol;\n}\n",
        {},
    ),
    # round 2: the rest of the executable subset (Appendix B of the survey), one feature group per case
    "mat4_subtract_mat3_product_mat2_sum": (      # SUBM with the column-3 typo, MULMM dim 3, ADDM dim 2
        VS_HEAD + "uniform mat4 A;\nuniform mat4 B;\nuniform mat3 N;\nuniform mat2 R;\nout vec4 vCol;\nvoid main()\n{\nmat4 M = A - B;\n"
        "gl_Position = M * aPos;\nmat3 P = N * N;\nvec3 n = P * aCol.xyz;\nmat2 Q = R + R;\nvec2 r = Q * aCol.yx;\nvCol = vec4(n.x, n.z, r.x, 1.0);\n}\n",
        "in vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = vCol;\n}\n",
        {"A": ("m4", [1.5, 0, 0, 0.01, 0, 1.5, 0, 0.02, 0.02, -0.01, 1.0, 0.03, 0.25, 0, 0.125, 1.5]),
         "B": ("m4", [0.5, 0.1, 0, 0, -0.1, 0.5, 0, 0, 0, 0, 0.5, 0, 0.05, 0.3, 0, 0.5]),
         "N": ("m3", [0.5, 0.1, 0, 0.2, 0.6, 0.1, 0, 0.3, 0.7]), "R": ("m2", [0.3, -0.2, 0.2, 0.3])},
    ),
    "int_arithmetic_and_conversions": (           # integer + - * /, int(float), float(int), a constructor with int arguments
        VS_HEAD + "out vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\n}\n",
        "in vec4 vCol;\nuniform int steps;\nout vec4 FragColor;\nvoid main()\n{\n"
        "int a = 7 / 2;\nint b = a * steps - 1;\nfloat q = vCol.x * 4.0;\nint c = int(q);\nfloat g = float(c);\n"
        "vec4 s = vec4(float(b), g, a, 1);\nFragColor = vCol * s / vec4(20.0, 4.0, 4.0, 1.0);\n}\n",
        {"steps": ("1i", [5])},
    ),
    "vec3_vec2_float_varyings": (                 # every varying width in one record, swizzles with repeats
        VS_HEAD + "out vec3 vN;\nout vec2 vUV;\nout float vF;\nvoid main()\n{\ngl_Position = aPos;\nvN = aCol.zxy;\nvUV = aCol.ww;\nvF = aCol.y;\n}\n",
        "in vec3 vN;\nin vec2 vUV;\nin float vF;\nout vec4 FragColor;\nvoid main()\n{\n"
        "vec3 n = vN * vN.zzx;\nFragColor = vec4(n.x, n.y + vF, vUV.y, 1.0);\n}\n",
        {},
    ),
    "uniform_vectors_and_parentheses": (          # vec2 / vec3 uniforms, signed literals, a swizzle on a parenthesis and on a builtin
        VS_HEAD + "out vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\n}\n",
        "in vec4 vCol;\nuniform vec2 shift;\nuniform vec3 gain;\nout vec4 FragColor;\nvoid main()\n{\n"
        "vec3 c = (vCol.xyz * gain).zyx;\nvec2 d = vCol.xy + shift;\n"
        "vec4 t = (vCol - vec4(0.5, 0.5, 0.5, 0.0)) * vec4(-1.0, 2.0, -0.5, 1.0);\n"
        "FragColor = max(t, vec4(c.x, d.y, c.z, 1.0)).yxzw;\n}\n",
        {"shift": ("2f", [0.25, -0.125]), "gain": ("3f", [0.9, 0.5, 1.5])},
    ),
    "vertex_stage_arithmetic": (                  # gl_Position from a local, no operator precedence in the vertex stage
        VS_HEAD + "uniform vec4 offs;\nout vec4 vCol;\nvoid main()\n{\nvec4 p = aPos * vec4(0.8, 0.8, 1.0, 1.0) + offs;\ngl_Position = p;\n"
        "vCol = sin(aCol * vec4(3.0, 3.0, 3.0, 3.0)).wzyx;\n}\n",
        "in vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\nvec4 m = min(tan(vCol), vCol.wwww);\nFragColor = vec4(m.x, m.y, m.z, 1.0);\n}\n",
        {"offs": ("4f", [0.1, -0.1, 0.0, 0.0])},
    ),
}
