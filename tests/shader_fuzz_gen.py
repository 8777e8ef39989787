"""Random well-typed shader pairs inside the executable GLSL subset of the reference (SURVEY.md appendix B:
`sub (op sub)*` folded left to right, operands of identical type, constructors with exactly N scalar arguments,
swizzles, sin / cos / tan / min / max, float / vec2 / vec3 / vec4 locals, uniforms and varyings of every width).
Shared by tests/test_shaders_gpu.py (a few seeds) and tools/shader_fuzz.py (many).

Kept out on purpose (non-reproducible corners): comparison operators, reads of never-written variables,
operands of different types (the statement is silently skipped and the variable keeps whatever an earlier invocation
left in its static storage), a first `out` that is not a vec4, division by anything but a constant."""
import numpy as np

WIDTH = {"float": 1, "vec2": 2, "vec3": 3, "vec4": 4}
TYPES = ["float", "vec2", "vec3", "vec4"]
COMPS = "xyzw"


class _Gen:
    def __init__(self, rng, env):
        self.rng = rng
        self.env = dict(env)          # name -> type, every one of them written before the fragment stage reads it

    def lit(self, lo=0.05, hi=1.5, signed=True):
        v = float(self.rng.uniform(lo, hi))
        if signed and self.rng.random() < 0.25:
            v = -v
        return f"{v:.4f}"

    def scalar(self):
        """A scalar sub-expression: float variable, one component of a vector, or a literal."""
        r = self.rng.random()
        names = sorted(self.env)
        if r < 0.25:
            return self.lit()
        n = names[int(self.rng.integers(len(names)))]
        w = WIDTH[self.env[n]]
        return n if w == 1 else f"{n}.{COMPS[int(self.rng.integers(w))]}"

    def leaf(self, ty):
        w = WIDTH[ty]
        r = self.rng.random()
        same = sorted(n for n, t in self.env.items() if t == ty)
        wider = sorted(n for n, t in self.env.items() if WIDTH[t] >= 2 and WIDTH[t] >= w)
        if r < 0.35 and same:
            return same[int(self.rng.integers(len(same)))]
        if r < 0.7 and wider:
            n = wider[int(self.rng.integers(len(wider)))]
            nw = WIDTH[self.env[n]]
            return n + "." + "".join(COMPS[int(self.rng.integers(nw))] for _ in range(w))
        if w == 1:
            return self.scalar()
        return f"{ty}(" + ", ".join(self.scalar() for _ in range(w)) + ")"

    def const(self, ty, lo, hi):
        w = WIDTH[ty]
        if w == 1:
            return self.lit(lo, hi, signed=False)
        return f"{ty}(" + ", ".join(self.lit(lo, hi, signed=False) for _ in range(w)) + ")"

    def sub(self, ty, depth):
        r = self.rng.random()
        if depth <= 0 or r < 0.4:
            return self.leaf(ty)
        if r < 0.6:
            inner = self.expr(ty, depth - 1)
            return f"({inner})"
        if r < 0.8:
            fn = ["sin", "cos", "tan"][int(self.rng.integers(3))] if self.rng.random() < 0.8 else "sin"
            return f"{fn}({self.expr(ty, depth - 1)})"
        fn = "min" if self.rng.random() < 0.5 else "max"
        return f"{fn}({self.expr(ty, depth - 1)}, {self.leaf(ty)})"

    def expr(self, ty, depth):
        n_ops = int(self.rng.integers(0, 3))
        s = self.sub(ty, depth)
        for _ in range(n_ops):
            op = "+-*/"[int(self.rng.integers(4))]
            if op == "/":
                s += " / " + self.const(ty, 0.5, 4.0)
            else:
                s += f" {op} " + self.sub(ty, depth - 1)
        return s


def make(seed):
    """-> (vs, fs, uniforms) in the form of tests/shader_cases.CASES."""
    rng = np.random.default_rng(77000 + seed)
    uniforms, decl_u = {}, ""
    for ty, kind in (("float", "1f"), ("vec2", "2f"), ("vec3", "3f"), ("vec4", "4f")):
        if rng.random() < 0.6:
            name = "u" + kind[0]
            uniforms[name] = (kind, [round(float(v), 4) for v in rng.uniform(0.1, 1.5, WIDTH[ty])])
            decl_u += f"uniform {ty} {name};\n"

    # vertex stage: pass-through position (sometimes scaled), a vec4 varying and up to three more of random width
    vary = [("vCol", "vec4")]
    for k in range(int(rng.integers(0, 4))):
        vary.append((f"v{k}", TYPES[int(rng.integers(4))]))
    vs = "layout (location = 0) vec4 aPos;\nlayout (location = 1) vec4 aCol;\n"
    vs += "".join(f"out {t} {n};\n" for n, t in vary) + "void main()\n{\n"
    if rng.random() < 0.4:
        vs += "gl_Position = aPos * vec4(0.9, 0.8, 1.0, 1.0);\n"
    else:
        vs += "gl_Position = aPos;\n"
    gv = _Gen(rng, {"aCol": "vec4", "aPos": "vec4"})
    vs += "vCol = aCol;\n"
    for n, t in vary[1:]:
        gv.env = {"aCol": "vec4"}
        vs += f"{n} = {gv.expr(t, 1)};\n"
    vs += "}\n"

    # fragment stage: a handful of typed locals, then a vec4 result
    env = {n: t for n, t in vary}
    for name, (kind, _) in uniforms.items():
        env[name] = TYPES[int(kind[0]) - 1]
    g = _Gen(rng, env)
    fs = "".join(f"in {t} {n};\n" for n, t in vary) + decl_u + "out vec4 FragColor;\nvoid main()\n{\n"
    for k in range(int(rng.integers(2, 7))):
        ty = TYPES[int(rng.integers(4))]
        fs += f"{ty} t{k} = {g.expr(ty, 2)};\n"
        g.env[f"t{k}"] = ty
    # the result reads the locals (and the colour varying) only, wrapped into [0, 1] by the reference's parabola sine so
    # that the clamp does not flatten it; every other program blends with a computed alpha
    g.env = {n: t for n, t in g.env.items() if n.startswith("t") or n == "vCol"}
    fs += f"vec4 res = sin({g.expr('vec4', 2)}) * vec4(0.5, 0.5, 0.5, 0.5) + vec4(0.5, 0.5, 0.5, 0.5);\n"
    fs += "FragColor = vec4(res.x, res.y, res.z, " + ("res.w" if rng.random() < 0.5 else "1.0") + ");\n}\n"
    return vs, fs, uniforms
