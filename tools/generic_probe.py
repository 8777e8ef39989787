"""Development: cost of the generic (IR evaluator) shader path against the recognised shapes on C2."""
import sys
sys.path.insert(0, ".")
from tools.perf_probe import probe
from swgl_b200 import scenes as S

sc = S.config(2)
probe(sc, reps=10)
sc2 = S.config(2)
sc2.name += "_generic_fs"
sc2.fs = "in vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = vCol + vCol;\n}\n"
probe(sc2, reps=10)
sc3 = S.config(2)
sc3.name += "_generic_vs"
sc3.vs = ("layout (location = 0) vec4 aPos;\nlayout (location = 1) vec4 aCol;\nout vec4 vCol;\nvoid main()\n{\n"
          "gl_Position = aPos;\nvCol = aCol * aCol;\n}\n")
probe(sc3, reps=10)
