"""Developer trace (GPU box): per-round FISTA diagnostics for one formulation on a C1-sized input."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(p))
import gml_b200
from gml_b200 import B200
from helpers import histogram_c1
form = getattr(gml_b200, sys.argv[1]); solver = sys.argv[2]; tol = float(sys.argv[3])
_, hist = histogram_c1(n=16, m_samples=100000, seed=16)
try:
    gml_b200.learn(hist, form(), B200(solver=solver, tol=tol, max_iter=400, verbose=2))
except Exception as e:
    print("ERR", e)
