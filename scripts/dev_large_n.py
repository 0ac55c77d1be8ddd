"""Developer check (GPU box): tensor-core vs CUDA-core backend at a large N (f, gradient agreement, no crash)."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import gml_b200
from gml_b200 import RISE, RPLE
n = int(sys.argv[1]); k = int(float(sys.argv[2]))
rng = np.random.default_rng(0)
spins = rng.choice(np.array([-1, 1], dtype=np.int8), size=(n, k))
counts = np.ones(k)
x = rng.normal(size=(n, n + 1)) * 0.05 * (rng.random((n, n + 1)) < 0.02)
x = np.round(x * 2**24) / 2**24
x[np.arange(n), np.arange(n)] = 0
sess = gml_b200.Session().upload(counts, spins)
for cls in (RISE, RPLE):
    f1, g1 = sess.eval_pairwise(cls(), x, "fista_cc")
    f2, g2 = sess.eval_pairwise(cls(), x, "fista_tc")
    print(cls.__name__, "N", n, "K", k, "f rel diff %.2e" % np.abs(f1 / f2 - 1).max(), "g abs diff %.2e" % np.abs(g1 - g2).max(), "|g|max %.2e" % np.abs(g2).max(), flush=True)
th = sess.solve_pairwise(RISE(0.4, True), gml_b200.B200())
print("solve ok, nnz/row", (th != 0).sum(1).mean())
