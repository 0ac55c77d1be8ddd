"""Developer check (GPU box): f / gradient of the cc and tc backends against the float64 oracle."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np
import gml_oracle as o
import gml_b200
from gml_b200 import RISE, RPLE, logRISE
from helpers import histogram_c1

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
backends = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fista_cc", "fista_tc"]
_, hist = histogram_c1(n=n, m_samples=m, seed=16)
K, N, M = o.data_info(hist)
print("K", K, "N", N, flush=True)
rng = np.random.default_rng(1)
x = rng.normal(size=(N, N + 1)) * 0.3 * (rng.random((N, N + 1)) < 0.4)
x = np.round(x * 2**24) / 2**24
for u in range(N):
    x[u, u] = 0.0
counts, spins = gml_b200.pack_histogram(hist)
sess = gml_b200.Session().upload(counts, spins)
w = hist[:, 0] / M
for name, cls in (("RISE", RISE), ("logRISE", logRISE), ("RPLE", RPLE)):
    fr = np.zeros(N); gr = np.zeros((N, N + 1))
    for u in range(N):
        stat = o.nodal_stat_pairwise(hist.astype(float), u)
        xv = x[u, :N].copy(); xv[u] = x[u, N]
        f, g, _ = o.smooth_parts(name, xv, stat, w, hess=False)
        fr[u] = f; gr[u, :N] = g; gr[u, N] = g[u]; gr[u, u] = 0.0
    for be in backends:
        f, g = sess.eval_pairwise(cls(), x, be)
        print(f"{name:8s} {be:9s} f relerr {np.abs(f / fr - 1).max():.2e}  g abserr {np.abs(g - gr).max():.2e} (|g|max {np.abs(gr).max():.2e})", flush=True)
