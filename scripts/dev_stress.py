"""Developer stress (GPU box): regimes off the benchmark path -- strong couplings (coarse-range overflow), logRISE at
scale, few samples.  Checks KKT with the independent CUDA-core gradient."""
import sys, pathlib, ctypes, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np, torch
import gml_b200
from gml_b200 import B200, RISE, RPLE, logRISE, _lib
from test_gpu_fullsize import lattice_model, rows_to_x

def kkt(sess, form, theta, lam, n):
    x = rows_to_x(theta)
    f, g = sess.eval_pairwise(form, x, backend="fista_cc")
    off = ~np.eye(n, dtype=bool)
    gz = np.abs(g[:, :n])[off & (x[:, :n] == 0)]
    gn = np.abs(g[:, :n] + lam * np.sign(x[:, :n]))[off & (x[:, :n] != 0)]
    return max(0.0, gz.max() - lam if gz.size else 0.0), gn.max() if gn.size else 0.0, np.abs(g[:, n]).max()

def run(name, coupling, side, k, form, sweeps=60):
    truth, row_ptr, col, val = lattice_model(side, coupling)
    n = side * side
    spins = torch.empty((n, k), dtype=torch.int8, device="cuda")
    counts = torch.ones(k, dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().gml_b200_sample_gibbs_device(0, n, row_ptr.ctypes.data, col.ctypes.data, val.ctypes.data, None, k, sweeps, 7,
                                                        ctypes.c_void_p(spins.data_ptr()), k, None))
    sess = gml_b200.Session(0).attach_device(counts.data_ptr(), spins.data_ptr(), k, n, k)
    m = B200(verbose=1)
    t = time.time()
    try:
        theta, info = sess.solve_pairwise(form, m, return_info=True)
    except gml_b200.GMLB200Error as e:
        print(name, "FAILED", e, m.last_stats); return
    dt = time.time() - t
    a, b, c = kkt(sess, form, theta, info["lambda"], n)
    off = theta - np.diag(np.diag(theta))
    print(f"{name}: rounds {info['iterations']} time {dt*1e3:.0f} ms  max|x| {np.abs(theta).max():.3f}  err vs truth {np.abs(off - truth).max():.4f}  "
          f"KKT zero-excess {a:.1e} support {b:.1e} field {c:.1e}", flush=True)

run("strong J=1.0 RISE N=100 K=1e6", 1.0, 10, 1_000_000, RISE(0.4, False))
run("strong J=1.2 RPLE N=100 K=1e6", 1.2, 10, 1_000_000, RPLE(0.2, False))
run("logRISE N=100 K=1e6", 0.4, 10, 1_000_000, logRISE(0.8, False))
run("few samples N=400 K=3000 RISE", 0.4, 20, 3000, RISE(0.4, False))
run("N=900 K=2e5 logRISE", 0.4, 30, 200_000, logRISE(0.8, False))
