"""Developer A/B (GPU box): FISTA rounds / time with and without the experimental mean-field warm start
(B200(warm_start=True), opts.reserved[7]) on a synthetic C3-style histogram, plus agreement of the two solutions.
    python scripts/dev_warmstart_gpu.py [N=1000] [K=2e6]"""
import ctypes, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import gml_b200
from gml_b200 import _lib
from bench import c3_model
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
k = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
lib = _lib.load()
row_ptr, col, val, truth = c3_model(n)
spins = torch.empty((n, k), dtype=torch.int8, device="cuda")
counts = torch.ones(k, dtype=torch.float64, device="cuda")
_lib.check(lib.gml_b200_sample_gibbs_device(0, n, row_ptr.ctypes.data, col.ctypes.data, val.ctypes.data, None, k, 40, 1000,
                                            ctypes.c_void_p(spins.data_ptr()), k, None))
sess = gml_b200.Session(0).attach_device(counts.data_ptr(), spins.data_ptr(), k, n, k)
out = {}
for warm in (False, True, False, True):
    m = gml_b200.B200(solver="fista_tc", warm_start=warm, verbose=1 if warm else 0)
    torch.cuda.synchronize(); t0 = time.time()
    theta, info = sess.solve_pairwise(gml_b200.RISE(0.4, False), m, return_info=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    out[warm] = theta
    print(f"warm_start={warm}: {dt * 1e3:.1f} ms, rounds {info['iterations']}, fg passes {info['n_fg_passes']}, evals {info['evals']:.3e}, "
          f"unconverged {info['n_unconverged']}, residual {info['max_residual']:.2e}", flush=True)
print("max |theta_warm - theta_cold| =", np.abs(out[True] - out[False]).max())
