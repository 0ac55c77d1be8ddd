"""Developer micro-benchmark (GPU box): per-kernel ms of the contraction passes on a synthetic histogram."""
import sys, pathlib, ctypes, os
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import gml_b200
from gml_b200 import _lib
sys.path.insert(0, str(ROOT))
from bench import c3_model
n = int(sys.argv[1]); k = int(float(sys.argv[2])); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
forms = sys.argv[4].split(",") if len(sys.argv) > 4 else ["RISE"]
nb, ne = (int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (0, 0)
lib = _lib.load()
row_ptr, col, val, _ = c3_model(n)
spins = torch.empty((n, k), dtype=torch.int8, device="cuda")
counts = torch.ones(k, dtype=torch.float64, device="cuda")
_lib.check(lib.gml_b200_sample_gibbs_device(0, n, row_ptr.ctypes.data, col.ctypes.data, val.ctypes.data, None, k, 5, 1,
                                            ctypes.c_void_p(spins.data_ptr()), k, None))
sess = gml_b200.Session(0).attach_device(counts.data_ptr(), spins.data_ptr(), k, n, k)
F = n + 1
nn = (ne - nb) if ne else n
for name in forms:
    r = sess.bench_passes(getattr(gml_b200, name)(), "fista_tc", reps, nb, ne)
    fl = 2.0 * k * F * nn
    print(f"{os.environ.get('TAG','')} {name} N={n} K={k} nodes={nn}: " + "  ".join(f"{a} {b:.3f} ms ({fl / b / 1e9:.0f} TF/s alg)" for a, b in list(r.items())[:3])
          + f"  full pass wall {r['full_pass_wall']:.3f} ms", flush=True)
