"""Developer ablation run (GPU box, -DGML_TC_ABLATE build): which pipeline stage bounds the contraction kernels.
GML_B200_DBG bits: 1 no epilogue, 2 no limb-tile loads, 4 no operand loads, 8 no histogram-tile loads (energy)."""
import sys, pathlib, ctypes, os
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import gml_b200
from gml_b200 import _lib
from bench import c3_model
n = int(sys.argv[1]); k = int(float(sys.argv[2])); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
lib = _lib.load()
row_ptr, col, val, _ = c3_model(n)
spins = torch.empty((n, k), dtype=torch.int8, device="cuda")
counts = torch.ones(k, dtype=torch.float64, device="cuda")
_lib.check(lib.gml_b200_sample_gibbs_device(0, n, row_ptr.ctypes.data, col.ctypes.data, val.ctypes.data, None, k, 5, 1,
                                            ctypes.c_void_p(spins.data_ptr()), k, None))
sess = gml_b200.Session(0).attach_device(counts.data_ptr(), spins.data_ptr(), k, n, k)
fl = 2.0 * k * (n + 1) * n
for dbg in [int(a) for a in (sys.argv[4].split(",") if len(sys.argv) > 4 else "0,1,2,8,4,5".split(","))]:
    os.environ["GML_B200_DBG"] = str(dbg)
    for coarse in ((True,) if os.environ.get('LEVELS') == 'coarse' else (False, True)):
        r = sess.bench_passes(gml_b200.RISE(), "fista_tc", reps, coarse=coarse)
        print(f"dbg={dbg} {'coarse' if coarse else 'fine  '}: " + "  ".join(f"{a} {b:.3f} ms ({fl / b / 1e9:.0f} TF/s alg)" for a, b in list(r.items())[:3]), flush=True)
