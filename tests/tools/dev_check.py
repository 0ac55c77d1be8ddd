"""Developer check (GPU box): error / iteration table of every solver x formulation on C1-sized input."""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np
import c_oracle as c
import gml_b200
from gml_b200 import B200, RISE, RPLE, logRISE
from helpers import histogram_c1

FORMS = {"RISE": RISE, "logRISE": logRISE, "RPLE": RPLE}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
solvers = sys.argv[3].split(",") if len(sys.argv) > 3 else ["newton", "fista_cc", "fista_tc"]
_, hist = histogram_c1(n=n, m_samples=m, seed=16)
print("K =", hist.shape[0], "N =", n)
for form in FORMS:
    t = time.time(); ref, rinfo = c.learn_pairwise(hist, form, return_info=True); tc = time.time() - t
    for solver in solvers:
        meth = B200(solver=solver, tol=0.0 if solver == "newton" else 1e-7, max_iter=3000, verbose=int(__import__('os').environ.get('GML_VERBOSE','0')))
        try:
            t = time.time()
            got, info = gml_b200.learn(hist, FORMS[form](), meth, return_info=True)
            dt = time.time() - t
            err = np.abs(got - ref).max()
            oerr = np.abs(info["objective"] / rinfo["objective"] - 1).max()
            print(f"{form:8s} {solver:9s} err {err:.2e} objrel {oerr:.2e} it {info['iterations']} fg {info['n_fg_passes']} f {info['n_f_passes']} "
                  f"solve_ms {info['solve_ms']:.1f} wall {dt*1e3:.0f}ms launches {info['kernel_launches']} resid {info['max_residual']:.1e} (oracle {tc:.2f}s)")
        except gml_b200.GMLB200Error as e:
            print(f"{form:8s} {solver:9s} FAILED {e} stats {meth.last_stats}")
