"""End-to-end learn() from the reference's own input type at the headline size (GPU box; needs ~100 GB of host memory):
the K x (N+1) column-major Float64 / Int64 `samples` matrix of learn() (src/GraphicalModelLearning.jl:69-81) goes through
gml_b200_learn_pairwise_matrix -- threaded host narrowing + H2D + validate + layouts + solve + symmetrise + D2H.

    python scripts/bench_matrix_input.py [K] [N] [dtype f64|i64] [devices]
Prints one JSON line: seconds per call (median of 3 after one warm-up), ingest seconds, bytes of the matrix.
"""
import ctypes, json, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import gml_b200
from gml_b200 import _lib
from bench import c3_model

k = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dtype = {"f64": np.float64, "i64": np.int64}[sys.argv[3] if len(sys.argv) > 3 else "f64"]
devices = int(sys.argv[4]) if len(sys.argv) > 4 else 1
lib = _lib.load()
row_ptr, col, val, truth = c3_model(n)
spins = torch.empty((n, k), dtype=torch.int8, device="cuda")
_lib.check(lib.gml_b200_sample_gibbs_device(0, n, row_ptr.ctypes.data, col.ctypes.data, val.ctypes.data, None, k, 40, 1000,
                                            ctypes.c_void_p(spins.data_ptr()), k, None))
torch.cuda.synchronize()
t0 = time.time()
mat = np.empty((k, n + 1), dtype=dtype, order="F")          # Julia-native layout
mat[:, 0] = 1
host = spins.cpu().numpy()
del spins
torch.cuda.empty_cache()
for i in range(n):
    mat[:, 1 + i] = host[i]
del host
build_s = time.time() - t0
form = gml_b200.RISE(0.4, True)
times, ingest = [], []
for it in range(4):
    m = gml_b200.B200(tol=1e-6, devices=devices)
    t0 = time.perf_counter()
    theta = gml_b200.learn_matrix(mat, form, m)
    dt = time.perf_counter() - t0
    if it:
        times.append(dt); ingest.append(m.last_stats["h2d_ms"] * 1e-3)
err = float(np.abs(theta - np.diag(np.diag(theta)) - truth).max())
st = m.last_stats
print(json.dumps({"what": "learn() from the K x (N+1) samples matrix (gml_b200_learn_pairwise_matrix)", "K": k, "N": n, "dtype": str(np.dtype(dtype)),
                  "devices": devices, "matrix_bytes": int(mat.nbytes), "learn_seconds_median": float(np.median(times)), "all_seconds": times,
                  "ingest_seconds_median": float(np.median(ingest)), "matrix_GBps": mat.nbytes / 1e9 / float(np.median(ingest)),
                  "solve_seconds": st["solve_ms"] * 1e-3, "pack_seconds": st["pack_ms"] * 1e-3, "iterations": st["iterations"],
                  "evals_per_s": st["evals"] / float(np.median(times)), "max_abs_coupling_error_vs_truth": err,
                  "host_matrix_build_seconds": build_s}))
