"""Small problems through every device path added in round 2, meant to run under compute-sanitizer (GPU box):
CTA-pair + streaming energy kernels on both levels with several feature chunks, retirement on the coarse level,
mean-field start, support polish, Newton beyond 64 features, multiRISE on the tensor path + device symmetrisation,
term-list sampler, matrix ingest, thresholding."""
import os, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np
import gml_b200
from gml_b200 import B200, RISE, RPLE, logRISE, multiRISE
from helpers import three_body_model

rng = np.random.default_rng(0)
n, k = 160, 6001
spins = rng.choice(np.array([-1, 1], dtype=np.int8), size=(n, k))
spins[1] = spins[0] * np.where(rng.random(k) < 0.8, 1, -1).astype(np.int8)
counts = np.ones(k)
for env in ("0", "1"):
    if env == "1":
        os.environ["GML_B200_NO_PAIR"] = "1"
    sess = gml_b200.Session().upload(counts, spins)
    x = np.round(rng.normal(size=(n, n + 1)) * 0.05 * (rng.random((n, n + 1)) < 0.2) * 2 ** 14) / 2 ** 14
    for lvl in (False, True):
        for form in (RISE, logRISE, RPLE):
            sess.eval_pairwise(form(), x, "fista_tc", coarse=lvl)
    a = sess.solve_pairwise(RISE(0.4, True), B200(tol=1e-6))                 # mean-field start, coarse retirement
    b = sess.solve_pairwise(RISE(0.4, True), B200(tol=1e-7, warm_start=False))
    c = sess.solve_pairwise(logRISE(0.8, False), B200(tol=1e-5, polish=True))
    print("pair" if env == "0" else "streaming", float(np.abs(a - b).max()))
    sess.close()
os.environ.pop("GML_B200_NO_PAIR", None)
m = gml_b200.learn_packed(counts[:3000], np.ascontiguousarray(spins[:100, :3000]), RISE(0.4, False), B200(barrier_mu=1e-9))      # Newton, 101 features
terms = three_body_model(20, 3)
sp = gml_b200.sample_terms_device(terms, 20, 20000, sweeps=20, seed=1).cpu().numpy()
s2 = gml_b200.Session().upload(np.ones(sp.shape[1]), sp)
fg = s2.solve_multibody_sym(multiRISE(0.4, True, 3), B200())                # 211 base features: tensor path
theta, nnz = s2.threshold(np.asarray(a[:20, :20]), 0.01)
mat = np.empty((k, n + 1), dtype=np.float64, order="F"); mat[:, 0] = 1; mat[:, 1:] = spins.T
d = gml_b200.learn_matrix(mat, RPLE(0.2, True), B200())
print("sanitize script done", len(fg.terms), nnz, d.shape)
