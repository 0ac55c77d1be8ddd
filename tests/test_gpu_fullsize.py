"""Full-size checks (BASELINE config C2: N=100 10x10 lattice spin glass, 1e6 samples) through size-independent
properties: KKT conditions of the returned point verified with an INDEPENDENT gradient (the CUDA-core backend),
node-shard consistency, recovery of the generating couplings, symmetry."""
import ctypes

import numpy as np
import pytest

import gml_b200
from gml_b200 import B200, RISE, RPLE, _lib

pytestmark = pytest.mark.gpu


def lattice_model(side=10, coupling=0.4, seed=100):
    rng = np.random.default_rng(seed)
    n = side * side
    truth = np.zeros((n, n))
    for r in range(side):
        for c in range(side):
            i = r * side + c
            for j in ([i + 1] if c + 1 < side else []) + ([i + side] if r + 1 < side else []):
                truth[i, j] = truth[j, i] = coupling * rng.choice([-1.0, 1.0])
    row_ptr = np.zeros(n + 1, dtype=np.int32)
    col, val = [], []
    for i in range(n):
        nz = np.nonzero(truth[i])[0]
        row_ptr[i + 1] = row_ptr[i] + len(nz)
        col += nz.tolist(); val += truth[i, nz].tolist()
    return truth, row_ptr, np.array(col, dtype=np.int32), np.array(val, dtype=np.float32)


@pytest.fixture(scope="module")
def c2_session():
    import torch
    truth, row_ptr, col, val = lattice_model()
    n, k = 100, 1_000_000
    spins = torch.empty((n, k), dtype=torch.int8, device="cuda")
    counts = torch.ones(k, dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().gml_b200_sample_gibbs_device(0, n, row_ptr.ctypes.data, col.ctypes.data, val.ctypes.data, None,
                                                        k, 60, 100, ctypes.c_void_p(spins.data_ptr()), k, None))
    torch.cuda.synchronize()
    sess = gml_b200.Session(0).attach_device(counts.data_ptr(), spins.data_ptr(), k, n, k)
    return sess, truth, spins


def rows_to_x(theta):
    """N x N rows (diagonal = field) -> N x (N+1) coefficient rows (self coupling 0, field last)."""
    n = theta.shape[0]
    x = np.zeros((n, n + 1))
    x[:, :n] = theta
    x[:, n] = np.diag(theta)
    x[np.arange(n), np.arange(n)] = 0.0
    return x


@pytest.mark.parametrize("form_cls,c", [(RISE, 0.4), (RPLE, 0.2)])
def test_c2_kkt_recovery_symmetry(c2_session, form_cls, c):
    sess, truth, _ = c2_session
    n = 100
    theta, info = sess.solve_pairwise(form_cls(c, False), B200(), return_info=True)
    assert info["solver_used"] == 3 and info["n_unconverged"] == 0
    lam = info["lambda"]
    # recovery of the generating model (statistical tolerance at 1e6 samples) and of its support
    off = theta - np.diag(np.diag(theta))
    assert np.abs(off - truth).max() <= 0.02
    assert np.abs(np.diag(theta)).max() <= 0.02
    # KKT with the independent fp32 CUDA-core gradient: |g_j| <= lambda on zeros, g_j = -lambda sign(x_j) on the
    # support, g = 0 for the free field.  Tolerance: FISTA tol (1e-6) x curvature + fp32 gradient noise
    f, g = sess.eval_pairwise(form_cls(c, False), rows_to_x(theta), backend="fista_cc")
    x = rows_to_x(theta)
    tol = 2e-5
    for u in range(n):
        for j in range(n):
            if j == u:
                continue
            if x[u, j] == 0.0:
                assert abs(g[u, j]) <= lam + tol
            else:
                assert abs(g[u, j] + lam * np.sign(x[u, j])) <= tol
        assert abs(g[u, n]) <= tol
    # objective reported by the solver = f + lambda |x|_1 with the independent f
    obj = f + lam * (np.abs(x[:, :n]).sum(axis=1))
    assert np.allclose(info["objective"], obj, rtol=2e-6, atol=0)
    # symmetrised output is symmetric and equals 0.5 (R + R')
    sym = sess.solve_pairwise(form_cls(c, True), B200())
    assert np.array_equal(sym, sym.T)
    assert np.abs(sym - 0.5 * (theta + theta.T)).max() <= 1e-12


def test_c2_node_shards_equal_full_solve(c2_session):
    """Node problems are independent: solving shards [0,37) and [37,100) gives the rows of the full solve, up to the
    solver tolerance (the round at which the passes switch from the coarse to the fine precision level is decided by
    the slowest node of the solve, so shard and full solves follow slightly different paths to the same optimum).
    With the coarse level disabled the rows agree to 1e-9."""
    import torch
    sess, _, _ = c2_session
    n = 100
    full = sess.solve_pairwise(RISE(0.4, False), B200())
    rows = []
    for b, e in ((0, 37), (37, 100)):
        out = torch.empty((e - b, n), dtype=torch.float64, device="cuda")
        sess.solve_pairwise_device(RISE(0.4, False), B200(), out.data_ptr(), b, e)
        torch.cuda.synchronize()
        rows.append(out.cpu().numpy())
    assert np.abs(np.vstack(rows) - full).max() <= 5e-6
    exact = B200(coarse_level=False)
    full2 = sess.solve_pairwise(RISE(0.4, False), exact)
    rows2 = []
    for b, e in ((0, 37), (37, 100)):
        out = torch.empty((e - b, n), dtype=torch.float64, device="cuda")
        sess.solve_pairwise_device(RISE(0.4, False), exact, out.data_ptr(), b, e)
        torch.cuda.synchronize()
        rows2.append(out.cpu().numpy())
    assert np.abs(np.vstack(rows2) - full2).max() <= 1e-9


def test_c2_active_set_compaction_changes_nothing(c2_session):
    """The passes drop parked / converged nodes (compacted slot list, gathered spin tiles).  Node problems are
    independent and the contractions exact, so at one precision level the compacted solve returns the rows of the
    uncompacted one; with the coarse level the paths differ only in when nodes park (solver tolerance)."""
    sess, _, _ = c2_session
    a, ia = sess.solve_pairwise(RISE(0.4, False), B200(coarse_level=False, compaction=True), return_info=True)
    b, ib = sess.solve_pairwise(RISE(0.4, False), B200(coarse_level=False, compaction=False), return_info=True)
    assert ia["n_unconverged"] == 0 and ib["n_unconverged"] == 0
    assert np.abs(a - b).max() <= 1e-9
    assert ia["evals"] < 0.9 * ib["evals"]          # the compacted passes did less work
    c = sess.solve_pairwise(RISE(0.4, False), B200(compaction=True))
    d = sess.solve_pairwise(RISE(0.4, False), B200(compaction=False))
    # (two solves that each stop at a prox-gradient mapping of 1e-6 -- on different precision levels -- are each within
    # ~tol / mu of the optimum)
    assert np.abs(c - d).max() <= 5e-6 and np.abs(c - a).max() <= 1e-5


def test_device_histogram_builder_and_sampler():
    """raw Gibbs samples (N=16, M=1e6) -> device dedup -> learn; the device histogram equals the
    host np.unique histogram, and learn() on it recovers the generating model (test/runtests.jl:105-127 style)."""
    import torch
    from helpers import random_ising
    n, m = 16, 1_000_000
    model = random_ising(n, 16)
    row_ptr = np.zeros(n + 1, dtype=np.int32); col, val = [], []
    for i in range(n):
        nz = [j for j in range(n) if j != i and model[i, j] != 0.0]
        row_ptr[i + 1] = row_ptr[i] + len(nz); col += nz; val += [model[i, j] for j in nz]
    col, val = np.array(col, dtype=np.int32), np.array(val, dtype=np.float32)
    field = np.ascontiguousarray(np.diag(model), dtype=np.float32)
    lib = _lib.load()
    raw = torch.empty((n, m), dtype=torch.int8, device="cuda")
    _lib.check(lib.gml_b200_sample_gibbs_device(0, n, row_ptr.ctypes.data, col.ctypes.data, val.ctypes.data, field.ctypes.data,
                                                m, 60, 5, ctypes.c_void_p(raw.data_ptr()), m, None))
    out_spins = torch.empty((n, m), dtype=torch.int8, device="cuda")
    out_counts = torch.empty(m, dtype=torch.float64, device="cuda")
    k = ctypes.c_int64(0)
    _lib.check(lib.gml_b200_build_histogram_device(0, ctypes.c_void_p(raw.data_ptr()), m, n, m, ctypes.c_void_p(out_spins.data_ptr()),
                                                   m, ctypes.c_void_p(out_counts.data_ptr()), ctypes.byref(k), None))
    K = k.value
    assert 1000 < K <= 65536
    host = raw.cpu().numpy()
    keys = (host.T > 0).astype(np.uint64) @ (np.uint64(1) << np.arange(n, dtype=np.uint64))
    uk, uc = np.unique(keys, return_counts=True)
    assert K == len(uk)
    dev_spins = out_spins[:, :K].cpu().numpy()
    dev_keys = (dev_spins.T > 0).astype(np.uint64) @ (np.uint64(1) << np.arange(n, dtype=np.uint64))
    assert np.array_equal(dev_keys, uk) and np.array_equal(out_counts[:K].cpu().numpy(), uc.astype(np.float64))
    hist = np.concatenate([out_counts[:K].cpu().numpy()[:, None], dev_spins.T.astype(np.float64)], axis=1)
    learned = gml_b200.learn(hist, RISE())
    assert np.abs(learned - model).max() <= 0.02


def test_regularisation_path_warm_starts(c2_session):
    """a lambda path with warm starts gives the same matrices as independent cold solves and needs
    fewer passes for the later points."""
    sess, _, _ = c2_session
    cs = [0.8, 0.4, 0.2]
    m = B200(solver="fista_tc")
    path = sess.solve_path(RISE(0.4, True), m, cs)
    warm_passes = m.last_stats["n_fg_passes"]
    cold_passes = 0
    for i, c in enumerate(cs):
        mc = B200(solver="fista_tc")
        cold = sess.solve_pairwise(RISE(c, True), mc)
        cold_passes += mc.last_stats["n_fg_passes"]
        assert np.abs(path[i] - cold).max() <= 5e-6
    print("path passes warm", warm_passes, "cold", cold_passes)
    assert warm_passes <= cold_passes + 2


def test_node_chunking_when_memory_is_short(monkeypatch):
    """A shard whose residual-limb buffer would not fit is solved in consecutive node chunks (here forced through the
    GML_B200_MEM_BUDGET_GB test hook: 300 nodes -> chunks of 128) and must give the same rows."""
    rng = np.random.default_rng(9)
    n, k = 300, 50_000
    spins = rng.choice(np.array([-1, 1], dtype=np.int8), size=(n, k))
    counts = np.ones(k)
    sess = gml_b200.Session(0).upload(counts, spins)
    # (node chunks cannot use the mean-field start -- it needs the correlation rows of all nodes -- so the comparison runs cold)
    whole = sess.solve_pairwise(RISE(0.4, False), B200(coarse_level=False, warm_start=False))
    monkeypatch.setenv("GML_B200_MEM_BUDGET_GB", "0.05")
    m = B200(coarse_level=False, warm_start=False)
    chunked = sess.solve_pairwise(RISE(0.4, False), m)
    monkeypatch.delenv("GML_B200_MEM_BUDGET_GB")
    assert m.last_stats["n_fg_passes"] > 0
    assert np.abs(chunked - whole).max() <= 1e-9
