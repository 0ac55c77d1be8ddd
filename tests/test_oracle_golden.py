"""Pins the CPU oracle (numpy + C twin) to the reference's own known-answer fixtures
(test/runtests.jl:68-101 -> tests/golden/*.csv) and relational tests (test/runtests.jl:132-158)."""
import numpy as np
import pytest

import c_oracle as c
import gml_oracle as o

FORMS = ("RISE", "logRISE", "RPLE")


@pytest.mark.parametrize("name", ["a", "b", "c"])
@pytest.mark.parametrize("form", FORMS)
def test_abc_goldens(golden, name, form):
    s = golden(f"{name}_samples.csv")
    gold = golden(f"{name}_{form}_learned.csv")
    exact = c.learn_pairwise(s, form)                      # default c, symmetrised (test/common.jl:9-13)
    barrier = c.learn_pairwise(s, form, mode="barrier", mu=1e-9)
    assert np.abs(exact - gold).max() <= 2e-8              # exact L1 vs Ipopt: barrier bias only
    assert np.abs(barrier - gold).max() <= 2e-9
    # the reference's own criterion: isapprox => relative Frobenius <= sqrt(eps)
    assert np.linalg.norm(barrier - gold) <= np.sqrt(np.finfo(float).eps) * np.linalg.norm(gold)


@pytest.mark.parametrize("name", ["a", "c"])
@pytest.mark.parametrize("form", FORMS)
def test_numpy_and_c_oracles_agree(golden, name, form):
    s = golden(f"{name}_samples.csv")
    assert np.abs(o.learn_pairwise(s, form) - c.learn_pairwise(s, form)).max() <= 1e-11
    assert np.abs(o.learn_pairwise(s, form, mode="barrier") - c.learn_pairwise(s, form, mode="barrier")).max() <= 1e-11


@pytest.mark.parametrize("form", FORMS)
def test_mvt_goldens(golden, form):
    """mvt (lambda = 5.4e-5): Ipopt's barrier point differs from the exact L1 optimum by up to 1.6e-4;
    the barrier restatement is within the north-star tolerance 1e-4 (test/runtests.jl:83-101)."""
    s = golden("mvt_samples.csv")
    gold = golden(f"mvt_{form}_learned.csv")
    exact = c.learn_pairwise(s, form, 0.2, False)
    barrier = c.learn_pairwise(s, form, 0.2, False, mode="barrier", mu=1e-9)
    assert np.abs(exact - gold).max() <= 2e-4
    assert np.abs(barrier - gold).max() <= 1e-4


@pytest.mark.parametrize("name", ["a", "b", "c", "mvt"])
def test_multirise_order2_equals_rise(golden, name):
    """test/runtests.jl:132-158: multiRISE(0.2,false,2) == RISE(0.2,false) key by key."""
    s = golden(f"{name}_samples.csv")
    # barrier mode = what Ipopt returns: no exact zeros, so no key is dropped by convert(Dict, .)
    r = o.matrix_to_dict(c.learn_pairwise(s, "RISE", 0.2, False, mode="barrier"))
    m = c.learn_multibody(s, 0.2, False, 2, mode="barrier")
    assert len(r) == len(m)
    for k, v in r.items():
        assert abs(m[k] - v) <= 1e-7


def test_multirise_order3_numpy_vs_c(golden):
    s = golden("c_samples.csv")
    a, b = o.learn_multibody(s, 0.2, True, 3), c.learn_multibody(s, 0.2, True, 3)
    assert a.keys() == b.keys()
    assert max(abs(a[k] - b[k]) for k in a) <= 1e-11


def test_objective_known_answers(golden):
    """Provisional objective KATs of SURVEY 8c re-derived: node 1, exact L1, lambda term included."""
    s = golden("a_samples.csv")
    for form, want in (("RISE", 0.969954153785), ("logRISE", -0.030242641570), ("RPLE", 0.663177462752)):
        _, info = c.learn_pairwise(s, form, return_info=True)
        assert abs(info["objective"][0] - want) <= 1e-9


@pytest.mark.parametrize("form", FORMS)
def test_c_eval_pairwise_matches_smooth_parts(form):
    """The histogram-sweeping f / grad f entry (gml_oracle_eval_pairwise) is the same function as the per-node
    float64 restatement smooth_parts over nodal_stat (src/GraphicalModelLearning.jl:162, 170/279/317)."""
    rng = np.random.default_rng(11)
    n, k = 23, 5003
    spins = rng.choice(np.array([-1, 1], dtype=np.int8), size=(n, k))
    counts = rng.integers(1, 6, size=k).astype(np.float64)
    hist = np.concatenate([counts[:, None], spins.T.astype(np.float64)], axis=1)
    w = counts / counts.sum()
    x = rng.normal(size=(n, n + 1)) * 0.2 * (rng.random((n, n + 1)) < 0.4)
    nodes = np.array([0, 7, 22, 13], dtype=np.int32)
    f, g = c.eval_pairwise(counts, spins, form, x[nodes], nodes)
    for q, u in enumerate(nodes):
        stat = o.nodal_stat_pairwise(hist, u)
        xv = x[u, :n].copy(); xv[u] = x[u, n]
        fu, gu, _ = o.smooth_parts(form, xv, stat, w, hess=False)
        gr = np.zeros(n + 1); gr[:n] = gu; gr[n] = gu[u]; gr[u] = 0.0
        assert abs(f[q] - fu) <= 1e-13 * max(1.0, abs(fu))
        assert np.abs(g[q] - gr).max() <= 1e-13
    f2, _ = c.eval_pairwise(counts, spins, form, x[nodes], nodes, want_grad=False)
    assert np.array_equal(f, f2)


def test_c_oracle_inner_threads_and_node_ranges():
    """Few nodes of a larger problem use the host threads inside the Hessian accumulation: same answer as the
    node-parallel solve; learn_multibody(nodes=...) returns the rows of the full solve."""
    rng = np.random.default_rng(5)
    n, k = 12, 20000
    spins = rng.choice(np.array([-1, 1], dtype=np.int8), size=(n, k))
    spins[3] = spins[2] * np.where(rng.random(k) < 0.8, 1, -1).astype(np.int8)
    counts = np.ones(k)
    c.set_threads(4)
    lam = 0.01
    full = c.learn_pairwise_packed(counts, spins, "RISE", lam, False)
    part = c.learn_pairwise_packed(counts, spins, "RISE", lam, False, nodes=(2, 4))     # 2 nodes x 2 inner threads
    assert np.abs(part[2:4] - full[2:4]).max() <= 1e-11
    hist = np.concatenate([counts[:, None], spins.T.astype(np.float64)], axis=1)[:4000, :7]
    a = c.learn_multibody(hist, 0.3, False, 3)
    b = c.learn_multibody(hist, 0.3, False, 3, nodes=(1, 3))
    assert set(b) == {k_ for k_ in a if k_[0] in (2, 3)}
    assert max(abs(a[k_] - b[k_]) for k_ in b) <= 1e-11
