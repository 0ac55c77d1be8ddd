"""CPU oracle for the `learn()` hot path of lanl-ansi/GraphicalModelLearning.jl (v0.2.2).

TEST INFRASTRUCTURE ONLY.  Nothing under `graphicalmodellearning.jl_b200/` may import this
module; only `tests/` (including the developer checks under `tests/tools/`), `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs use it, and only as the checker or as the timed
CPU baseline.

What it restates (float64 numpy; every function cites the reference lines it follows, all
relative to /root/reference/):

  * data_info                    src/GraphicalModelLearning.jl:76-81
  * lambda                       src/GraphicalModelLearning.jl:157 (=:86,213,266,304)
  * pairwise nodal statistics    src/GraphicalModelLearning.jl:162
  * RISE / logRISE / RPLE        src/GraphicalModelLearning.jl:169-177 / 278-286 / 316-324
  * row placement, symmetrise    src/GraphicalModelLearning.jl:181-186
  * multiRISE keys / objective   src/GraphicalModelLearning.jl:91-133, src/models.jl:228-246
  * multiRISE symmetrisation     src/GraphicalModelLearning.jl:135-149

The arithmetic of the reference itself lives in third-party packages that are NOT in the
reference tree and cannot run in the build container (no Julia): JuMP (compat
"~0.21, ~0.22, ~0.23, ^1", Project.toml:16) building the expression graph and Ipopt (compat
"~0.4 ... ^1", Project.toml:15; CHANGELOG.md:8 says fixtures were produced with "JuMP v1,
Ipopt v1") solving it with a primal-dual log-barrier interior point method.  No versions are
pinned (Manifest.toml is git-ignored).  Two solution concepts are therefore restated:

  mode "exact"   : the exact minimiser of  f_u(x) + lambda * sum_{j != u} |x_j|
                   (proximal Newton with a coordinate-descent inner loop, to ~1e-13).
  mode "barrier" : the point Ipopt actually returns -- the minimiser of the log-barrier
                   subproblem at its final barrier parameter mu.  Eliminating the slack z_j
                   from  lambda*z_j - mu*[log(z_j - x_j) + log(z_j + x_j)]  gives
                   z_j = eps + sqrt(eps^2 + x_j^2), eps = mu/lambda, i.e. a smooth penalty
                   with  phi'(x) = lambda * x / (eps + sqrt(eps^2 + x^2)).
                   Ipopt's default monotone schedule ends at mu = 1e-9 (or stops one step
                   earlier at ~2.5e-9 when `tol` is already met).

PINNING: `tests/test_oracle_golden.py` checks this oracle against all 12 known-answer
matrices the reference's own test-suite holds (test/data/{a,b,c,mvt}_{RISE,logRISE,RPLE}_learned.csv,
test/runtests.jl:68-101), copied verbatim as data fixtures into tests/golden/.
multiRISE at interaction_order=3 has NO stored output in the reference: parity for it is
pinned only through the order-2 identity with RISE (test/runtests.jl:132-146).
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass

import numpy as np

FORMULATIONS = ("RISE", "logRISE", "RPLE")


# --------------------------------------------------------------------------------------
# reference scalars
# --------------------------------------------------------------------------------------
def data_info(samples: np.ndarray):
    """(num_conf, num_spins, num_samples) -- src/GraphicalModelLearning.jl:76-81."""
    num_conf, num_row = samples.shape
    num_spins = num_row - 1
    num_samples = samples[:, 0].sum()
    return num_conf, num_spins, num_samples


def regularizer_lambda(c: float, num_spins: int, num_samples: float) -> float:
    """lambda = c*sqrt(log(N^2/0.05)/M) -- src/GraphicalModelLearning.jl:157."""
    return c * math.sqrt(math.log((num_spins ** 2) / 0.05) / num_samples)


# --------------------------------------------------------------------------------------
# nodal statistics (features)
# --------------------------------------------------------------------------------------
def nodal_stat_pairwise(samples: np.ndarray, u: int) -> np.ndarray:
    """K x N matrix stat[k,i] = s_u^k * s_i^k (i != u), stat[k,u] = s_u^k.

    src/GraphicalModelLearning.jl:162.  `u` is 0-based here.
    """
    spins = samples[:, 1:].astype(np.float64)
    stat = spins * spins[:, u:u + 1]
    stat[:, u] = spins[:, u]
    return stat


def combinations_sorted(items, order):
    """`permutations(items, order)` with asymmetric=false: strictly ascending tuples in
    lexicographic order -- src/models.jl:228-246."""
    return list(itertools.combinations(sorted(items), order))


def multirise_keys(num_spins: int, u: int, inter_order: int):
    """Keys of node u (1-based spin labels, as the reference builds them):
    (u,), then (u, j) ..., then (u, j<k) ... -- src/GraphicalModelLearning.jl:94-104."""
    neighbours = [i for i in range(1, num_spins + 1) if i != u]
    keys = []
    for p in range(1, inter_order + 1):
        if p == 1:
            keys.append((u,))
        else:
            keys.extend((u,) + perm for perm in combinations_sorted(neighbours, p - 1))
    return keys


def nodal_stat_multibody(samples: np.ndarray, keys) -> np.ndarray:
    """stat[k, key] = prod_{i in key} s_i^k -- src/GraphicalModelLearning.jl:106-108."""
    spins = samples[:, 1:].astype(np.float64)
    cols = [np.prod(spins[:, [i - 1 for i in key]], axis=1) for key in keys]
    return np.stack(cols, axis=1)


# --------------------------------------------------------------------------------------
# objectives: value, gradient, Hessian of the smooth part f_u
# --------------------------------------------------------------------------------------
def smooth_parts(form: str, x: np.ndarray, stat: np.ndarray, w: np.ndarray, hess: bool = True):
    """f, grad f, Hess f for one node.

    RISE    sum_k w_k exp(-t_k)              src/GraphicalModelLearning.jl:170
    logRISE log(sum_k w_k exp(-t_k))         src/GraphicalModelLearning.jl:279
    RPLE    sum_k w_k log(1+exp(-2 t_k))     src/GraphicalModelLearning.jl:317
    with t_k = sum_i x_i stat[k,i], w_k = samples[k,1]/num_samples.
    """
    t = stat @ x
    if form == "RISE":
        e = w * np.exp(-t)
        f = e.sum()
        g = -(stat.T @ e)
        H = (stat.T * e) @ stat if hess else None
    elif form == "logRISE":
        tmin = (-t).max()
        e = w * np.exp(-t - tmin)
        Z = e.sum()
        f = math.log(Z) + tmin
        p = e / Z
        g = -(stat.T @ p)
        H = (stat.T * p) @ stat - np.outer(g, g) if hess else None
    elif form == "RPLE":
        a = -2.0 * t
        f = (w * (np.maximum(a, 0.0) + np.log1p(np.exp(-np.abs(a))))).sum()
        sig = 0.5 * (1.0 - np.tanh(t))           # 1/(1+exp(2t))
        g = -2.0 * (stat.T @ (w * sig))
        H = 4.0 * ((stat.T * (w * sig * (1.0 - sig))) @ stat) if hess else None
    else:
        raise ValueError(form)
    return f, g, H


def smooth_value(form, x, stat, w):
    return smooth_parts(form, x, stat, w, hess=False)[0]


def full_objective(form, x, stat, w, lam, pen):
    """Objective including the L1 term on penalised coordinates (the value Ipopt minimises
    once z_j = |x_j|): src/GraphicalModelLearning.jl:169-172."""
    return smooth_value(form, x, stat, w) + lam * np.abs(x[pen]).sum()


# --------------------------------------------------------------------------------------
# solvers for one node
# --------------------------------------------------------------------------------------
def solve_node_exact(form, stat, w, lam, pen, x0=None, tol=1e-13, max_outer=200):
    """Exact L1 minimiser by proximal Newton + cyclic coordinate descent."""
    F = stat.shape[1]
    x = np.zeros(F) if x0 is None else x0.copy()
    iters = 0
    for outer in range(max_outer):
        iters += 1
        f, g, H = smooth_parts(form, x, stat, w)
        Fx = f + lam * np.abs(x[pen]).sum()
        # --- coordinate descent on the L1-regularised quadratic model
        d = np.zeros(F)
        Hd = np.zeros(F)
        diag = np.maximum(np.diag(H), 1e-300)
        for sweep in range(10000):
            maxchg = 0.0
            for j in range(F):
                gj = g[j] + Hd[j]
                cur = x[j] + d[j]
                v = cur - gj / diag[j]
                if pen[j]:
                    thr = lam / diag[j]
                    v = math.copysign(max(abs(v) - thr, 0.0), v)
                delta = v - cur
                if delta != 0.0:
                    d[j] += delta
                    Hd += H[:, j] * delta
                    maxchg = max(maxchg, abs(delta))
            if maxchg < 1e-16 + 1e-3 * tol:
                break
        step = np.abs(d).max()
        if step < tol:
            x = x + d
            break
        # --- Armijo line search on the composite objective
        delta_model = g @ d + lam * (np.abs((x + d)[pen]).sum() - np.abs(x[pen]).sum())
        alpha = 1.0
        while True:
            xn = x + alpha * d
            Fn = smooth_value(form, xn, stat, w) + lam * np.abs(xn[pen]).sum()
            if Fn <= Fx + 1e-4 * alpha * delta_model + 1e-13 * max(abs(Fx), 1.0) or alpha < 1e-10:
                break
            alpha *= 0.5
        x = xn
    return x, iters


def barrier_penalty(x, lam, mu):
    """phi, phi', phi'' of the slack-eliminated Ipopt barrier term (see module docstring)."""
    eps = mu / lam
    r = np.sqrt(eps * eps + x * x)
    z = eps + r
    # lambda*z - mu*log(z^2 - x^2), and z^2 - x^2 = 2*eps*z
    phi = lam * z - mu * np.log(2.0 * eps * z)
    dphi = lam * x / z
    d2phi = lam * (z - x * x / r) / (z * z)
    return phi, dphi, d2phi


def solve_node_barrier(form, stat, w, lam, pen, mu, x0, tol=1e-15, max_iter=200):
    """Minimiser of f(x) + sum_{pen} phi_mu(x_j) by damped Newton from x0."""
    x = x0.copy()
    idx = np.where(pen)[0]
    iters = 0

    def merit(xx):
        return smooth_value(form, xx, stat, w) + barrier_penalty(xx[idx], lam, mu)[0].sum()

    for it in range(max_iter):
        iters += 1
        f, g, H = smooth_parts(form, x, stat, w)
        phi, dphi, d2phi = barrier_penalty(x[idx], lam, mu)
        grad = g.copy()
        grad[idx] += dphi
        Hm = H.copy()
        Hm[idx, idx] += d2phi
        try:
            d = -np.linalg.solve(Hm, grad)
        except np.linalg.LinAlgError:
            d = -np.linalg.lstsq(Hm, grad, rcond=None)[0]
        m0 = f + phi.sum()
        slope = grad @ d
        alpha = 1.0
        while True:
            xn = x + alpha * d
            if merit(xn) <= m0 + 1e-4 * alpha * slope + 1e-13 * max(abs(m0), 1.0) or alpha < 1e-12:
                break
            alpha *= 0.5
        x = xn
        if np.abs(alpha * d).max() < tol:
            break
    return x, iters


def solve_node(form, stat, w, lam, pen, mode="exact", mu=1e-9):
    x, it = solve_node_exact(form, stat, w, lam, pen)
    if mode == "exact":
        return x, it
    if mode == "barrier":
        # warm start: exact solution; exact zeros are moved to the first-order barrier value so
        # Newton starts inside the basin of the smooth problem.
        _, g, _ = smooth_parts(form, x, stat, w, hess=False)
        x0 = x.copy()
        for j in np.where(pen)[0]:
            if x0[j] == 0.0:
                denom = max(lam * lam - g[j] * g[j], 1e-300)
                x0[j] = -2.0 * mu * g[j] / denom
        xb, it2 = solve_node_barrier(form, stat, w, lam, pen, mu, x0)
        return xb, it + it2
    raise ValueError(mode)


# --------------------------------------------------------------------------------------
# learn() restatements
# --------------------------------------------------------------------------------------
@dataclass
class LearnInfo:
    lam: float
    objective: np.ndarray      # per node, f_u(x) + lam*|x_pen|_1 at the returned point
    iterations: np.ndarray


def learn_pairwise(samples, form="RISE", regularizer=None, symmetrization=True,
                   mode="exact", mu=1e-9, nodes=None, return_info=False):
    """learn(samples, RISE/logRISE/RPLE(regularizer, symmetrization), NLP())
    -- src/GraphicalModelLearning.jl:154-189 / 263-298 / 301-336."""
    if regularizer is None:
        regularizer = {"RISE": 0.4, "logRISE": 0.8, "RPLE": 0.2}[form]   # :35,:49,:56
    samples = np.asarray(samples, dtype=np.float64)
    K, N, M = data_info(samples)
    lam = regularizer_lambda(regularizer, N, M)
    w = samples[:, 0] / M
    recon = np.zeros((N, N))
    node_list = range(N) if nodes is None else nodes
    objs = np.zeros(N)
    its = np.zeros(N, dtype=np.int64)
    for u in node_list:
        stat = nodal_stat_pairwise(samples, u)
        pen = np.ones(N, dtype=bool)
        pen[u] = False                                  # :171  (j != current_spin)
        x, it = solve_node(form, stat, w, lam, pen, mode, mu)
        recon[u, :] = x                                 # :181
        objs[u] = full_objective(form, x, stat, w, lam, pen)
        its[u] = it
    if symmetrization and nodes is None:
        recon = 0.5 * (recon + recon.T)                 # :184-186
    if return_info:
        return recon, LearnInfo(lam, objs, its)
    return recon


def learn_multibody(samples, regularizer=0.4, symmetrization=True, interaction_order=2,
                    mode="exact", mu=1e-9):
    """learn(samples, multiRISE(regularizer, symmetrization, interaction_order), NLP())
    -- src/GraphicalModelLearning.jl:83-152.  Returns dict {key tuple (1-based): value}."""
    samples = np.asarray(samples, dtype=np.float64)
    K, N, M = data_info(samples)
    lam = regularizer_lambda(regularizer, N, M)
    w = samples[:, 0] / M
    recon = {}
    for u in range(1, N + 1):
        keys = multirise_keys(N, u, interaction_order)
        stat = nodal_stat_multibody(samples, keys)
        pen = np.array([len(k) > 1 for k in keys])      # :118
        x, _ = solve_node("RISE", stat, w, lam, pen, mode, mu)
        for k, v in zip(keys, x):
            recon[k] = float(v)                          # :129-132
    if symmetrization:                                   # :135-149
        groups = {}
        for k, v in recon.items():
            groups.setdefault(tuple(sorted(k)), []).append(v)
        recon = {k: float(np.mean(v)) for k, v in groups.items()}
    return recon


def matrix_to_dict(m: np.ndarray):
    """convert(Dict, matrix): (i,) -> diag, (i,j) i != j -> m[i,j]; entries isapprox 0 dropped
    -- src/models.jl:157-182 (1-based keys)."""
    n = m.shape[0]
    out = {}
    for i in range(n):
        if m[i, i] != 0.0:
            out[(i + 1,)] = float(m[i, i])
    for i in range(n):
        for j in range(n):
            if i != j and m[i, j] != 0.0:
                out[(i + 1, j + 1)] = float(m[i, j])
    return out


# --------------------------------------------------------------------------------------
# exact sampler (input generator used by the reference's statistical tests)
# --------------------------------------------------------------------------------------
def sample_exact(terms: dict, num_spins: int, number_sample: int, rng: np.random.Generator):
    """Exact iid sampling by enumeration of all 2^N configurations -- src/sampling.jl:58-88.
    `terms` maps 1-based index tuples to weights; spin_i = 2*bit_i - 1, bit 0 = spin 1
    (src/sampling.jl:11-14).  Returns the histogram matrix [count, s_1..s_N] (Int64)."""
    ncfg = 1 << num_spins
    ids = np.arange(ncfg)
    spins = 2 * ((ids[:, None] >> np.arange(num_spins)[None, :]) & 1) - 1
    energy = np.zeros(ncfg)
    for term, wgt in terms.items():
        energy += wgt * np.prod(spins[:, [i - 1 for i in term]], axis=1)
    p = np.exp(energy - energy.max())
    p /= p.sum()
    counts = rng.multinomial(number_sample, p)
    keep = counts > 0
    return np.concatenate([counts[keep, None], spins[keep]], axis=1).astype(np.int64)


def matrix_to_terms(m: np.ndarray):
    """FactorGraph(matrix): diagonal -> 1-body, upper triangle -> 2-body -- src/models.jl:105-135."""
    n = m.shape[0]
    terms = {}
    for i in range(n):
        if m[i, i] != 0.0:
            terms[(i + 1,)] = float(m[i, i])
    for i in range(n):
        for j in range(i + 1, n):
            if m[i, j] != 0.0:
                terms[(i + 1, j + 1)] = float(m[i, j])
    return terms
