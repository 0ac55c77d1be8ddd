"""ctypes front-end of oracle/libgml_oracle.so (the C twin of gml_oracle.py).

TEST INFRASTRUCTURE ONLY -- see the header of gml_oracle.c.  Importable only from tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

import ctypes
import pathlib
import subprocess

import numpy as np

from gml_oracle import data_info, multirise_keys, regularizer_lambda

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None
FORM_ID = {"RISE": 0, "logRISE": 1, "RPLE": 2}


def build(force: bool = False) -> pathlib.Path:
    so = _HERE / "libgml_oracle.so"
    src = _HERE / "gml_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libgml_oracle.so"], check=True,
                       capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        _LIB.gml_oracle_learn_pairwise.restype = ctypes.c_int
        _LIB.gml_oracle_learn_multibody.restype = ctypes.c_int
        _LIB.gml_oracle_learn_multibody_nodes.restype = ctypes.c_int
        _LIB.gml_oracle_eval_pairwise.restype = ctypes.c_int
        _LIB.gml_oracle_num_threads.restype = ctypes.c_int
    return _LIB


def pack(samples: np.ndarray):
    """[count, s_1..s_N] histogram -> (counts f64[K], spins int8 spin-major [N x K])."""
    samples = np.asarray(samples)
    counts = np.ascontiguousarray(samples[:, 0], dtype=np.float64)
    spins = np.ascontiguousarray(samples[:, 1:].T.astype(np.int8))
    return counts, spins


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def learn_pairwise(samples, form="RISE", regularizer=None, symmetrization=True, mode="exact",
                   mu=1e-9, nodes=None, return_info=False):
    if regularizer is None:
        regularizer = {"RISE": 0.4, "logRISE": 0.8, "RPLE": 0.2}[form]
    K, N, M = data_info(np.asarray(samples))
    lam = regularizer_lambda(regularizer, N, float(M))
    counts, spins = pack(samples)
    return learn_pairwise_packed(counts, spins, form, lam, symmetrization, mode, mu, nodes,
                                 return_info)


def learn_pairwise_packed(counts, spins, form, lam, symmetrization=True, mode="exact", mu=1e-9,
                          nodes=None, return_info=False):
    N, K = spins.shape
    out = np.zeros((N, N))
    obj = np.zeros(N)
    stats = np.zeros(2, dtype=np.int64)
    nb, ne = (0, N) if nodes is None else nodes
    lib().gml_oracle_learn_pairwise(
        _p(counts, ctypes.c_double), _p(spins, ctypes.c_int8), ctypes.c_int64(K), ctypes.c_int(N),
        ctypes.c_int64(spins.strides[0]), ctypes.c_int(FORM_ID[form]), ctypes.c_double(lam),
        ctypes.c_int(int(symmetrization)), ctypes.c_int(0 if mode == "exact" else 1),
        ctypes.c_double(mu), ctypes.c_int(nb), ctypes.c_int(ne), _p(out, ctypes.c_double),
        _p(obj, ctypes.c_double), _p(stats, ctypes.c_int64))
    if return_info:
        return out, {"lam": lam, "objective": obj, "n_fgh": int(stats[0]), "n_f": int(stats[1])}
    return out


def eval_pairwise(counts, spins, form, x, nodes=None, want_grad=True):
    """f_u and grad f_u (float64, no solve) at the rows of x for the listed nodes (default: all).
    counts f64[K], spins int8 [N x K] spin-major; x: len(nodes) x (N+1) in the C ABI's feature order
    (couplings to spins 0..N-1, self entry ignored, then the field).  Returns (f, g)."""
    N, K = spins.shape
    nodes = np.arange(N, dtype=np.int32) if nodes is None else np.ascontiguousarray(nodes, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert x.shape == (len(nodes), N + 1) and spins.dtype == np.int8 and spins.strides[1] == 1
    counts = np.ascontiguousarray(counts, dtype=np.float64)
    f = np.zeros(len(nodes))
    g = np.zeros_like(x) if want_grad else None
    lib().gml_oracle_eval_pairwise(
        _p(counts, ctypes.c_double), _p(spins, ctypes.c_int8), ctypes.c_int64(K), ctypes.c_int(N),
        ctypes.c_int64(spins.strides[0]), ctypes.c_int(FORM_ID[form]), _p(nodes, ctypes.c_int32),
        ctypes.c_int(len(nodes)), _p(x, ctypes.c_double), _p(f, ctypes.c_double),
        _p(g, ctypes.c_double) if want_grad else None)
    return f, g


def learn_multibody(samples, regularizer=0.4, symmetrization=True, interaction_order=2,
                    mode="exact", mu=1e-9, nodes=None):
    """nodes=(begin, end): solve only those node problems and return the un-symmetrised {(u, ...): value} entries of
    these nodes (symmetrisation needs all nodes)."""
    K, N, M = data_info(np.asarray(samples))
    lam = regularizer_lambda(regularizer, N, float(M))
    counts, spins = pack(samples)
    return learn_multibody_packed(counts, spins, lam, symmetrization, interaction_order, mode, mu, nodes)


def learn_multibody_packed(counts, spins, lam, symmetrization=True, interaction_order=2, mode="exact", mu=1e-9,
                           nodes=None):
    N, K = spins.shape
    keys = [multirise_keys(N, u, interaction_order) for u in range(1, N + 1)]
    n_keys = len(keys[0])
    kidx = -np.ones((N, n_keys, interaction_order), dtype=np.int32)
    klen = np.zeros((N, n_keys), dtype=np.int32)
    for u in range(N):
        for f, key in enumerate(keys[u]):
            klen[u, f] = len(key)
            kidx[u, f, :len(key)] = np.array(key) - 1
    vals = np.zeros((N, n_keys))
    obj = np.zeros(N)
    nb, ne = (0, N) if nodes is None else nodes
    lib().gml_oracle_learn_multibody_nodes(
        _p(counts, ctypes.c_double), _p(spins, ctypes.c_int8), ctypes.c_int64(K), ctypes.c_int(N),
        ctypes.c_int64(spins.strides[0]), ctypes.c_int(interaction_order), ctypes.c_int(n_keys),
        _p(kidx, ctypes.c_int32), _p(klen, ctypes.c_int32), ctypes.c_double(lam),
        ctypes.c_int(0 if mode == "exact" else 1), ctypes.c_double(mu), ctypes.c_int(nb), ctypes.c_int(ne),
        _p(vals, ctypes.c_double), _p(obj, ctypes.c_double))
    recon = {}
    for u in range(nb, ne):
        for f, key in enumerate(keys[u]):
            recon[key] = float(vals[u, f])
    if symmetrization and nodes is None:
        groups = {}
        for k, v in recon.items():
            groups.setdefault(tuple(sorted(k)), []).append(v)
        recon = {k: float(np.mean(v)) for k, v in groups.items()}
    return recon


def set_threads(n: int) -> None:
    lib().gml_oracle_set_threads(ctypes.c_int(int(n)))


def num_threads() -> int:
    return int(lib().gml_oracle_num_threads())
