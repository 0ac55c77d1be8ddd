"""Seeded synthetic inputs shared by the tests (BASELINE.json configs, reduced where noted)."""
import numpy as np

import gml_oracle as o

# test/common.jl:15-32
MODEL_A = np.array([[0.0, 0.1, 0.2], [0.1, 0.0, 0.3], [0.2, 0.3, 0.0]])
MODEL_B = np.array([[0.3, 0.1, 0.2], [0.1, 0.2, 0.3], [0.2, 0.3, 0.1]])
MODEL_C = np.array([[0.0, 0.1, 0.2, 0.3], [0.1, 0.0, 0.2, 0.3], [0.2, 0.2, 0.0, 0.3], [0.3, 0.3, 0.3, 0.0]])
MODELS = {"a": MODEL_A, "b": MODEL_B, "c": MODEL_C}
DEFAULT_C = {"RISE": 0.4, "logRISE": 0.8, "RPLE": 0.2}


def random_ising(n, seed, p_edge=0.25, jlo=0.2, jhi=0.6, hmax=0.2):
    """C1: Erdos-Renyi couplings +-U[jlo,jhi], fields U[-hmax,hmax]."""
    rng = np.random.default_rng(seed)
    m = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            if rng.random() < p_edge:
                m[i, j] = m[j, i] = rng.choice([-1.0, 1.0]) * rng.uniform(jlo, jhi)
        m[i, i] = rng.uniform(-hmax, hmax)
    return m


def histogram_c1(n=16, m_samples=100_000, seed=16):
    model = random_ising(n, seed)
    rng = np.random.default_rng(seed + 1)
    return model, o.sample_exact(o.matrix_to_terms(model), n, m_samples, rng)


def three_body_model(n, seed, n_triples=None):
    """C4-style: ring of pair couplings +-0.3 plus random triples +-0.4."""
    rng = np.random.default_rng(seed)
    terms = {}
    for i in range(n):
        terms[(i + 1, (i + 1) % n + 1) if i + 1 < (i + 1) % n + 1 else ((i + 1) % n + 1, i + 1)] = \
            float(rng.choice([-0.3, 0.3]))
    for _ in range(n_triples or n):
        t = tuple(sorted(int(x) + 1 for x in rng.choice(n, 3, replace=False)))
        terms[t] = float(rng.choice([-0.4, 0.4]))
    return terms
