"""N>1 host logic on CPU: world_size-2 gloo run of the shard partition + row all-gather
(the only exchange of the node-sharded path), checked against the oracle's full matrix."""
import os
import pathlib
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _worker(rank, world, port, samples, expected, result_q):
    for p in (ROOT, ROOT / "oracle"):
        sys.path.insert(0, str(p))
    import c_oracle as c
    import gml_b200  # noqa: F401
    from gml_b200.distributed import gather_rows, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = samples.shape[1] - 1
    b, e = shard_bounds(n, world, rank)
    # the shard's rows (what gml_b200_solve_pairwise_device leaves on the GPU), here from the oracle
    rows = c.learn_pairwise(samples, "RISE", 0.3, False, nodes=(b, e))[b:e]
    full = gather_rows(torch.from_numpy(np.ascontiguousarray(rows)), n)
    sym = 0.5 * (full + full.T)
    result_q.put((rank, float(np.abs(sym.numpy() - expected).max()), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_all_nodes():
    sys.path.insert(0, str(ROOT))
    import gml_b200  # noqa: F401
    from gml_b200.distributed import shard_bounds
    for n in (1, 3, 7, 16, 125, 1000):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in cuts]
            if n // world >= 64:        # 16-aligned cuts
                assert all(b % 16 == 0 for b, _ in cuts) and max(sizes) - min(sizes) <= 32
            else:
                assert max(sizes) - min(sizes) <= 1


def test_gather_rows_handles_aligned_uneven_shards():
    """N=1000 over 8 ranks: 16-aligned cuts give 7 x 128 + 104 rows; the gather must pad to the LARGEST shard."""
    sys.path.insert(0, str(ROOT))
    import gml_b200  # noqa: F401
    from gml_b200.distributed import shard_bounds
    cuts = [shard_bounds(1000, 8, r) for r in range(8)]
    sizes = [e - b for b, e in cuts]
    assert sum(sizes) == 1000 and max(sizes) == 128 and all(b % 16 == 0 for b, _ in cuts)


def test_two_rank_gather_matches_oracle():
    for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
        sys.path.insert(0, str(p))
    import c_oracle as c
    from helpers import histogram_c1
    _, samples = histogram_c1(n=7, m_samples=4000, seed=11)     # 7 nodes over 2 ranks: ragged shards 4 + 3
    expected = c.learn_pairwise(samples, "RISE", 0.3, True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, samples, expected, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, shape in results:
        assert shape == (7, 7)
        assert err <= 1e-12


def test_sample_slices_cover_all_rows():
    sys.path.insert(0, str(ROOT))
    import gml_b200  # noqa: F401
    from gml_b200.distributed import sample_slice
    for k in (1, 255, 256, 30_001, 10_000_000):
        for world in (1, 2, 3, 8):
            cuts = [sample_slice(k, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == k
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert all(b % 256 == 0 or b == k for b, _ in cuts)


def _slice_worker(rank, world, port, counts, spins, x, result_q):
    for p in (ROOT, ROOT / "oracle"):
        sys.path.insert(0, str(p))
    import c_oracle as c
    import gml_b200  # noqa: F401
    from gml_b200.distributed import sample_slice
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sample_slice(spins.shape[1], world, rank)
    # what a rank of the sample-sharded mode holds: its rows, weights c_k / M_global (comm_globalize_histogram), and per
    # pass the un-normalised partial sums of f and of the gradient, combined with ONE all-reduce (csrc/comm.cu)
    m_local = torch.tensor([counts[b:e].sum()], dtype=torch.float64)
    m_global = m_local.clone()
    dist.all_reduce(m_global)
    f, g = c.eval_pairwise(counts[b:e], np.ascontiguousarray(spins[:, b:e]), "RISE", x)
    scale = float(m_local / m_global)
    part = torch.from_numpy(np.concatenate([f[:, None], g], axis=1) * scale)
    dist.all_reduce(part)
    result_q.put((rank, part.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_sharded_sums_match_oracle():
    """Host logic of the sample-sharded partition on CPU (gloo, world 2): slices + global weights + one all-reduce of the
    partial (f, grad f) sums reproduce the full-histogram evaluation of the oracle, identically on both ranks."""
    for p in (ROOT, ROOT / "oracle"):
        sys.path.insert(0, str(p))
    import c_oracle as c
    rng = np.random.default_rng(2)
    n, k = 9, 1000
    spins = rng.choice(np.array([-1, 1], dtype=np.int8), size=(n, k))
    counts = rng.integers(1, 4, size=k).astype(np.float64)
    x = rng.normal(size=(n, n + 1)) * 0.2
    f, g = c.eval_pairwise(counts, spins, "RISE", x)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_slice_worker, args=(r, 2, port, counts, spins, x, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(results[0][1], results[1][1])
    assert np.abs(results[0][1][:, 0] - f).max() <= 1e-13 and np.abs(results[0][1][:, 1:] - g).max() <= 1e-13
