"""Sample-sharded mode: histogram rows split over 2 GPUs, NCCL all-reduce of the exact int64 gradient
sums per pass; must reproduce the single-GPU solution.  Needs two GPUs (skipped otherwise)."""
import os
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, counts, spins, expected, q):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import gml_b200
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    k = spins.shape[1]
    lo, hi = k * rank // world, k * (rank + 1) // world
    sess = gml_b200.Session(rank).upload(np.ascontiguousarray(counts[lo:hi]), np.ascontiguousarray(spins[:, lo:hi]))
    sess.comm_init()
    assert abs(sess.num_samples - counts.sum()) < 1e-6          # global M after the globalize step
    m = gml_b200.B200(solver="fista_tc", sample_sharded=True, device=rank, tol=1e-7)
    got = sess.solve_pairwise(gml_b200.RISE(0.4, False), m)
    q.put((rank, float(np.abs(got - expected).max()), m.last_stats["iterations"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sample_sharded_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
        sys.path.insert(0, str(p))
    import gml_b200
    from helpers import histogram_c1
    _, hist = histogram_c1(n=16, m_samples=300_000, seed=16)
    counts, spins = gml_b200.pack_histogram(hist)
    expected = gml_b200.learn(hist, gml_b200.RISE(0.4, False), gml_b200.B200(solver="fista_tc", tol=1e-7))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, spins, expected, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, iters in results:
        assert err <= 1e-6, (rank, err)      # both solves stop at a prox-gradient mapping of 1e-7


def test_single_process_multi_device_learn():
    """B200(devices=2): the one-shot C entry point shards the nodes over two GPUs from ONE process (what a single
    Julia process uses) and must return the single-GPU matrix."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
        sys.path.insert(0, str(p))
    import gml_b200
    from helpers import random_ising
    import gml_oracle as o
    n = 160                                   # >= 64 nodes per device: 16-aligned cut at 80
    rng = np.random.default_rng(4)
    spins = rng.choice(np.array([-1, 1], dtype=np.int8), size=(n, 60_000))
    hist = np.concatenate([np.ones((60_000, 1)), spins.T.astype(np.float64)], axis=1)
    one = gml_b200.learn(hist, gml_b200.RISE(0.4, True), gml_b200.B200(coarse_level=False, warm_start=False))
    m2 = gml_b200.B200(devices=2, coarse_level=False, warm_start=False)
    two = gml_b200.learn(hist, gml_b200.RISE(0.4, True), m2)
    assert np.abs(one - two).max() <= 1e-9
    assert np.array_equal(two, two.T)


def _worker_helpers(rank, world, port, counts, spins, expected, q):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import gml_b200
    from gml_b200.distributed import learn_sample_sharded, upload_sample_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    sess = upload_sample_sharded(gml_b200.Session(rank), counts, spins)
    sess.comm_init()                                                   # second call: the globalize step must be idempotent
    assert abs(sess.num_samples - counts.sum()) < 1e-6
    m = gml_b200.B200(solver="fista_tc", device=rank, tol=1e-7)
    got = learn_sample_sharded(sess, gml_b200.RISE(0.4, False), m, symmetrize=False).cpu().numpy()
    q.put((rank, got, m.last_stats["iterations"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sample_sharded_n200_vs_single_gpu_and_oracle():
    """The bench's multi-GPU path (distributed.upload_sample_sharded / learn_sample_sharded) at N = 200, K = 150 000:
    both ranks end with the same matrix, equal to the single-GPU solve and to the oracle on a node subset."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
        sys.path.insert(0, str(p))
    import c_oracle as c
    import gml_b200
    from test_gpu_headline_parity import gibbs_pairwise, sparse_model
    model = sparse_model(200, 7)
    spins = gibbs_pairwise(model, 150_000, 50, 11).cpu().numpy()
    counts = np.ones(spins.shape[1])
    sess = gml_b200.Session(0).upload(counts, spins)
    one, info = sess.solve_pairwise(gml_b200.RISE(0.4, False), gml_b200.B200(solver="fista_tc", tol=1e-7), return_info=True)
    sess.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_helpers, args=(r, 2, port, counts, spins, one, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(results[0][1], results[1][1])                  # exact integer gradient sums: identical decisions
    assert np.abs(results[0][1] - one).max() <= 1e-6
    ref = c.learn_pairwise_packed(counts, spins, "RISE", info["lambda"], False, nodes=(63, 65))
    assert np.abs(results[0][1][63:65] - ref[63:65]).max() <= 1e-5


def test_single_process_multi_device_sample_slices():
    """B200(devices=2) with enough rows per device: the one-shot C entry point splits the histogram ROWS over two GPUs
    from ONE process (host threads + the library's NCCL communicator) and symmetrises on the device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
        sys.path.insert(0, str(p))
    import gml_b200
    rng = np.random.default_rng(4)
    n, k = 100, 140_000
    spins = rng.choice(np.array([-1, 1], dtype=np.int8), size=(n, k))
    spins[1] = spins[0] * np.where(rng.random(k) < 0.7, 1, -1).astype(np.int8)
    counts = np.ones(k)
    one = gml_b200.learn_packed(counts, spins, gml_b200.RISE(0.4, True), gml_b200.B200(tol=1e-7))
    m2 = gml_b200.B200(devices=2, tol=1e-7)
    two = gml_b200.learn_packed(counts, spins, gml_b200.RISE(0.4, True), m2)
    assert np.abs(one - two).max() <= 1e-6
    assert np.array_equal(two, two.T)
