"""Multi-GPU plumbing: node subproblems are independent (the reference's serial loop at
src/GraphicalModelLearning.jl:161 has no cross-iteration dependency), so ranks own contiguous node
shards, the histogram is replicated, and the only exchange is ONE all-gather of the learned rows
before the symmetrisation (:184-186).  One process per GPU, torch.distributed for the collective."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_nodes: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [begin, end) of rank.  Small problems: balanced to within one node.  Once every rank gets
    at least 64 nodes the cuts are moved to multiples of 16: the kernels pad shards to 64-node tiles anyway, and a
    16-aligned first spin column lets the energy epilogue fetch its spins with vector loads (TMA alignment)."""
    base, extra = divmod(n_nodes, world)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    if base >= 64:
        up = lambda v: min(n_nodes, -(-v // 16) * 16)
        begin, end = (0 if rank == 0 else up(begin)), (n_nodes if rank == world - 1 else up(end))
    return begin, end


def gather_rows(local_rows: torch.Tensor, n_nodes: int, group=None) -> torch.Tensor:
    """All-gather the (n_local x N) row blocks into the full N x N row-major matrix on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_rows
    n_cols = local_rows.shape[1]
    cuts = [shard_bounds(n_nodes, world, r) for r in range(world)]
    max_rows = max(e - b for b, e in cuts)
    padded = torch.zeros((max_rows, n_cols), dtype=local_rows.dtype, device=local_rows.device)
    padded[: local_rows.shape[0]] = local_rows
    out = torch.empty((world * max_rows, n_cols), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    blocks = []
    for r, (b, e) in enumerate(cuts):
        blocks.append(out[r * max_rows: r * max_rows + (e - b)])
    return torch.cat(blocks, dim=0)


def learn_sharded(session, formulation, method, symmetrize: bool = True, group=None) -> torch.Tensor:
    """learn() for a pairwise formulation with the node loop sharded over the ranks of `group`.
    `session` holds the (replicated) histogram on this rank's GPU.  Returns the N x N device matrix
    (row-major, row u = node u) on every rank."""
    from . import _lib
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = session.N
    b, e = shard_bounds(n, world, rank)
    rows = torch.empty((e - b, n), dtype=torch.float64, device=f"cuda:{session.device}")
    stream = torch.cuda.current_stream().cuda_stream
    session.solve_pairwise_device(formulation, method, rows.data_ptr(), b, e, stream=stream)
    full = gather_rows(rows, n, group).contiguous()
    if symmetrize:
        import ctypes
        _lib.check(_lib.load().gml_b200_symmetrize_device(ctypes.c_void_p(full.data_ptr()), n,
                                                          ctypes.c_void_p(stream)))
    return full


def upload_replicated(session, counts, spins, group=None):
    """Make the full histogram resident on every rank with ONE pass over the host link per byte: each rank
    copies only its 1/world block of spin rows host->device, the blocks are exchanged with an NCCL
    all-gather over NVLink, and the library packs the result.  `counts` (f64 [K]) and `spins` (int8 [N x K],
    spin-major) are host arrays (pinned for full H2D speed) holding the same data on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    # torch tensors keep their pinned-memory attribute (numpy views of pinned memory are copied as pageable)
    t_spins = spins if isinstance(spins, torch.Tensor) else torch.from_numpy(spins)
    t_counts = counts if isinstance(counts, torch.Tensor) else torch.from_numpy(counts)
    if world == 1:
        return session.upload(t_counts.numpy(), t_spins.numpy())
    n, k = t_spins.shape
    # Split by SPIN ROWS: in the spin-major layout a block of rows is one contiguous host range (full-speed
    # H2D, no strided copy), and the all-gather along dim 0 reassembles [N x K] directly.
    rows = -(-n // world)
    dev = torch.device(f"cuda:{session.device}")
    lo, hi = min(n, rank * rows), min(n, (rank + 1) * rows)
    part = torch.ones((rows, k), dtype=torch.int8, device=dev)
    part[: hi - lo].copy_(t_spins[lo:hi], non_blocking=True)
    full = torch.empty((world * rows, k), dtype=torch.int8, device=dev)
    dist.all_gather_into_tensor(full, part, group=group)
    d_counts = t_counts.to(dev, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    session.attach_device(d_counts.data_ptr(), full.data_ptr(), k, n, k)
    return session


# ---- sample-sharded partition (secondary): histogram ROWS split over the ranks, every rank advances all
# node problems in lockstep; per pass the int64 gradient sums and the fp64 objective sums are all-reduced by the
# library's own NCCL communicator (csrc/comm.cu).  Each rank streams only K/world samples per pass (node shards stream
# the whole replicated histogram on every rank), so it is also the cheaper partition whenever K/world still fills the
# GPU -- measured at C3: profiles/r2_*.
def sample_slice(n_rows: int, world: int, rank: int, align: int = 256) -> Tuple[int, int]:
    """Histogram rows [begin, end) of rank: equal parts, cuts on multiples of `align` (CTA pairs sweep 2 x 128 rows)."""
    per = -(-n_rows // world)
    per = -(-per // align) * align
    return min(n_rows, rank * per), min(n_rows, (rank + 1) * per)


def upload_sample_sharded(session, counts, spins, group=None):
    """Make this rank's slice of the histogram resident (ONE H2D of K/world rows per rank, no collective), create the
    library communicator over `group` and make the weights global.  counts f64 [K], spins int8 [N x K] spin-major
    host arrays (pinned for full H2D speed) holding the same data on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    t_spins = spins if isinstance(spins, torch.Tensor) else torch.from_numpy(spins)
    t_counts = counts if isinstance(counts, torch.Tensor) else torch.from_numpy(counts)
    n, k = t_spins.shape
    b, e = sample_slice(k, world, rank)
    if e <= b:
        raise ValueError("sample-sharded upload: fewer histogram rows than ranks")
    # the library copies the column range with ONE strided cudaMemcpy2DAsync (N rows of e-b contiguous bytes, pitch K)
    # straight into the histogram's final rows; a torch copy_ of the non-contiguous view goes through a host-side gather
    # (measured: 3.1 s instead of 0.1 s for 5 GB)
    session.upload(t_counts[b:e].numpy(), t_spins[:, b:e].numpy())
    if world > 1:
        session.comm_init(group)
    return session


def learn_sample_sharded(session, formulation, method, symmetrize: bool = True) -> torch.Tensor:
    """learn() for a pairwise formulation on a sample-sharded session (upload_sample_sharded): every rank solves all
    nodes and ends with the same N x N device matrix (row-major); no gather."""
    import ctypes
    import dataclasses
    from . import _lib
    n = session.N
    sharded = dist.is_initialized() and dist.get_world_size() > 1
    m = dataclasses.replace(method, sample_sharded=sharded)
    rows = torch.empty((n, n), dtype=torch.float64, device=f"cuda:{session.device}")
    stream = torch.cuda.current_stream().cuda_stream
    try:
        session.solve_pairwise_device(formulation, m, rows.data_ptr(), 0, n, stream=stream)
    finally:
        method.last_stats = m.last_stats
    if symmetrize:
        _lib.check(_lib.load().gml_b200_symmetrize_device(ctypes.c_void_p(rows.data_ptr()), n, ctypes.c_void_p(stream)))
    return rows
