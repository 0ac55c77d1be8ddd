#!/usr/bin/env python
"""bench.py -- learn() with RISE on the BASELINE.json headline workload (C3):
N = 1000 sparse random-graph Ising (4-regular, J = +-0.4, h = 0), 1e7-sample histogram.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank/GPU)
    python bench.py --impl reference ...                     (CPU oracle port on the host cores)

A "step" is one full learn(): all node problems solved to the stated tolerance on the resident
histogram (+ the row all-gather and symmetrisation).  metric = node*sample evals/s
= N*K*(n_fg + 0.5*n_f)/t  (SURVEY 8d); ms_per_step is the learn() time.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def regular_graph(n: int, degree: int, seed: int):
    """Random `degree`-regular simple graph by repeated pairing (seeded, deterministic)."""
    rng = np.random.default_rng(seed)
    while True:
        stubs = np.repeat(np.arange(n), degree)
        rng.shuffle(stubs)
        a, b = stubs[0::2], stubs[1::2]
        if np.any(a == b):
            continue
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        if len(set(zip(lo.tolist(), hi.tolist()))) != len(lo):
            continue
        return lo, hi


def c3_model(n: int, seed: int = 1000, degree: int = 4, coupling: float = 0.4):
    lo, hi = regular_graph(n, degree, seed)
    rng = np.random.default_rng(seed + 1)
    sign = rng.choice([-1.0, 1.0], size=len(lo))
    rows = [[] for _ in range(n)]
    for i, j, s in zip(lo, hi, sign):
        rows[i].append((j, coupling * s))
        rows[j].append((i, coupling * s))
    row_ptr = np.zeros(n + 1, dtype=np.int32)
    col, val = [], []
    for i in range(n):
        rows[i].sort()
        row_ptr[i + 1] = row_ptr[i] + len(rows[i])
        col += [c for c, _ in rows[i]]
        val += [v for _, v in rows[i]]
    truth = np.zeros((n, n))
    for i, j, s in zip(lo, hi, sign):
        truth[i, j] = truth[j, i] = coupling * s
    return row_ptr, np.array(col, dtype=np.int32), np.array(val, dtype=np.float32), truth


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for name, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1409.8), d.get("hbm_gbs", 6544.3), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (reference = Julia + Ipopt cannot run here)
# ------------------------------------------------------------------------------------------------
def cpu_baseline(spins_sample: np.ndarray, counts_sample: np.ndarray, nodes, lam: float):
    """Times the C oracle (per-node second-order solve, dense Hessian every iteration -- what the
    reference asks Ipopt to do) on a bounded sample: `nodes` node problems x K_s histogram rows."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import c_oracle
    c_oracle.set_threads(os.cpu_count() or 1)      # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    t0 = time.time()
    _, info = c_oracle.learn_pairwise_packed(counts_sample, spins_sample, "RISE", lam, False, "exact", 1e-9,
                                             nodes=nodes, return_info=True)
    dt = time.time() - t0
    n_nodes = nodes[1] - nodes[0]
    ks = spins_sample.shape[1]
    evals = n_nodes * ks * (info["n_fgh"] / n_nodes + 0.5 * info["n_f"] / n_nodes)
    return {"value": evals / dt, "unit": "node*sample evals/s", "cores": c_oracle.num_threads(), "kind": "port",
            "seconds": dt, "passes_fgh_per_node": info["n_fgh"] / n_nodes,
            "sample": f"{n_nodes} of N={spins_sample.shape[0]} node problems x first {ks} histogram rows, "
                      f"exact-L1 prox-Newton (f, grad, dense Hessian per iteration), OpenMP over nodes"}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's per-node solve on the host cores.
    (Julia + JuMP + Ipopt are not in the image, /root/reference has no compilable sources: oracle port.)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.nspins
    ks = args.cpu_rows
    row_ptr, col, val, _ = c3_model(n)
    # CPU-side Gibbs sample of the bounded row count (same model, same protocol as the device sampler)
    rng = np.random.default_rng(7)
    s = rng.choice(np.array([-1, 1], dtype=np.int8), size=(n, ks))
    for _ in range(args.sweeps):
        for i in range(n):
            fld = np.zeros(ks, dtype=np.float32)
            for q in range(row_ptr[i], row_ptr[i + 1]):
                fld += val[q] * s[col[q]]
            s[i] = np.where(rng.random(ks) < 1.0 / (1.0 + np.exp(-2.0 * fld)), 1, -1)
    counts = np.ones(ks)
    lam = 0.4 * np.sqrt(np.log(n * n / 0.05) / args.nsamples)
    sys.path.insert(0, str(ROOT / "oracle"))
    import c_oracle
    c_oracle.set_threads(os.cpu_count() or 1)
    nodes = (0, min(n, c_oracle.num_threads() * args.cpu_nodes_per_core))
    times, last = [], None
    for step in range(args.warmup + args.steps):
        last = cpu_baseline(np.ascontiguousarray(s), counts, nodes, lam)
        if step >= args.warmup:
            times.append(last["seconds"])
    value = last["value"] * last["seconds"] / np.mean(times)
    line = {"impl": "reference", "metric": "node_sample_evals_per_s", "value": value, "unit": "node*sample evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3: learn() RISE(0.4,true), N={n} random 4-regular Ising J=+-0.4, M=K={int(args.nsamples):g} Gibbs samples "
                                   f"({args.sweeps} sweeps/chain) -- CPU arm: bounded sample of this workload, see cpu_baseline.sample",
                       "arithmetic": "float64 per-node prox-Newton (f, gradient, dense Hessian per iteration), OpenMP over nodes"},
            "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": value, "unit": "node*sample evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    line["cpu_baseline"]["value"] = value
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import gml_b200
    from gml_b200 import _lib
    from gml_b200.distributed import learn_sharded, shard_bounds, upload_replicated

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    lib = _lib.load()

    n, k = args.nspins, int(args.nsamples)
    row_ptr, col, val, truth = c3_model(n)
    # ---- synthetic histogram, generated once on the device (replicated: same seed on every rank)
    spins = torch.empty((n, k), dtype=torch.int8, device=dev)
    counts = torch.ones(k, dtype=torch.float64, device=dev)
    t0 = time.time()
    _lib.check(lib.gml_b200_sample_gibbs_device(local, n, row_ptr.ctypes.data, col.ctypes.data, val.ctypes.data, None,
                                                k, args.sweeps, 1000, ctypes.c_void_p(spins.data_ptr()), k, None))
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    sess = gml_b200.Session(local).attach_device(counts.data_ptr(), spins.data_ptr(), k, n, k)
    form = gml_b200.RISE(0.4, True)
    lam = gml_b200.regularizer_lambda(0.4, n, float(k))
    method = gml_b200.B200(solver=args.solver, tol=args.tol, device=local, profile=True, verbose=args.verbose, multilevel=args.multilevel, coarse_level=not args.no_coarse)
    b, e = shard_bounds(n, world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        full = learn_sharded(sess, form, method, symmetrize=True)
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), full, dict(method.last_stats)

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    times, stats, full = [], None, None
    for _ in range(args.steps):
        ms, full, stats = one_step()
        times.append(ms)
    clocks = sampler.stop() if rank == 0 else {}
    ms_step = float(np.mean(times))
    # evals: passes weighted by the fraction of the histogram they sweep (coarse continuation levels count 1/stride),
    # summed over the ranks' shards
    ev = torch.tensor([stats["evals"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ev)
    value = float(ev.item()) / (ms_step * 1e-3)
    learned = full.cpu().numpy()
    recon_err = float(np.abs(learned - np.diag(np.diag(learned)) - truth).max())
    nnz_row = (learned != 0.0).sum(axis=1)          # symmetrised support (union of the two directed estimates)

    # ---- roofline of the dominant kernel (per-launch CUDA-event times from the library, this rank)
    nn_local = e - b
    F = n + 1
    flops_per_launch = 2.0 * k * F * nn_local                    # algorithmic: 2*K*F per node per contraction
    nfull = max(1, stats["timed_full_passes"])          # launches that swept the whole histogram (finest level)
    ker = {"tc_energy_pair_kernel(full)": (stats["energy_fg_ms"], nfull),
           "tc_grad_kernel": (stats["grad_ms"], nfull),
           "tc_energy_pair_kernel(objective)": (stats["energy_f_ms"], stats["n_f_passes"])}
    dom = max(ker, key=lambda kk: ker[kk][0])
    peak_tf, peak_gbs, peak_src = load_peaks()
    roof = None
    traffic = None
    tfile = ROOT / "profiles" / "r1_traffic.json"
    if tfile.exists() and world == 1 and n == 1000 and k == 10_000_000:
        traffic = json.loads(tfile.read_text()).get("c3_1gpu", {}).get(dom.split("(")[0])      # from the committed ncu --set full capture
    if ker[dom][0] > 0:
        avg_ms = ker[dom][0] / max(1, ker[dom][1])
        ach = flops_per_launch / (avg_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": traffic, "avg_launch_ms": avg_ms, "launches": ker[dom][1], "peak_source": peak_src,
                "algorithmic_flops_per_launch": flops_per_launch,
                "kernel_ms_share": {kk: v[0] / stats["solve_ms"] for kk, v in ker.items()},
                "note": "launches over the whole shard only (compacted tail passes are not timed, so kernel_ms_share sums to < 1); "
                        "algorithmic flops = 2*K*F*nodes per contraction; the int8 limb split executes 3x (energy) / 2x (gradient) that "
                        "many int8 MACs on the coarse precision level, 4x / 3x on the fine one"}

    # ---- e2e: host (pinned) buffers in, host matrix out, through the same C ABI
    e2e = None
    if not args.skip_e2e:
        h_spins = torch.empty((n, k), dtype=torch.int8, pin_memory=True)
        h_spins.copy_(spins)
        h_counts = torch.ones(k, dtype=torch.float64, pin_memory=True)
        del spins
        sess.close()
        torch.cuda.empty_cache()
        np_spins, np_counts = h_spins.numpy(), h_counts.numpy()
        e2e_method = gml_b200.B200(solver=args.solver, tol=args.tol, device=local, multilevel=args.multilevel, coarse_level=not args.no_coarse)
        e_times = []
        for it in range(1 + args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            s2 = upload_replicated(gml_b200.Session(local), h_counts, h_spins)    # H2D (1/world per rank) + NVLink all-gather + validate + layout
            torch.cuda.synchronize(); t1 = time.perf_counter()
            out = learn_sharded(s2, form, e2e_method, symmetrize=True)        # solve + all-gather + symmetrise
            host = out.cpu()                                                  # D2H of the N x N result
            torch.cuda.synchronize()
            if args.verbose and rank == 0:
                print(f"[bench e2e] upload {t1 - t0:.3f} s, solve+gather+d2h {time.perf_counter() - t1:.3f} s", file=sys.stderr)
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            st2 = dict(e2e_method.last_stats)
            s2.close()
            if it > 0:
                e_times.append(float(dt.item()))
        ev2 = torch.tensor([st2["evals"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ev2)
        e_med = float(np.median(e_times))                 # median of the timed end-to-end calls
        e2e = {"value": float(ev2.item()) / e_med, "unit": "node*sample evals/s",
               "h2d_bytes_per_step": int(n * k + 8 * k), "d2h_bytes_per_step": int(8 * n * n * world),
               "learn_seconds": e_med, "steps": len(e_times), "all_seconds": e_times}

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:      # reported at N=1 only
        sys.path.insert(0, str(ROOT / "oracle"))
        import c_oracle
        c_oracle.set_threads(os.cpu_count() or 1)
        ks = args.cpu_rows
        src = h_spins if not args.skip_e2e else spins.cpu()
        sample = np.ascontiguousarray(src[:, :ks].numpy())
        nodes = (0, min(n, c_oracle.num_threads() * args.cpu_nodes_per_core))
        cpu = cpu_baseline(sample, np.ones(ks), nodes, lam)
        cpu["learn_seconds_extrapolated"] = (n * k * cpu["passes_fgh_per_node"]) / cpu["value"]

    if rank == 0:
        line = {"metric": "node_sample_evals_per_s", "value": value, "unit": "node*sample evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int8",
                "data": "synthetic",
                "config": {"workload": f"C3: learn() RISE(0.4,true), N={n} random 4-regular Ising J=+-0.4, M=K={k:g} Gibbs samples "
                                       f"({args.sweeps} sweeps/chain), node-sharded over {world} GPU(s), histogram replicated",
                           "arithmetic": "int8 tensor-core contractions with s32/s64 accumulation (exact), f32 per-sample epilogue, f64 solver state",
                           "tol": args.tol, "solver": args.solver, "l2_note": "inputs (10 GB int8 + 20-30 GB residual digits) exceed the 126 MB L2",
                           "lambda": lam, "sampler_seconds": gen_s},
                "learn_seconds": ms_step * 1e-3, "passes": {"fg": stats["n_fg_passes"], "f": stats["n_f_passes"], "iterations": stats["iterations"]},
                "max_abs_coupling_error_vs_truth": recon_err, "max_residual": stats["max_residual"],
                "support": {"mean_nnz_per_row": float(nnz_row.mean()), "max_nnz_per_row": int(nnz_row.max()), "true_degree": 4},
                "gpu_launches": int(stats["kernel_launches"]) * args.steps,
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nspins", type=int, default=1000)
    ap.add_argument("--nsamples", type=float, default=1e7)
    ap.add_argument("--sweeps", type=int, default=40)
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--solver", default="fista_tc")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-rows", type=int, default=2048)
    ap.add_argument("--cpu-nodes-per-core", type=int, default=1)
    ap.add_argument("--verbose", type=int, default=0)
    ap.add_argument("--multilevel", action="store_true")
    ap.add_argument("--no-coarse", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
