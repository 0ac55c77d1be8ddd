"""Developer probe (2+ GPUs, torchrun): where does the replicated upload spend its time?"""
import os, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch, torch.distributed as dist
import gml_b200
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = torch.device(f"cuda:{local}")
n, k = 1000, int(float(sys.argv[1]))
h_spins = torch.ones((n, k), dtype=torch.int8, pin_memory=True)
h_counts = torch.ones(k, dtype=torch.float64, pin_memory=True)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    dist.barrier(); t0 = T()
    ks = -(-k // world); lo, hi = min(k, rank * ks), min(k, (rank + 1) * ks)
    part = torch.ones((n, ks), dtype=torch.int8, device=dev)
    part[:, : hi - lo].copy_(h_spins[:, lo:hi], non_blocking=True)
    cpart = torch.ones(ks, dtype=torch.float64, device=dev); cpart[: hi - lo].copy_(h_counts[lo:hi], non_blocking=True)
    t1 = T()
    allp = torch.empty((world, n, ks), dtype=torch.int8, device=dev); allc = torch.empty((world, ks), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allp, part); dist.all_gather_into_tensor(allc, cpart)
    t2 = T()
    full = allp.permute(1, 0, 2).reshape(n, world * ks)
    t3 = T()
    sess = gml_b200.Session(local)
    t4 = T()
    sess.attach_device(allc.data_ptr(), full.data_ptr(), k, n, world * ks)
    t5 = T()
    sess.close()
    t6 = T()
    if rank == 0:
        print(f"it {it}: h2d {t1-t0:.3f} allgather {t2-t1:.3f} permute {t3-t2:.3f} create {t4-t3:.3f} attach {t5-t4:.3f} close {t6-t5:.3f}", flush=True)
dist.destroy_process_group()
