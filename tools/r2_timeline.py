"""Look-ahead timeline of the bench workload on N GPUs (torchrun): per rank, the times at which panel k was usable and bulk
update k finished (option la_trace), printed by rank 0 for ranks 0 and N-1; plus warm solve timings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch, torch.distributed as dist
import dhqr_b200 as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
h = D.init_distributed(local)
m, n = 32768, 4096
b = D.splits(world, n); c0, nl = b[rank], b[rank + 1] - b[rank]
Al = D.colmajor_empty(m, nl, dev); al = torch.zeros(n, dtype=torch.float64, device=dev)
Ad = D.ColumnBlockMatrix(Al, n, c0, h)
K = n // 128
for rep in range(3):
    D.fill_uniform_(Al, 0, 0, c0, h); torch.cuda.synchronize(); dist.barrier(); h.set_option("la_trace", 1 if rep == 2 else 0)
    D.householder_(Ad, al, 0, h); torch.cuda.synchronize()
buf = torch.zeros(3 * K, dtype=torch.float64, device=dev)
D._lib.call("dhqr_debug_copy_f64", h.raw, b"la_times", C.c_void_p(buf.data_ptr()), 3 * K, None)
h.set_option("la_trace", 0)
t = buf.cpu().numpy().reshape(K, 3)
allt = [None] * world
dist.all_gather_object(allt, t)
if rank == 0:
    for r in (0, world - 1):
        tt = allt[r]
        print(f"rank {r}: total {tt[:, [0, 2]].max():.2f} ms; panel-ready steps (ms):", np.round(np.diff(np.concatenate([[0], tt[:, 0]])), 2).tolist())
        print(f"rank {r}: bulk-done steps (ms):", np.round(np.diff(np.concatenate([[0], tt[:, 2]])), 2).tolist(), flush=True)
D.shutdown_distributed(); dist.destroy_process_group()
