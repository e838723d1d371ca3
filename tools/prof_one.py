"""One factorisation of the bench workload, for ncu: python tools/prof_one.py [m n nb]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dhqr_b200 as D
m, n, nb = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (32768, 4096, 0)
dev = torch.device("cuda:0")
A = D.colmajor_empty(m, n, dev); D.fill_uniform_(A, 0)
al = torch.zeros(n, dtype=torch.float64, device=dev)
D.householder_(A, al, nb)
torch.cuda.synchronize()
print("launches", D.default_handle(0).launch_count())
