"""One warm H \\ b on the bench workload, for ncu (the Q'b sweep: k_pack, k_gram128, k_wreduce4, k_tinv, k_qt_dot, k_qt_axpy,
then k_backsolve_wave):  python tools/prof_solve.py [m n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dhqr_b200 as D
m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32768, 4096)
dev = torch.device("cuda:0")
A = D.colmajor_empty(m, n, dev); D.fill_uniform_(A, 0)
H = D.qr_(A)
b = torch.rand(m, dtype=torch.float64, device=dev)
x = D.ldiv(H, b)
torch.cuda.synchronize()
print("launches", D.default_handle(0).launch_count())
