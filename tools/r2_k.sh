#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "unblocked or golden or ragged or zero_pivot or host_buffer or smoke" --timeout 200 --timeout-method=thread 2>&1 | tail -4
timeout 120 python - <<'PY'
import sys; sys.path.insert(0,'.')
import torch, numpy as np, dhqr_b200 as D
dev=torch.device('cuda:0'); h=D.default_handle(0)
m,n=8192,1024
A=D.colmajor_empty(m,n,dev); al=torch.zeros(n,dtype=torch.float64,device=dev)
def t(opt):
    for k,v in opt.items(): h.set_option(k,v)
    ts=[]
    for _ in range(6):
        D.fill_uniform_(A,0); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); D.householder_(A,al,1); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
print('wave  :', t({'unblocked_wave':1}))
ref=A.clone(); aref=al.clone()
print('fused :', t({'unblocked_wave':0,'fuse_house':1}))
print('max|H_wave-H_fused|', float((ref-A).abs().max()), float((aref-al).abs().max()))
print('2-launch:', t({'unblocked_wave':0,'fuse_house':0}))
h.set_option('unblocked_wave',1); h.set_option('fuse_house',1)
PY
timeout 300 python bench.py --config 2 --no-cpu > gpurun_out/k_bench_c2.json 2> gpurun_out/k_bench_c2.err; echo "bench c2 rc=$?"; cut -c1-260 gpurun_out/k_bench_c2.json; tail -2 gpurun_out/k_bench_c2.err
