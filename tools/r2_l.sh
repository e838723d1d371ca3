#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "block_reflector or determinism or oracle or unblocked or lookahead" --timeout 200 --timeout-method=thread 2>&1 | tail -4
timeout 200 python - <<'PY'
import sys; sys.path.insert(0,'.')
import torch, numpy as np, dhqr_b200 as D
dev=torch.device('cuda:0'); h=D.default_handle(0)
m,n=32768,4096
A=D.colmajor_empty(m,n,dev); al=torch.zeros(n,dtype=torch.float64,device=dev)
def t(opt, nb=0, mm=None):
    for k,v in opt.items(): h.set_option(k,v)
    ts=[]
    for _ in range(6):
        D.fill_uniform_(A,0); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); D.householder_(A,al,nb); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
print('cvy_q=1 lookahead:', t({'cvy_q':1}))
print('cvy_q=0 lookahead:', t({'cvy_q':0}))
print('cvy_q=1 persist=2:', t({'cvy_q':1,'cvy_persist':2}))
h.set_option('cvy_persist',1)
print('cvy_q=1 serial:', t({'cvy_q':1,'lookahead':0}))
print('cvy_q=0 serial:', t({'cvy_q':0,'lookahead':0}))
h.set_option('lookahead',1); h.set_option('cvy_q',1)
h.set_option("profile", 1); D.fill_uniform_(A, 0); torch.cuda.synchronize(); h.profile_reset()
D.householder_(A, al, 0); torch.cuda.synchronize(); p = h.profile(); h.set_option("profile", 0)
print({k:(round(v['ms'],2), round(v['work']/v['ms']/1e9,1)) for k,v in p.items() if k.startswith('k_gemm')})
B=D.colmajor_empty(8192,1024,dev); a2=torch.zeros(1024,dtype=torch.float64,device=dev)
ts=[]
for _ in range(6):
    D.fill_uniform_(B,0); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); D.householder_(B,a2,1); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print('config 2 wave:', min(ts[1:]))
PY
