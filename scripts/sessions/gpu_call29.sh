#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call29.log
: > $L
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flux.py tests/test_gpu_qwen.py tests/test_gpu_hy15.py tests/test_lora.py -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
for sm in 0 1; do
python - >> $L 2>&1 <<PY
import os, sys, torch
os.environ["B200_LINEAR_SMALLM"] = "$sm"
sys.path.insert(0, ".")
from apex_studio_b200 import ops
def t(M,N,K,epi):
    x=torch.randn(M,K,device="cuda").bfloat16(); w=(torch.randn(N,K,device="cuda")*0.02).bfloat16(); b=torch.randn(N,device="cuda").bfloat16()
    out=torch.zeros(M,N,device="cuda",dtype=torch.bfloat16)
    f=lambda: ops.linear(x,w,b,epilogue=epi,out=out,gate=b if epi==2 else None)
    ref=(x.float()@w.float().t()+b.float())
    f(); torch.cuda.synchronize()
    err=((out.float()-(ref if epi==0 else out.float())).norm()/ref.norm()).item() if epi==0 else 0.0
    for _ in range(3): f()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(300): f()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1)/300*1e3,1), round(err,5)
print("SMALLM=$sm us,err [txt qkv, txt out, txt ff1, txt ff2, hy ctx qkv 1985x6144x2048, 300x520x264]:", [t(512,9216,3072,0), t(512,3072,3072,0), t(512,12288,3072,0), t(512,3072,12288,0), t(1985,6144,2048,0), t(300,520,264,0)])
PY
done
tail -c 2000 $L
