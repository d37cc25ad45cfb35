import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from apex_studio_b200 import ops
def t(M,N,K,epi):
    x=torch.randn(M,K,device="cuda").bfloat16(); w=(torch.randn(N,K,device="cuda")*0.02).bfloat16(); b=torch.randn(N,device="cuda").bfloat16()
    out=torch.zeros(M,N,device="cuda",dtype=torch.bfloat16)
    f=lambda: ops.linear(x,w,b,epilogue=epi,out=out,gate=b if epi==2 else None)
    for _ in range(3): f()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    it=30 if M>20000 else 200
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/it
    return round(2.0*M*N*K/ms/1e9,1)
print([t(75600,15360,5120,0), t(75600,5120,5120,2), t(75600,13824,5120,1), t(75600,5120,13824,2), t(4608,9216,3072,0), t(8192,3072,12288,2)])
''' % ROOT
for g in (4, 8, 16, 32):
    env = dict(os.environ, B200_LINEAR_GROUP_M=str(g))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    print("GROUP_M", g, "[wan qkv, wan out, wan ff1, wan ff2, flux single qkv, qwen ff2] TF/s:", r.stdout.strip() or r.stderr[-300:], flush=True)
