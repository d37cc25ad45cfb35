#!/usr/bin/env python
"""NVLink traffic of the fused norm+RoPE+exchange kernel, measurable by ncu: ONE process, two GPUs.  The kernel runs on cuda:0
as "sequence-parallel rank 0 of 2"; peer 0's receive plane is a local buffer, peer 1's lives on cuda:1 (peer access), so half
of every token's channels is stored over NVLink -- exactly what b200_rmsnorm_rope_scatter does inside the 2-rank job.

    ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum -k regex:rmsnorm_rope python scripts/nvlink_single_process.py

Prints the payload the kernel must push to the peer and the bandwidth derived from CUDA-event timing."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apex_studio_b200 import ops

assert torch.cuda.device_count() >= 2
torch.cuda.set_device(0)
S, heads, hd, P = 75600, 40, 128, 2
n_local, d, width = S // P, heads * hd, (heads // P) * hd
x = torch.randn(n_local, d, device="cuda:0").bfloat16()
w = torch.ones(d, device="cuda:0", dtype=torch.bfloat16)
rope = torch.randn(n_local, hd, device="cuda:0").bfloat16()
plane0 = torch.zeros(S, width, device="cuda:0", dtype=torch.bfloat16)
plane1 = torch.zeros(S, width, device="cuda:1", dtype=torch.bfloat16)
plane1.copy_(plane0)                      # first cross-device copy makes torch enable peer access 0 <-> 1
assert torch.cuda.can_device_access_peer(0, 1)
peers = (ctypes.c_void_p * P)(plane0.data_ptr(), plane1.data_ptr())
for _ in range(2):
    ops.rmsnorm_rope_scatter(x, w, rope, heads, 1e-6, peers, P, 0, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    ops.rmsnorm_rope_scatter(x, w, rope, heads, 1e-6, peers, P, 0, 0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
remote = n_local * width * 2
ok = torch.equal(plane1[:n_local].to("cuda:0"), ops.rmsnorm_rope_(x.clone(), w, rope, heads, 1e-6)[:, width:])
print(json.dumps({"kernel": "rmsnorm_rope_scatter (rank 0 of 2)", "rows": n_local, "remote_payload_bytes_per_launch": remote,
                  "local_payload_bytes_per_launch": remote, "ms_per_launch": ms, "nvlink_GBps_from_timing": remote / ms / 1e6,
                  "remote_result_correct": bool(ok)}))
