#!/usr/bin/env python
"""NVLink evidence for the fused compute+exchange kernels (run under torchrun, 2 or 4 ranks, one sequence-parallel group):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 scripts/nvlink_bytes.py

Runs ONE Wan-shaped layer's exchange (S = 75,600 tokens, 40 heads x 128) `reps` times through b200_rmsnorm_rope_scatter
(q, k, v) and b200_attn_fwd_scatter, and reads GPU 0's NVLink data counters (`nvidia-smi nvlink -gt d`) before and after.
Expected bytes leaving a rank per layer: (P-1)/P of its q|k|v shard (3 x S/P x 5120 x 2 B) + (P-1)/P of the attention rows it
computes for other ranks ((H/P) x 128 x 2 B x S x (P-1)/P).  Prints one JSON line from rank 0."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def nvlink_kib(gpu: int):
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True).stdout
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    if os.environ.get("NVLINK_RAW"):
        sys.stderr.write(out[:1500] + "\n")
    return tx, rx, len(re.findall(r"Data Tx:", out))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from apex_studio_b200 import ops
    from apex_studio_b200.parallel import ParallelContext
    from apex_studio_b200.wan.rope import wan_rope_table_bf16

    S, heads, hd, reps = 75600, 40, 128, 4
    par = ParallelContext.create(use_cfg=False, use_p2p=True)
    P = par.sp_size
    ex = par.peer_exchange(S, heads, hd, dev)
    n_local, d = S // P, heads * hd
    qkv = torch.randn(n_local, 3 * d, device=dev).bfloat16()
    w = torch.ones(d, device=dev, dtype=torch.bfloat16)
    rope = wan_rope_table_bf16(hd, (21, 45, 80), dev)[par.sp_rank * n_local:(par.sp_rank + 1) * n_local].contiguous()
    as4 = lambda t, n: t.view(1, n, -1, hd).transpose(1, 2)

    def layer():
        q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
        ops.rmsnorm_rope_scatter(q, w, rope, heads, 1e-6, ex.qkv_peers, ex.P, 0, ex.row0)
        ops.rmsnorm_rope_scatter(k, w, rope, heads, 1e-6, ex.qkv_peers, ex.P, ex.plane, ex.row0)
        ops.rmsnorm_rope_scatter(v, None, None, heads, 1e-6, ex.qkv_peers, ex.P, 2 * ex.plane, ex.row0, norm=False)
        ex.barrier(0)
        ops.attention_scatter(as4(ex.qkv[0], S), as4(ex.qkv[1], S), as4(ex.qkv[2], S), ex.o_peers, ex.P, ex.n_local, ex.head_off, d)
        ex.barrier(1)

    layer()
    torch.cuda.synchronize()
    dist.barrier()
    before = nvlink_kib(0) if rank == 0 else None
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        layer()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        after = nvlink_kib(0)
        qkv_out = 3 * n_local * d * 2 * (P - 1) / P
        o_out = (heads // P) * hd * 2 * S * (P - 1) / P
        exp = (qkv_out + o_out) * reps
        tx = (after[0] - before[0]) * 1024
        rx = (after[1] - before[1]) * 1024
        print(json.dumps({"world": world, "sp": P, "reps": reps, "links_reported": after[2], "gpu0_nvlink_tx_bytes": tx,
                          "gpu0_nvlink_rx_bytes": rx, "expected_payload_bytes_out": exp, "tx_over_expected": tx / exp if exp else None,
                          "rx_over_expected": rx / exp if exp else None, "ms_per_layer_exchange_plus_attention": e0.elapsed_time(e1) / reps,
                          "note": "counters include protocol overhead / acks; payload expected = (P-1)/P x (q|k|v shard + attention rows "
                                  "computed for other ranks)"}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
