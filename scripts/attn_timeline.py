#!/usr/bin/env python
"""Timeline of the attention kernel's softmax / MMA hand-offs (b200_attn_fwd_prof): per key tile, SM-clock timestamps of CTA
(0,0,0) while the whole grid runs.  Prints medians of the intervals that make up the S -> softmax -> P -> PV dependency loop.
B200_ATTN_2CTA selects the kernel form (read once per process)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from apex_studio_b200 import _lib  # noqa: E402


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 75600
    steps = 400
    lib = _lib.load()
    q, k, v = (torch.randn(1, H, S, 128, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    o = torch.empty_like(q)
    prof = torch.zeros(steps, 32, dtype=torch.int64, device="cuda")
    st = lambda t: [t.stride(0), t.stride(1), t.stride(2)]
    for _ in range(2):
        rc = lib.b200_attn_fwd_prof(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), 1, H, S, S, 128, *st(q), *st(k), *st(v),
                                    *st(o), 128 ** -0.5, prof.data_ptr(), steps, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
    torch.cuda.synchronize()
    if os.environ.get("ATTN_TIMELINE_RAW"):
        raw = prof.cpu()[200:204]
        base = int(raw[0, 0])
        names = ["t0:S_seen", "t0:S_in_regs", "t0:max_done", "t0:P_stored", "t0:P_arrived", "t1:S_seen", "t1:S_in_regs", "t1:max_done",
                 "t1:P_stored", "t1:P_arrived", "mma:pre_PV0", "mma:PV0_issued", "mma:pre_PV1", "mma:PV1_issued", "mma:K_landed", "mma:S_free_seen",
                 "t0w0:P_arrived", "t0w1:P_arrived", "t0w2:P_arrived", "t0w3:P_arrived", "t1w0:P_arrived", "t1w1:P_arrived", "t1w2:P_arrived",
                 "t1w3:P_arrived", "mma:V_landed(0)", "mma:P0_seen", "mma:V_landed(1)", "mma:P1_seen"] + ["?"] * 4
        ev = sorted((int(raw[r, c]) - base, f"j={200 + r} {names[c]}") for r in range(4) for c in range(32) if int(raw[r, c]))
        for tt, nm in ev:
            print(f"{tt:7d}  {nm}", flush=True)
    t = prof.cpu().double()[100:380]          # steady state
    med = lambda x: float(x.median())
    out = {"pair": os.environ.get("B200_ATTN_2CTA", "default"), "heads": H, "S": S}
    out["step_period"] = med(t[1:, 11] - t[:-1, 11])
    for tile, b in ((0, 0), (1, 5)):
        out[f"t{tile}_ld_S"] = med(t[:, b + 1] - t[:, b + 0])
        out[f"t{tile}_rowmax"] = med(t[:, b + 2] - t[:, b + 1])
        out[f"t{tile}_exp_and_P_store_issue"] = med(t[:, b + 3] - t[:, b + 2])
        out[f"t{tile}_st_wait_arrive"] = med(t[:, b + 4] - t[:, b + 3])
        out[f"t{tile}_softmax_total"] = med(t[:, b + 4] - t[:, b + 0])
        out[f"t{tile}_wait_for_S"] = med(t[1:, b + 0] - t[:-1, b + 4])
    # MMA thread: when it sees P ready relative to the softmax warp's arrive, and how long it waits
    out["mma_P0_seen_after_arrive"] = med(t[:, 11] - t[:, 4])
    out["mma_P1_seen_after_arrive"] = med(t[:, 13] - t[:, 9])
    out["mma_wait_P0"] = med(t[:, 11] - t[:, 10])
    out["mma_wait_P1"] = med(t[:, 13] - t[:, 12])
    out["mma_PV0_to_S0next_issued"] = med(t[:, 12] - t[:, 11])
    # S0(j+1) issued (stamp 12 of step j) -> softmax 0 of step j+1 sees S ready: MMA execution + commit latency
    out["S0_issue_to_seen"] = med(t[1:, 0] - t[:-1, 12])
    out["S1_issue_to_seen"] = med(t[1:, 5] - t[:-1, 13])
    print(json.dumps({k_: (round(v_, 1) if isinstance(v_, float) else v_) for k_, v_ in out.items()}), flush=True)


if __name__ == "__main__":
    main()
