#!/usr/bin/env python
"""Raw timeline of consecutive WORK ITEMS of the persistent attention grid (b200_attn_fwd_prof, CTA 0): for short key
sequences (Wan cross-attention: 4 key tiles per item) the per-item overhead -- epilogue, Q / K hand-over, pipeline fill --
is what matters, not the steady-state period.  Usage: attn_timeline_items.py H Sq Sk [first_item n_items]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from apex_studio_b200 import _lib  # noqa: E402

NAMES = {0: "t0:S_seen", 1: "t0:S_in_regs", 2: "t0:max_done", 3: "t0:P_stored", 4: "t0:P_arrived", 5: "t1:S_seen", 6: "t1:S_in_regs",
         7: "t1:max_done", 8: "t1:P_stored", 9: "t1:P_arrived", 10: "mma:pre_PV0", 11: "mma:PV0_issued", 12: "mma:pre_PV1",
         13: "mma:PV1_issued", 14: "mma:K(j+2)_landed", 15: "mma:S_free_seen", 16: "t0:epilogue_begin", 17: "t0:epilogue_end",
         18: "t1:epilogue_begin", 19: "t1:epilogue_end", 20: "mma:Q_landed", 21: "mma:K0_landed", 22: "mma:S0(0)_issued",
         23: "t0:epi q0 box drained", 24: "t0:epi q0 O in regs", 25: "t0:epi q0 box written+fenced", 26: "t0:epi q0 store issued",
         27: "t0:epi q1 box drained", 28: "t0:epi q1 O in regs", 29: "t0:epi q1 box written+fenced", 30: "t0:epi q1 store issued"}


def main():
    H, Sq, Sk = (int(a) for a in sys.argv[1:4])
    first, n = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (20, 3)
    n_kv = (Sk + 127) // 128
    steps = (first + n + 2) * n_kv
    lib = _lib.load()
    q = torch.randn(1, H, Sq, 128, device="cuda", dtype=torch.bfloat16)
    k, v = (torch.randn(1, H, Sk, 128, device="cuda", dtype=torch.bfloat16) for _ in range(2))
    o = torch.empty_like(q)
    prof = torch.zeros(steps, 32, dtype=torch.int64, device="cuda")
    st = lambda t: [t.stride(0), t.stride(1), t.stride(2)]
    for _ in range(2):
        rc = lib.b200_attn_fwd_prof(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), 1, H, Sq, Sk, 128, *st(q), *st(k), *st(v),
                                    *st(o), 128 ** -0.5, prof.data_ptr(), steps, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
    torch.cuda.synchronize()
    raw = prof.cpu()
    r0, r1 = first * n_kv, (first + n) * n_kv
    base = int(raw[r0, 0])
    ev = sorted((int(raw[r, c]) - base, f"item {r // n_kv} j={r % n_kv} {NAMES.get(c, '?')}")
                for r in range(r0, r1) for c in range(32) if int(raw[r, c]))
    for tt, nm in ev:
        print(f"{tt:7d}  {nm}")
    if int(raw[r1, 0]):
        per_item = (int(raw[r1, 0]) - int(raw[r0, 0])) / n
        print(f"cycles per item (t0:S_seen of item {first} -> item {first + n}): {per_item:.0f}")
    print(f"persist={os.environ.get('B200_ATTN_PERSIST', '1')}")


if __name__ == "__main__":
    main()
