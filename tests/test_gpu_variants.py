"""Opt-in kernel forms stay correct: the selection environment variables are read once per process, so each form runs in
a subprocess.  Covered: the attention A/B partners -- the round-1 pipeline (B200_ATTN_PIPE=0) and the CTA-pair form
(B200_ATTN_2CTA=1) --, the 1-CTA GEMM for large M (B200_LINEAR_2CTA=0) and the 128x64 small-M GEMM form (B200_LINEAR_SMALLM=1).  Bars: attention rel-L2 <= 5e-3 vs fp32 math,
GEMM rel-L2 <= 4e-3 vs fp32 math of the same bf16 operands."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

GEMM_CODE = r"""
import json, sys, torch
sys.path.insert(0, %r)
from apex_studio_b200 import ops
res = {}
for (M, N, K, epi) in [(300, 520, 264, 0), (512, 3072, 3072, 0), (1024, 1024, 1024, 1), (2048, 768, 512, 2), (130, 256, 128, 0)]:
    torch.manual_seed(M)
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    acc = x.float() @ w.float().t() + b.float()
    if epi == 0:
        ref, out = acc, ops.linear(x, w, b)
    elif epi == 1:
        ref, out = torch.nn.functional.gelu(acc, approximate="tanh"), ops.linear(x, w, b, epilogue=ops.EPI_GELU_TANH)
    else:
        h = torch.randn(M, N, device="cuda").bfloat16()
        g = torch.randn(N, device="cuda").bfloat16()
        ref, out = h.float() + g.float() * acc, h.clone()
        ops.linear(x, w, b, epilogue=ops.EPI_GATE_RES, out=out, gate=g)
    res["%%dx%%dx%%d_%%d" %% (M, N, K, epi)] = ((out.float() - ref).norm() / ref.norm()).item()
print(json.dumps(res))
""" % ROOT


def _run(cmd, env):
    r = subprocess.run(cmd, env=dict(os.environ, **env), capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stderr[-800:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("env", [{"B200_ATTN_PIPE": "0"}, {"B200_ATTN_2CTA": "1"}, {"B200_ATTN_PIPE": "1", "B200_ATTN_2CTA": "0"},
                                 {"B200_ATTN_PERSIST": "0"}])
def test_attention_form_in_subprocess(env):
    res = _run([sys.executable, os.path.join(ROOT, "scripts", "attn_variant_ab.py")], env)
    assert res["nan"] is False
    errs = {k: v for k, v in res.items() if k.startswith("rel_")}
    assert len(errs) == 8 and all(v <= 5e-3 for v in errs.values()), errs


@pytest.mark.parametrize("env", [{"B200_LINEAR_2CTA": "0"}, {"B200_LINEAR_2CTA": "1"}, {"B200_LINEAR_SMALLM": "1"},
                                 {"B200_LINEAR_QUAD": "1"}, {"B200_LINEAR_QUAD": "0"}])
def test_linear_kernel_forms_in_subprocess(env):
    res = _run([sys.executable, "-c", GEMM_CODE], env)
    assert len(res) == 5 and all(v <= 4e-3 for v in res.values()), (env, res)
