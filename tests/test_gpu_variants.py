"""Opt-in kernel forms stay correct: the selection environment variables are read once per process, so each form runs in
a subprocess.  Covered: the attention A/B partners -- the round-1 pipeline (B200_ATTN_PIPE=0) and the CTA-pair form
(B200_ATTN_2CTA=1) --, the 1-CTA GEMM for large M (B200_LINEAR_2CTA=0) and the 128x64 small-M GEMM form (B200_LINEAR_SMALLM=1).  Bars: attention rel-L2 <= 5e-3 vs fp32 math,
GEMM rel-L2 <= 4e-3 vs fp32 math of the same bf16 operands; the conv forms (B200_CONV_TB / _KC / _PAD64) <= 4e-3 vs the fp32 oracle conv."""
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


CONV_CODE = r"""
import json, sys, torch
sys.path.insert(0, %r)
sys.path.insert(0, %r + "/oracle")
import wan_vae
from apex_studio_b200.vae.wan import conv3d_cl, AutoencoderKLWan
res = {}
for (T, H, W, cin, cout, taps) in [(5, 18, 16, 96, 96, (3, 3, 3)), (7, 12, 28, 64, 96, (3, 3, 3)), (6, 9, 5, 128, 192, (3, 3, 3)),
                                   (6, 6, 10, 64, 128, (3, 1, 1)), (9, 20, 40, 96, 16, (3, 3, 3)), (3, 8, 8, 32, 32, (3, 3, 3))]:
    g = torch.Generator().manual_seed(T * 100 + H + W + cin)
    x = torch.randn(1, cin, T, H, W, generator=g).bfloat16()
    w = (torch.randn(cout, cin, *taps, generator=g) * (cin * taps[0] * taps[1] * taps[2]) ** -0.5).bfloat16()
    b = (torch.randn(cout, generator=g) * 0.1).bfloat16()
    r = torch.randn(1, cout, T, H, W, generator=g).bfloat16()
    ref = wan_vae.causal_conv3d(x.float(), {"c.weight": w.float(), "c.bias": b.float()}, "c") + r.float()
    cl = lambda t: t[0].permute(1, 2, 3, 0).contiguous().cuda()
    out = conv3d_cl(cl(x), AutoencoderKLWan._tap_major(w.float()).cuda().bfloat16(), b.cuda(), taps, cout, residual=cl(r))
    got = out.permute(3, 0, 1, 2).unsqueeze(0).float().cpu()
    res["%%dx%%dx%%d_%%d_%%d_%%d" %% (T, H, W, cin, cout, taps[1])] = ((got - ref).norm() / ref.norm()).item()
print(json.dumps(res))
""" % (ROOT, ROOT)


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


@pytest.mark.parametrize("env", [{"B200_CONV_TB": "4"}, {"B200_CONV_TB": "2"}, {"B200_CONV_KC": "1"}, {"B200_CONV_PAD64": "1"}, {}])
def test_conv_kernel_forms_in_subprocess(env):
    """The opt-in conv forms: temporal blocking (conv3d_tb_kernel, TB = 4 / 2: ragged last frame block, causal skip of the
    left frames, time_conv taps, N = 16 / 96 / 192), one 32-channel chunk per stage, zero-filled 64-channel chunks."""
    res = _run([sys.executable, "-c", CONV_CODE], env)
    assert len(res) == 6 and all(v <= 4e-3 for v in res.values()), (env, res)
