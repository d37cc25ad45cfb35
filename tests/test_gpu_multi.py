"""Multi-GPU parity the driver can run: `pytest -m gpu` on a box with >= 2 GPUs self-launches `torch.distributed.run` on
scripts/gpu_mp_check.py / gpu_mp_check_mmdit.py and asserts the bit-identity the design claims:

  * sequence-parallel DiT forward (NCCL all-to-all AND the exchange fused into the kernels over NVLink peer memory) == the
    unsharded forward:                                   max |diff| == 0.0, and the fused path repeatable;
  * CFG x SP `moe_denoise` == the sequential loop:       max |diff| == 0.0, integer trace equal;
  * tile-parallel VAE decode == single-GPU tiled decode:  max |diff| == 0.0;
  * the same for the dual-stream families (HunyuanVideo-1.5, QwenImage).

On a 1-GPU box the tests skip (the round-end driver box is 1 GPU for the test tier; `gpurun --gpus 2` runs them)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _launch(script, nproc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", script)]
    env = dict(os.environ, NCCL_DEBUG=os.environ.get("NCCL_DEBUG", "WARN"))
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, r.stdout[-2000:]
    return json.loads(lines[-1])["per_rank"]


@pytest.mark.parametrize("nproc", [2, 4])
def test_wan_sharded_paths_are_bit_identical(nproc):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs, found {_ngpu()}")
    per_rank = _launch("gpu_mp_check.py", nproc, 29511 + nproc)
    assert len(per_rank) == nproc
    for res in per_rank:
        assert "p2p_error" not in res, res.get("p2p_error")
        assert res["sp_max_abs_diff"] == 0.0 and res["p2p_max_abs_diff"] == 0.0 and res["p2p_repeat_max_abs_diff"] == 0.0, res
        assert res["denoise_max_abs_diff"] == 0.0 and res["trace_equal"] is True, res
        assert res["vae_max_abs_diff"] == 0.0, res
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(per_rank, open(os.path.join(out, f"multi_gpu_parity_wan_n{nproc}.json"), "w"), indent=1)


@pytest.mark.parametrize("nproc", [2])
def test_dual_stream_sharded_paths_are_bit_identical(nproc):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs, found {_ngpu()}")
    per_rank = _launch("gpu_mp_check_mmdit.py", nproc, 29531 + nproc)
    for res in per_rank:
        diffs = {k: v for k, v in res.items() if k.endswith("max_abs_diff")}
        assert diffs and all(v == 0.0 for v in diffs.values()), res
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(per_rank, open(os.path.join(out, f"multi_gpu_parity_mmdit_n{nproc}.json"), "w"), indent=1)
