"""The C-ABI library loads and exports every symbol include/apex_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "apex_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("b200_attn_fwd", "b200_linear", "b200_layernorm_modulate", "b200_rmsnorm_rope", "b200_gate_residual",
              "b200_cfg_combine", "b200_version", "b200_strerror"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from apex_studio_b200 import _lib

    lib = _lib.load()
    for s in _declared_symbols():
        assert hasattr(lib, s), f"libapex_b200.so does not export {s}"
        assert s in _lib.SIGNATURES, f"ctypes signature table misses {s}"
    assert lib.b200_version() >= 100
    assert lib.b200_strerror(0) == b"ok"
    assert b"aligned" in lib.b200_strerror(-2)


def test_sass_contains_blackwell_instructions():
    """tcgen05.mma -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld -> LDTM must be present in the built library."""
    import shutil
    import subprocess

    from apex_studio_b200 import _lib

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass  # no legacy mma.sync path


def test_error_codes_map_to_reference_exceptions():
    from apex_studio_b200 import _lib

    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(ValueError):
        _lib.check(-2, "x")
    with pytest.raises(RuntimeError):
        _lib.check(-5, "x")
    _lib.check(0, "x")


def test_argument_validation_needs_no_gpu():
    """Null pointers / bad shapes are rejected by the ABI before any CUDA call."""
    from apex_studio_b200 import _lib

    lib = _lib.load()
    assert lib.b200_linear(None, None, None, None, None, 1, 1, 8, 8, 8, 8, 0, None) == -6
    assert lib.b200_attn_fwd(1, 1, 1, 16, 1, 1, 8, 8, 64, *([8] * 12), 1.0, None) == -1  # head_dim 64
    assert lib.b200_layernorm_modulate(16, 16, None, None, None, None, 4, 12, 16, 16, 0, 1e-6, None) == -2
    # round-2 entry points: the NORMW epilogue only through b200_linear_normw (it needs the row_sumsq output), which checks its
    # extra operands and the capacity of the partial-sum buffer; the q-norm attention needs the partial sums; the fused
    # conv + norm epilogue rejects channel counts that span two N tiles; the joint scatter checks its peer table
    assert lib.b200_linear(16, 16, None, 16, 16, 64, 64, 64, 64, 64, 64, 6, None) == -6
    assert lib.b200_linear_normw(16, 16, None, None, 16, 16, 1, 16, 64, 64, 64, 64, 64, 64, None) == -6       # no norm weight
    assert lib.b200_linear_normw(16, 16, None, 16, 16, 16, 1, 16, 64, 256, 64, 64, 64, 256, None) == -1       # capacity < N / 64
    assert lib.b200_attn_fwd_qnorm(16, 16, 16, 16, 1, 1, 8, 8, 128, *([8] * 12), 1.0, None, 1, 128, 1e-6, None) == -6
    assert lib.b200_conv3d_cl_norm_silu(16, 16, None, 16, 16, 2, 8, 8, 64, 384, 3, 3, 3, None) == -1          # 384 = two N tiles
    assert lib.b200_conv3d_cl_norm_silu(16, 16, None, None, 16, 2, 8, 8, 64, 96, 3, 3, 3, None) == -6         # no gamma
    assert lib.b200_attn_fwd_scatter_joint(16, 16, 16, 2, 64, 64, 128, 128, 256, 128, 256, 128, 256, None, 2, 32, 0, 128, 512,
                                           8, 0, 1.0, None) == -6


def test_ops_fail_loudly_without_cuda_tensors():
    from apex_studio_b200 import ops

    x = torch.zeros(4, 128, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="no CPU fallback"):
        ops.linear(x, torch.zeros(8, 128, dtype=torch.bfloat16))
    with pytest.raises(ValueError, match="no CPU fallback"):
        ops.layernorm_modulate(x)
    with pytest.raises(ValueError):
        ops.attention(torch.zeros(1, 1, 8, 128), torch.zeros(1, 1, 8, 128), torch.zeros(1, 1, 8, 128))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "apex-studio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|wan_dit|unipc|ref_import)\b", src, flags=re.M), f
