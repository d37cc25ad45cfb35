"""FunctionRegister / attention_register surface == apps/api/src/register/__init__.py:8-143 semantics."""
import pytest
import torch

from apex_studio_b200.register import FunctionRegister


def test_register_call_default_and_errors():
    reg = FunctionRegister()

    @reg("double")
    def double(x):
        return 2 * x

    @reg
    def triple(x):
        return 3 * x

    assert reg.call(3, key="double") == 6 and reg.call(3, key="triple") == 9
    reg.set_default("double")
    assert reg.get_default() == "double" and reg.call(5) == 10
    with pytest.raises(KeyError, match="already registered"):
        reg("double")(lambda x: x)
    reg("double", overwrite=True)(lambda x: x + 1)
    assert reg.call(1, key="double") == 2
    with pytest.raises(KeyError, match="not found"):
        reg.get("nope")
    reg("gone", available=False)(lambda: 1)
    assert not reg.is_available("gone") and "gone" in reg.all() and "gone" not in reg.all_available()
    with pytest.raises(RuntimeError, match="Function 'gone' is not available."):
        reg.call(key="gone")
    reg.set_availability("gone", True)
    assert reg.call(key="gone") == 1
    assert len(reg) == 3 and set(iter(reg)) == {"double", "triple", "gone"} and reg["triple"](2) == 6


def test_matches_reference_register_when_available():
    """Differential test against the reference's own class (container only)."""
    from ref_import import bootstrap

    if not bootstrap.available():
        pytest.skip("/root/reference not present")
    ref_cls = bootstrap.ref("src.register").FunctionRegister
    for cls in (ref_cls, FunctionRegister):
        r = cls()
        r("a")(lambda: "a")
        r("b", available=False)(lambda: "b")
        r.set_default("a")
        assert r.call() == "a"
        assert sorted(r.all()) == ["a", "b"] and sorted(r.all_available()) == ["a"]
        with pytest.raises(RuntimeError) as e:
            r.call(key="b")
        assert str(e.value) == "Function 'b' is not available."
        with pytest.raises(KeyError):
            r("a")(lambda: 0)


def test_b200_backend_registered_and_strict():
    from apex_studio_b200.attention import attention_register, b200_attention

    assert "b200" in attention_register.all()
    if not torch.cuda.is_available():
        assert not attention_register.is_available("b200")
        with pytest.raises(RuntimeError, match="Function 'b200' is not available."):
            attention_register.call(None, None, None, key="b200")
    q = torch.zeros(1, 1, 8, 128, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="mask"):
        b200_attention(q, q, q, attn_mask=torch.ones(8, 8))
    with pytest.raises(ValueError, match="causal"):
        b200_attention(q, q, q, is_causal=True)
    with pytest.raises(ValueError, match="dropout"):
        b200_attention(q, q, q, dropout_p=0.1)
