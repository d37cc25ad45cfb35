"""Import the reference's OWN, unmodified hot-path modules from /root/reference (test infrastructure).

Used only by oracle/make_golden.py and by oracle-validation tests that skip when /root/reference is
absent (it does not exist on the GPU box).  Mechanism:
  * put oracle/ref_import (holding the `diffusers` stand-in) and /root/reference/apps/api on sys.path;
  * pre-seed bare `src.transformer`, `src.vae`, `src.scheduler`, `src.engine` packages so that their
    auto-discovery `__init__`s (transformer/__init__.py:81, vae/__init__.py:65), which import every model
    family and need diffusers/accelerate/ray, never run;
  * point APEX_HOME_DIR at a scratch dir because src.utils.defaults creates directories on import.
Nothing is written to /root/reference (PYTHONDONTWRITEBYTECODE).
"""
import importlib
import os
import sys
import tempfile
import types

# APEX_REFERENCE_API: an alternative location of the same unmodified files (baseline/_ref/apps/api on the GPU box)
REFERENCE_API = os.environ.get("APEX_REFERENCE_API") or "/root/reference/apps/api"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_API, "src"))


def setup():
    if not available():
        raise RuntimeError("/root/reference is not present (expected on the GPU box)")
    sys.dont_write_bytecode = True
    os.environ.setdefault("APEX_HOME_DIR", tempfile.mkdtemp(prefix="apex_home_"))
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, REFERENCE_API):
        if p not in sys.path:
            sys.path.insert(0, p)
    import src  # noqa: F401  (namespace/regular package at apps/api/src)

    for name in ("transformer", "vae", "scheduler", "engine", "transformer.wan", "transformer.wan.base",
                 "vae.wan", "engine.wan", "transformer.flux", "transformer.flux.base", "transformer.hunyuanvideo15",
                 "transformer.hunyuanvideo15.base", "transformer.qwenimage", "transformer.qwenimage.base", "vae.hunyuanvideo15", "transformer.flux2", "transformer.flux2.base"):
        full = "src." + name
        if full in sys.modules:
            continue
        m = types.ModuleType(full)
        m.__path__ = [os.path.join(REFERENCE_API, "src", *name.split("."))]
        m.__package__ = full
        sys.modules[full] = m
    # families other than Wan import the registry from the package itself (flux/base/model.py:56)
    tr = sys.modules["src.transformer"]
    if not hasattr(tr, "TRANSFORMERS_REGISTRY"):
        tr.TRANSFORMERS_REGISTRY = importlib.import_module("src.transformer.base").TRANSFORMERS_REGISTRY


def ref(module: str):
    """importlib.import_module after setup(); e.g. ref('src.transformer.wan.base.model')."""
    setup()
    return importlib.import_module(module)
