#!/usr/bin/env python
"""Builds baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun): the UNMODIFIED reference files of the Wan hot path --
exactly the import closure of `src.transformer.wan.base.model`, `src.attention.functions` and `src.vae.wan.model` under
/root/reference/apps/api -- plus the `diffusers` stand-in (a copy of oracle/ref_import; `diffusers` is an un-vendored dependency
of the reference).  Run in the build container (the reference does not exist on the GPU box):

    python baseline/make_ref.py

Used by baseline/ref_gpu_block.py (bench.py's `reference_gpu` block): the reference's own WanTransformerBlock on the B200."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/apps/api"
DST = os.path.join(ROOT, "baseline", "_ref")

PROBE = r"""
import sys, os
sys.path.insert(0, %r)
from ref_import import bootstrap
bootstrap.setup()
import importlib
for m in ("src.transformer.wan.base.model", "src.attention.functions", "src.vae.wan.model"):
    importlib.import_module(m)
for name, mod in sorted(sys.modules.items()):
    f = getattr(mod, "__file__", None)
    if f and f.startswith(%r):
        print(f)
""" % (os.path.join(ROOT, "oracle"), REF)


def main():
    if not os.path.isdir(REF):
        print("no /root/reference here: nothing to do (baseline/_ref is built in the container that has it)")
        return 0
    out = subprocess.run([sys.executable, "-c", PROBE], capture_output=True, text=True, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    if out.returncode != 0:
        sys.stderr.write(out.stderr[-2000:])
        return 1
    files = [ln for ln in out.stdout.splitlines() if ln.startswith(REF)]
    shutil.rmtree(DST, ignore_errors=True)
    n = 0
    for f in files:
        rel = os.path.relpath(f, REF)
        d = os.path.join(DST, "apps", "api", rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(f, d)
        n += 1
    shutil.copytree(os.path.join(ROOT, "oracle", "ref_import"), os.path.join(DST, "ref_import"),
                    ignore=shutil.ignore_patterns("__pycache__"))
    print(f"baseline/_ref: {n} reference files (unmodified) + the diffusers stand-in")
    return 0


if __name__ == "__main__":
    sys.exit(main())
