"""apex-studio_b200 -- sm_100a (B200) implementation of the denoising hot path of Apex Studio's
generation server (reference: totokunda/apex-studio, apps/api): the per-step Wan DiT forward, the
attention operator behind ``attention_register``, the UniPC step indexing and the Wan 3D-VAE decode.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/apex_b200.h) -> libapex_b200.so
  _lib.py          ctypes loader for libapex_b200.so (fails loudly when the library is missing)
  ops.py           torch-tensor front end of the C ABI (pointers/strides/stream extraction, error mapping)
  register.py      FunctionRegister mirror of apps/api/src/register/__init__.py
  attention.py     ``attention_register`` mirror + the "b200" backend (boundary #1 of SURVEY.md section 8b)
  wan/             WanTransformer3DModel / WanTransformerBlock host-side mirror driving the kernels
  scheduler.py     flow-UniPC scheduler (integer step indexing bit-exact with the reference)
  denoise.py       moe_denoise / base_denoise loop mirror (CFG, expert switch, scheduler step)
  parallel.py      one-process-per-GPU sharding (CFG pair, token shards) over torch.distributed/NCCL

The CPU oracle (``oracle/``) is test infrastructure and is never imported from here.
"""
__version__ = "0.1.0"
