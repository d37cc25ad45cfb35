"""CUDA-graph capture of a launch-bound inner loop.

The image models run 500-900 kernels of ~100 us per denoise step (FLUX.1-dev: 524 launches in 73 ms).  Every tensor the
forward touches is either a weight, a preallocated workspace or an input, and every kernel is launched through the C ABI on
the CURRENT stream with host-built TMA descriptors passed by value, so a whole step can be captured once and replayed:
``GraphedCallable`` copies new inputs into static buffers, replays the graph and returns the static outputs.

Results are bit-identical to eager execution (same kernels, same order; tests/test_gpu_flux.py).  The reference has no
counterpart: its ``torch.compile`` hooks (engine/base_engine.py) trace PyTorch ops; here there is nothing to trace.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Sequence, Tuple

import torch


class GraphedCallable:
    """Capture ``fn(*tensor_args, **static_kwargs)`` into a CUDA graph; tensor args are refreshed by copy on each call.

    * ``fn`` must be shape-static, allocation-stable after warm-up and free of host<->device synchronisation
      (the B200 model forwards are, once their RoPE tables / workspaces exist -- hence the warm-up runs);
    * non-tensor keyword arguments (ids living on the host, flags) are bound at capture time;
    * the returned tensors are the graph's static outputs: consume or clone them before the next call.
    """

    def __init__(self, fn: Callable[..., Any], example_args: Sequence[torch.Tensor], static_kwargs: Optional[Dict[str, Any]] = None,
                 warmup: int = 2):
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA graphs need a CUDA device (the b200 path has no CPU fallback)")
        for a in example_args:
            if not (isinstance(a, torch.Tensor) and a.is_cuda):
                raise ValueError("every positional argument must be a CUDA tensor (host inputs go into static_kwargs)")
        self._fn = fn
        self._kwargs = dict(static_kwargs or {})
        self._static_in = [a.clone() for a in example_args]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn(*self._static_in, **self._kwargs)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._static_out = fn(*self._static_in, **self._kwargs)

    def __call__(self, *args: torch.Tensor):
        if len(args) != len(self._static_in):
            raise ValueError(f"expected {len(self._static_in)} tensor arguments, got {len(args)}")
        for dst, src in zip(self._static_in, args):
            if dst.shape != src.shape or dst.dtype != src.dtype:
                raise ValueError(f"graph captured for {tuple(dst.shape)} {dst.dtype}, got {tuple(src.shape)} {src.dtype}")
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src)
        self.graph.replay()
        return self._static_out
