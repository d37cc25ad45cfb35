"""Minimal stand-in for the `diffusers` package (NOT installed in this image, and un-pinned by the
reference: `diffusers @ git+https://github.com/huggingface/diffusers.git`, apps/api/requirements/requirements.txt).

TEST INFRASTRUCTURE ONLY.  It exists so that oracle/make_golden.py can import the reference's own,
unmodified Wan transformer / VAE / scheduler modules from /root/reference in this container and record
golden vectors.  Only the primitives those modules use are restated, from knowledge of upstream
diffusers semantics (not from source in the tree):
  FeedForward("gelu-approximate") = Linear(+bias) -> F.gelu(approximate="tanh") -> Dropout(0) -> Linear(+bias)
  FP32LayerNorm                   = F.layer_norm(x.float(), ..., eps).to(x.dtype)
  Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0) = cat[cos, sin](t * exp(-ln(1e4) * i / 128))
  TimestepEmbedding               = Linear -> SiLU -> Linear
  PixArtAlphaTextProjection("gelu_tanh") = Linear -> GELU(tanh) -> Linear
Nothing in the product package imports this.
"""
__version__ = "0.0-standin"
