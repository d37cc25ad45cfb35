from .wan import AutoencoderKLWan, WanVAEConfig  # noqa: F401
