from .wan import AutoencoderKLWan, WanVAEConfig  # noqa: F401
from .hunyuanvideo15 import AutoencoderKLHunyuanVideo15, HunyuanVideo15VAEConfig  # noqa: F401
