from .model import QwenImageConfig, QwenImageTransformer2DModel, qwen_rope_tables  # noqa: F401
