from .model import HunyuanVideo15Config, HunyuanVideo15Transformer3DModel, hy15_rope_table  # noqa: F401
