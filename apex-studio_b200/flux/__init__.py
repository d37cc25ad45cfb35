from .model import FluxConfig, FluxTransformer2DModel, flux_rope_table  # noqa: F401
