from .model import Flux2Config, Flux2Transformer2DModel  # noqa: F401
