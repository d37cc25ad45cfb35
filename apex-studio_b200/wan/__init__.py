from .model import WanConfig, WanTransformer3DModel  # noqa: F401
