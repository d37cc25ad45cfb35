class FromOriginalModelMixin:
    pass


class PeftAdapterMixin:
    pass


class FluxTransformer2DLoadersMixin:
    pass
