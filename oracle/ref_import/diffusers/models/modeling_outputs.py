from ..utils import BaseOutput


class Transformer2DModelOutput(BaseOutput):
    pass


class AutoencoderKLOutput(BaseOutput):
    pass
