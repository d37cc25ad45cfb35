import logging as _logging

USE_PEFT_BACKEND = False


class _Logging:
    @staticmethod
    def get_logger(name):
        return _logging.getLogger(name)


logging = _Logging()


def scale_lora_layers(model, weight):
    return None


def unscale_lora_layers(model, weight=None):
    return None


def deprecate(*args, **kwargs):
    return None


class BaseOutput(dict):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.__dict__.update(kw)


def is_torch_npu_available():
    return False
