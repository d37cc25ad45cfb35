from enum import Enum

from ..utils import BaseOutput


class KarrasDiffusionSchedulers(Enum):
    UniPCMultistepScheduler = 1


class SchedulerOutput(BaseOutput):
    pass


class SchedulerMixin:
    config_name = "scheduler_config.json"
