"""B200-native ConvLSTM recurrence: host-side mirror of the reference nn.Module API over libplc.so.

Import name: ``plconv`` (this directory's name has hyphens; ``plconv/__init__.py`` aliases it).
"""
from . import _lib, build, functional, gan, generator, losses, nn, parallel, rollout, trainer, training  # noqa: F401
from .losses import CombinedLoss  # noqa: F401
from .trainer import Trainer, TrainerConfig  # noqa: F401
from .generator import Generator  # noqa: F401
from .nn import ConvLSTMCell, ConvLSTMStack, EncoderForecaster  # noqa: F401
from .rollout import NowcastGenerator, NowcastRunner  # noqa: F401
from .gan import Discriminator, GanTrainStep  # noqa: F401
from ._lib import PLC_MODE_BF16_TC, PLC_MODE_FP32, invalidate_packed_weights  # noqa: F401
