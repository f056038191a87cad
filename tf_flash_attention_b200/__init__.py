"""B200-native drop-in for nothingstopsme/tf_flash_attention."""
from . import flash_attention  # noqa: F401
from .flash_attention import (causal_1d, causal_2d, full_1d, full_2d, local_1d, local_2d)  # noqa: F401
