"""1-D group: `python -m tf_flash_attention_b200.tests.test_1d TestGroup.{list,verify,benchmark}`
(the reference's `flash_attention/tests/test_1d.py`; shape ranges from its table, test_1d.py:57-66)."""
import sys

import torch

from . import test_base


class TestGroup(test_base.TestGroup):
    SEQUENCE_DIMS = 1
    #           dtype: (minimum shape, maximum shape) of [batch, heads, channels, length]
    SHAPE_TABLE = {
        torch.float16: ([1, 8, 8, 256], [1, 8, 32, 4096]),
        torch.float32: ([1, 8, 8, 256], [1, 8, 32, 2048]),
        torch.float64: ([1, 8, 8, 256], [1, 8, 32, 1024]),
    }


if __name__ == "__main__":
    sys.exit(test_base.main(TestGroup))
