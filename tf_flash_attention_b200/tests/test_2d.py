"""2-D group: `python -m tf_flash_attention_b200.tests.test_2d TestGroup.{list,verify,benchmark}`
(the reference's `flash_attention/tests/test_2d.py`; shape ranges from its table, test_2d.py:85-94)."""
import sys

import torch

from . import test_base


class TestGroup(test_base.TestGroup):
    SEQUENCE_DIMS = 2
    #           dtype: (minimum shape, maximum shape) of [batch, heads, channels, height, width]
    SHAPE_TABLE = {
        torch.float16: ([1, 8, 8, 16, 16], [1, 8, 32, 64, 64]),
        torch.float32: ([1, 8, 8, 16, 16], [1, 8, 32, 32, 64]),
        torch.float64: ([1, 8, 8, 16, 16], [1, 8, 32, 32, 32]),
    }


if __name__ == "__main__":
    sys.exit(test_base.main(TestGroup))
