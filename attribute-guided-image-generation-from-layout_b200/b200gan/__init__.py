"""b200gan — host side of the B200-native G+D training step (PyTorch modules over libb200gan.so)."""
from . import _lib, ops  # noqa: F401
from .ops import set_precision, get_precision, bump_weight_epoch  # noqa: F401
