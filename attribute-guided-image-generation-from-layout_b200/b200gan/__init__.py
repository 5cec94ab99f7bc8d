"""b200gan — host side of the B200-native G+D training step (PyTorch modules over libb200gan.so)."""
from . import _lib, ops  # noqa: F401
from .ops import set_precision, get_precision, bump_weight_epoch  # noqa: F401

# packed GEMM operands follow every torch.optim.Optimizer.step() (ops.WeightPacks explains why this is not left to
# tensor versions): one global post-step hook, installed with the package
ops.install_optimizer_hook()
