"""128x128 generator — module surface of the reference's models/generator_obj_att128.py.

Same blocks as the 64x64 model plus AdaptiveAvgPool2d(8) in the LayoutEncoder (generator_obj_att128.py:486,505) and
the nearest x2 upsample + c5 / spade_4 / c6 / spade_5 / c7 refinement in the Decoder (generator_obj_att128.py:549-604).
"""
from models import generator_obj_att as _g64
from models.generator_obj_att import (ConditionalBatchNorm2d, ResidualBlock, ConvLSTMCell, LayoutConvLSTM,  # noqa: F401
                                      CropEncoder, GlobalEncoder, AttributeEncoder, get_z_random)


class LayoutEncoder(_g64.LayoutEncoder):
    def __init__(self, *a, **kw):
        kw["pool_to_8"] = True
        super().__init__(*a, **kw)


class Decoder(_g64.Decoder):
    def __init__(self, nf=64, conv_dim=64):
        super().__init__(nf=nf, conv_dim=conv_dim, image_size=128)


class Generator(_g64.Generator):
    _image_size = 128
