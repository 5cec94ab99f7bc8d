"""Placeholder for the reference's models/layers.py (sg2im builder helpers).  The reference's discriminator imports
build_cnn / GlobalAvgPool from it but never calls them, so they are outside the hot path."""


def build_cnn(*args, **kwargs):
    raise NotImplementedError("models.layers.build_cnn is not on the G+D hot path")


class GlobalAvgPool:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("models.layers.GlobalAvgPool is not on the G+D hot path")
