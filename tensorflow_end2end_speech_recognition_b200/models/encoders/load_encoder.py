"""Encoder registry, same entry point as ``models/encoders/load_encoder.py:26-57``."""
from .core.blstm import BLSTMEncoder
from .core.vgg_blstm import VGGBLSTMEncoder

ENCODERS = {"blstm": BLSTMEncoder, "vgg_blstm": VGGBLSTMEncoder}


def load(encoder_type):
    """Select & load encoder (reference: load_encoder.py:46-57)."""
    if encoder_type not in ENCODERS:
        raise ValueError("encoder_type should be one of [%s], you provided %s." %
                         (", ".join(ENCODERS), encoder_type))
    return ENCODERS[encoder_type]
