"""Encoder registry, same entry point as ``models/encoders/load_encoder.py:26-57``."""
from .core.blstm import BLSTMEncoder
from .core.multitask_blstm import MultitaskBLSTMEncoder
from .core.vgg_blstm import VGGBLSTMEncoder

# the encoders on the B200 hot path; the reference's other entries (uni-directional lstm, gru/bgru, vgg_lstm,
# cnn_zhang, vgg_wang, pyramid_blstm, cldnn_wang, student_*; load_encoder.py:26-44) are not built
ENCODERS = {"blstm": BLSTMEncoder, "vgg_blstm": VGGBLSTMEncoder, "multitask_blstm": MultitaskBLSTMEncoder}


def load(encoder_type):
    """Select & load encoder (reference: load_encoder.py:46-57)."""
    if encoder_type not in ENCODERS:
        raise ValueError("encoder_type should be one of [%s], you provided %s." %
                         (", ".join(ENCODERS), encoder_type))
    return ENCODERS[encoder_type]
