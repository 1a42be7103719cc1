"""Encoder registry, same entry point as ``models/encoders/load_encoder.py:26-57``."""
from .core.blstm import BLSTMEncoder
from .core.gru import BGRUEncoder, GRUEncoder
from .core.lstm import LSTMEncoder, MultitaskLSTMEncoder
from .core.multitask_blstm import MultitaskBLSTMEncoder
from .core.vgg_blstm import VGGBLSTMEncoder
from .core.vgg_lstm import VGGLSTMEncoder

# the encoders built on the B200 kernels; the reference's other entries (cnn_zhang, vgg_wang, pyramid_blstm,
# cldnn_wang, student_*; load_encoder.py:26-44) are not built
ENCODERS = {"blstm": BLSTMEncoder, "lstm": LSTMEncoder, "bgru": BGRUEncoder, "gru": GRUEncoder, "vgg_blstm": VGGBLSTMEncoder, "vgg_lstm": VGGLSTMEncoder,
            "multitask_blstm": MultitaskBLSTMEncoder, "multitask_lstm": MultitaskLSTMEncoder}


def load(encoder_type):
    """Select & load encoder (reference: load_encoder.py:46-57)."""
    if encoder_type not in ENCODERS:
        raise ValueError("encoder_type should be one of [%s], you provided %s." %
                         (", ".join(ENCODERS), encoder_type))
    return ENCODERS[encoder_type]
