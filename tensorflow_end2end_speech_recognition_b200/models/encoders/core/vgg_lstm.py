"""VGG + unidirectional LSTM encoder -- host mirror of ``models/encoders/core/vgg_lstm.py``: the VGG
front-end of ``core/vgg_blstm.py`` (``b2_vgg_frontend_forward/backward``) feeding ``core/lstm.py``."""
from .lstm import LSTMEncoder
from .vgg_blstm import VGGBLSTMEncoder


class VGGLSTMEncoder(VGGBLSTMEncoder):
    def __init__(self, input_size, splice, num_stack, num_units, num_proj, num_layers, lstm_impl,
                 use_peephole, parameter_init, clip_activation, time_major=False,
                 name="vgg_lstm_encoder", precision="fp32", tf_version="1.2.0"):
        super(VGGLSTMEncoder, self).__init__(input_size, splice, num_stack, num_units, num_proj, num_layers,
                                             lstm_impl, use_peephole, parameter_init, clip_activation,
                                             time_major=time_major, name=name, precision=precision,
                                             tf_version=tf_version)
        # same attribute, one direction: VGGBLSTMEncoder drives it through the shared call / backward signature
        self.blstm = LSTMEncoder(num_units, num_proj, num_layers, lstm_impl, use_peephole, parameter_init,
                                 clip_activation, time_major=True, precision=precision, tf_version=tf_version)
        self.num_proj = self.blstm.num_proj
