"""Encoder -> decoder bridge -- host mirror of ``models/attention/bridge.py``.

``InitialStateBridge(encoder_outputs, decoder_state_size, parameter_init)()`` flattens the
encoder's final state ((c_fw, h_fw), (c_bw, h_bw)), concatenates it along the feature axis,
applies one fully connected layer (identity activation, bias) of width sum(decoder state
sizes) and splits the result into the decoder cell's (c, h) (reference ``_create``).
``ZeroBridge`` / ``PassThroughBridge`` are the other two reference classes.
"""
from collections import namedtuple

import numpy as np
import torch

from ... import ops

LSTMStateTuple = namedtuple("LSTMStateTuple", ("c", "h"))


def _flatten(state):
    if isinstance(state, torch.Tensor):
        return [state]
    out = []
    for s in state:
        out.extend(_flatten(s))
    return out


class InitialStateBridge(object):
    def __init__(self, encoder_outputs, decoder_state_size, parameter_init, name="bridge"):
        if encoder_outputs is not None and not hasattr(encoder_outputs, "final_state"):
            raise ValueError("Invalid bridge_input not in encoder outputs.")
        self.encoder_outputs = encoder_outputs
        self.decoder_state_size = decoder_state_size          # LSTMStateTuple(c=Hd, h=Hd)
        self.parameter_init = parameter_init
        self.name = name
        self.variables = None

    def create_variables(self, bridge_input_size, rng, device):
        n_out = int(sum(self.decoder_state_size))
        std = self.parameter_init
        x = rng.normal(0, std, size=(bridge_input_size, n_out))
        bad = np.abs(x) > 2 * std                              # tf.truncated_normal: resample
        while bad.any():
            x[bad] = rng.normal(0, std, size=int(bad.sum()))
            bad = np.abs(x) > 2 * std
        self.variables = {"bridge/weights": torch.tensor(x.astype(np.float32), device=device),
                          "bridge/biases": torch.zeros(n_out, device=device)}
        return self.variables

    def __call__(self, encoder_outputs=None):
        enc = encoder_outputs if encoder_outputs is not None else self.encoder_outputs
        flat = torch.cat(_flatten(enc.final_state), dim=1).contiguous()      # [B, 4*H_enc]
        out = ops.gemm(flat, self.variables["bridge/weights"], bias=self.variables["bridge/biases"])
        sizes = list(self.decoder_state_size)
        parts, o = [], 0
        for s in sizes:
            parts.append(out[:, o:o + s].contiguous())
            o += s
        return LSTMStateTuple(*parts)


class ZeroBridge(object):
    def __init__(self, encoder_outputs, decoder_state_size, name="bridge"):
        self.encoder_outputs, self.decoder_state_size = encoder_outputs, decoder_state_size

    def __call__(self, encoder_outputs=None):
        enc = encoder_outputs if encoder_outputs is not None else self.encoder_outputs
        B, dev = enc.outputs.shape[0], enc.outputs.device
        return LSTMStateTuple(*[torch.zeros((B, s), device=dev) for s in self.decoder_state_size])


class PassThroughBridge(object):
    def __init__(self, encoder_outputs, decoder_state_size=None, name="bridge"):
        self.encoder_outputs = encoder_outputs

    def __call__(self, encoder_outputs=None):
        enc = encoder_outputs if encoder_outputs is not None else self.encoder_outputs
        return enc.final_state
