"""BLSTM-CTC model and train step, CPU restatement (torch-CPU).  TEST INFRASTRUCTURE.

Follows ``models/ctc/ctc.py:175-238`` (``_build``: encoder -> output
fully_connected -> logits [T,B,C]), ``:256-323`` (``compute_loss``: mean of
``tf.nn.ctc_loss``), ``models/model_base.py:97-166`` (``train``: gradients ->
per-tensor clip_by_norm -> optimizer) and ``utils/training/multi_gpu.py:13-48``.
This is also the CPU baseline that bench.py times next to the GPU path (the real
TF-1.x CPU path cannot run here: TensorFlow is not installable).
"""
import numpy as np
import torch

from . import lstm as olstm
from . import optim as oopt


def layers_from_variables(variables, num_layers, use_peephole=True):
    layers = []
    for i in range(1, num_layers + 1):
        layer = {}
        for d in ("fw", "bw"):
            scope = "blstm_hidden%d/%s/lstm_cell/" % (i, d)
            p = {"kernel": variables[scope + "kernel"], "bias": variables[scope + "bias"]}
            if use_peephole:
                for k in ("w_i_diag", "w_f_diag", "w_o_diag"):
                    p[k] = variables[scope + k]
            if scope + "projection/kernel" in variables:                 # LSTMCell(num_proj), blstm.py:215-228
                p["projection"] = variables[scope + "projection/kernel"]
            layer[d] = p
        layers.append(layer)
    return layers


def uni_layers_from_variables(variables, num_layers, use_peephole=True):
    """parameter dicts of the unidirectional stack (MultiRNNCell variable names, encoders/core/lstm.py:127-166)"""
    layers = []
    for i in range(num_layers):
        scope = "multi_lstm/multi_rnn_cell/cell_%d/lstm_cell/" % i
        p = {"kernel": variables[scope + "kernel"], "bias": variables[scope + "bias"]}
        if use_peephole:
            for k in ("w_i_diag", "w_f_diag", "w_o_diag"):
                p[k] = variables[scope + k]
        layers.append(p)
    return layers


def gru_layers_from_variables(variables, num_layers, bidirectional):
    layers = []
    for i in range(1, num_layers + 1):
        layer = {}
        for d in (("fw", "bw") if bidirectional else ("fw",)):
            scope = ("bgru_hidden%d/%s/gru_cell/" % (i, d)) if bidirectional else \
                ("multi_gru/multi_rnn_cell/cell_%d/gru_cell/" % (i - 1))
            layer[d] = {k: variables[scope + k] for k in ("gates/kernel", "gates/bias", "candidate/kernel",
                                                          "candidate/bias")}
        layers.append(layer)
    return layers


def ctc_model_forward(variables, inputs_btd, seq_len, labels, num_layers, use_peephole=True,
                      cell_clip=None, keep_prob=1.0, dropout_masks=None, vgg=None, unidirectional=False, gru=None):
    """variables: dict name -> torch tensor.  Returns (mean loss, logits [T,B,C], per-utt losses).
    vgg = (num_channels, width): run the VGG front-end (oracle/vgg.py) before the BLSTM stack
    (encoder_type 'vgg_blstm', ctc.py:135-147)."""
    layers = None if (unidirectional or gru) else layers_from_variables(variables, num_layers, use_peephole)
    if vgg is not None:
        from . import vgg as ovgg
        inputs_btd = ovgg.vgg_frontend(inputs_btd, variables, vgg[0], vgg[1])
    if gru:                     # encoder_type 'gru' / 'bgru' (gru = "gru" | "bgru")
        enc, _ = olstm.gru_forward(inputs_btd, seq_len, gru_layers_from_variables(variables, num_layers, gru == "bgru"),
                                   gru == "bgru")
    elif unidirectional:        # encoder_type 'lstm' / 'vgg_lstm'
        enc, _ = olstm.lstm_forward(inputs_btd, seq_len, uni_layers_from_variables(variables, num_layers, use_peephole),
                                    keep_prob=keep_prob, dropout_masks=dropout_masks, cell_clip=cell_clip)
    else:
        enc, _ = olstm.blstm_forward(inputs_btd, seq_len, layers, keep_prob=keep_prob,
                                     dropout_masks=dropout_masks, cell_clip=cell_clip)
    T, B, E = enc.shape
    feat = enc.reshape(T * B, E)
    if "bottleneck/weights" in variables:                               # ctc.py:200-213 (keep_prob 1)
        feat = torch.relu(feat @ variables["bottleneck/weights"] + variables["bottleneck/biases"])
    logits = (feat @ variables["output/weights"] + variables["output/biases"])
    logits = logits.reshape(T, B, -1)
    C = logits.shape[-1]
    lens = torch.tensor([len(l) for l in labels], dtype=torch.long)
    flat = torch.tensor([v for l in labels for v in l], dtype=torch.long)
    losses = torch.nn.functional.ctc_loss(torch.log_softmax(logits, -1), flat,
                                          torch.as_tensor(np.asarray(seq_len), dtype=torch.long), lens,
                                          blank=C - 1, reduction="none", zero_infinity=False)
    # ignore_longer_outputs_than_inputs=True: utterances with L > T_b contribute 0 (ctc.py:296)
    skip = lens > torch.as_tensor(np.asarray(seq_len), dtype=torch.long)
    losses = torch.where(skip, torch.zeros_like(losses), losses)
    return losses.mean(), logits, losses


class OracleTrainer(object):
    """variables as numpy fp64/fp32 arrays; autograd for the backward pass; clip and
    optimizer from oracle.optim (TF-1.x rules)."""

    def __init__(self, variables, num_layers, optimizer="rmsprop", learning_rate=1e-3,
                 clip_grad_norm=None, use_peephole=True, cell_clip=None, dtype=torch.float64):
        self.names = list(variables)
        self.params = [np.array(variables[n], dtype=np.float64 if dtype == torch.float64 else np.float32)
                       for n in self.names]
        self.num_layers, self.use_peephole, self.cell_clip = num_layers, use_peephole, cell_clip
        self.clip = clip_grad_norm
        self.opt = oopt.Optimizer(optimizer, learning_rate)
        self.dtype = dtype

    def loss_and_grads(self, inputs_btd, seq_len, labels):
        vs = {n: torch.tensor(p, dtype=self.dtype, requires_grad=True) for n, p in zip(self.names, self.params)}
        x = torch.tensor(np.asarray(inputs_btd), dtype=self.dtype)
        loss, logits, losses = ctc_model_forward(vs, x, seq_len, labels, self.num_layers,
                                                 self.use_peephole, self.cell_clip)
        loss.backward()
        grads = [vs[n].grad.numpy() if vs[n].grad is not None else None for n in self.names]
        return float(loss.detach()), logits.detach().numpy(), grads

    def step(self, inputs_btd, seq_len, labels, tower_grads=None):
        loss, logits, grads = self.loss_and_grads(inputs_btd, seq_len, labels)
        if self.clip is not None:
            grads = [oopt.clip_by_norm(g, self.clip) if g is not None else None for g in grads]
        self.opt.step(self.params, grads)
        return loss, logits, grads


def multitask_ctc_forward(variables, inputs_btd, seq_len, labels_main, labels_sub, num_layers_main,
                          num_layers_sub, main_task_weight, use_peephole=True, unidirectional=False):
    """Hierarchical CTC (models/ctc/multitask_ctc.py:109-191,225-296): main head on the top BLSTM layer, sub head on
    layer ``num_layers_sub`` (blstm.py:325-331); total = w * mean(ctc_main) + (1 - w) * mean(ctc_sub).
    variables: dict name -> torch tensor.  Returns (total, logits_main [T,B,Cm], logits_sub [T,B,Cs])."""
    layers = (uni_layers_from_variables if unidirectional else layers_from_variables)(variables, num_layers_main,
                                                                                      use_peephole)
    x = inputs_btd
    enc_sub = None
    for i, layer in enumerate(layers, 1):
        y, _ = (olstm.lstm_forward if unidirectional else olstm.blstm_forward)(x, seq_len, [layer])
        if i == num_layers_sub:
            enc_sub = y
        x = y.transpose(0, 1)
    enc = y
    T, B, E = enc.shape
    lens_t = torch.as_tensor(np.asarray(seq_len), dtype=torch.long)

    def head(feat, w, b, labels):
        logits = (feat.reshape(T * B, E) @ variables[w] + variables[b]).reshape(T, B, -1)
        C = logits.shape[-1]
        lens = torch.tensor([len(l) for l in labels], dtype=torch.long)
        flat = torch.tensor([v for l in labels for v in l], dtype=torch.long)
        losses = torch.nn.functional.ctc_loss(torch.log_softmax(logits, -1), flat, lens_t, lens, blank=C - 1,
                                              reduction="none", zero_infinity=False)
        return losses.mean(), logits
    l_main, logits_main = head(enc, "output_main/weights", "output_main/biases", labels_main)
    l_sub, logits_sub = head(enc_sub, "output_sub/weights", "output_sub/biases", labels_sub)
    return main_task_weight * l_main + (1.0 - main_task_weight) * l_sub, logits_main, logits_sub
