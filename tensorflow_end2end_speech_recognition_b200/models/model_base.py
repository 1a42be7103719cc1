"""Base class for all models -- host mirror of ``models/model_base.py``.

Keeps the reference's train-step surface: ``OPTIMIZER_CLS_NAMES``
(model_base.py:12-20), ``_set_optimizer`` (:68-95), ``train(loss, optimizer,
learning_rate)`` (:97-133) and ``_clip_gradients`` (:135-166), executed eagerly:
``train`` runs backward -> per-tensor clip_by_norm -> (NCCL all-reduce mean across
ranks, replacing utils/training/multi_gpu.py's tower averaging) -> optimizer.
All arithmetic is CUDA (optim.cu); torch tensors are containers.
"""
import numpy as np
import torch

from .. import ops
from ..compat.graph import graph_op

OPTIMIZER_CLS_NAMES = {          # name -> (kernel kind, number of state slots, slot-0 init)
    "adagrad": ("adagrad", 1, 0.1),
    "adadelta": ("adadelta", 2, 0.0),
    "adam": ("adam", 2, 0.0),
    "rmsprop": ("rmsprop", 1, 1.0),
    "sgd": ("sgd", 0, 0.0),
    "momentum": ("momentum", 1, 0.0),
    "nestrov": ("nestrov", 1, 0.0),
}


class Variable(object):
    """A named view into the model's flat fp32 parameter buffer (TF variable names)."""

    def __init__(self, name, tensor, grad):
        self.name, self.tensor, self.grad = name, tensor, grad

    def get_shape(self):
        return tuple(self.tensor.shape)


class Optimizer(object):
    """Eager stand-in for tf.train.*Optimizer: compute_gradients / apply_gradients."""

    def __init__(self, name, learning_rate, model):
        self.name, self.learning_rate, self.model = name, learning_rate, model
        kind, nstate, init0 = OPTIMIZER_CLS_NAMES[name]
        self.kind = kind
        flat = model.flat_params
        self.state0 = torch.full_like(flat, init0) if nstate >= 1 else None
        self.state1 = torch.zeros_like(flat) if nstate >= 2 else None
        self._p = ops.TensorList([flat])
        self._g = ops.TensorList([model.flat_grads])
        self._s0 = ops.TensorList([self.state0]) if self.state0 is not None else None
        self._s1 = ops.TensorList([self.state1]) if self.state1 is not None else None
        self.global_step = 0

    def load_state(self, state):
        """slots + step counter from a checkpoint (compat.tf.train.Saver.restore)"""
        for k in ("state0", "state1"):
            st = getattr(self, k)
            if st is not None and k in state:
                st.copy_(torch.as_tensor(state[k]).to(st.device))
        self.global_step = int(state.get("global_step", self.global_step))

    def compute_gradients(self, loss):
        """Runs the backward pass of the last compute_loss -> [(grad, var)]."""
        self.model._backward()
        return [(v.grad, v) for v in self.model.trainable_variables()]

    def apply_gradients(self, grads_and_vars, global_step=None, learning_rate=None):
        lr = self.learning_rate if learning_rate is None else learning_rate
        self.global_step += 1
        # parameters, gradients and optimizer state are flat buffers with one layout,
        # so the whole update is a single elementwise launch
        ops.optimizer_step_multi(self.kind, self._p, self._g, self._s0, self._s1, float(lr),
                                 self.global_step)


class ModelBase(object):
    def __init__(self, *args, **kwargs):
        self.clip_grad_norm = None
        self._variables = []
        self.flat_params = None
        self.flat_grads = None
        self.world_size = 1
        self._pending = []

    # ------------------------------------------------------------ variables
    def _allocate_variables(self, named_arrays, device):
        """named_arrays: ordered [(tf_name, numpy)] -> flat parameter / gradient buffers + views.
        Every variable starts on a 16-byte boundary (TMA / vector loads)."""
        offs, total = [], 0
        for _, a in named_arrays:
            offs.append(total)
            total += (a.size + 3) // 4 * 4
        host = np.zeros(total, np.float32)
        for (n, a), o in zip(named_arrays, offs):
            host[o:o + a.size] = a.reshape(-1)
        self.flat_params = torch.tensor(host, device=device)
        self.flat_grads = torch.zeros_like(self.flat_params)
        self._variables = []
        self.variables, self.grads = {}, {}
        for (n, a), o in zip(named_arrays, offs):
            t = self.flat_params[o:o + a.size].view(a.shape)
            g = self.flat_grads[o:o + a.size].view(a.shape)
            self._variables.append(Variable(n, t, g))
            self.variables[n], self.grads[n] = t, g
        self._grad_list = ops.TensorList([v.grad for v in self._variables])

    def trainable_variables(self):
        return list(self._variables)

    # ------------------------------------------------------------ optimizer
    def _set_optimizer(self, optimizer, learning_rate):
        """(reference: model_base.py:68-95)"""
        optimizer = optimizer.lower()
        if optimizer not in OPTIMIZER_CLS_NAMES:
            raise ValueError("Optimizer name should be one of [%s], you provided %s." %
                             (", ".join(OPTIMIZER_CLS_NAMES), optimizer))
        return Optimizer(optimizer, learning_rate, self)

    def _clip_gradients(self, grads_and_vars):
        """Per-tensor tf.clip_by_norm, in place (reference: model_base.py:135-166).  The
        1/world_size of the tower mean is folded into the same launch."""
        ops.clip_by_norm_multi(self._grad_list, self.clip_grad_norm, post_scale=1.0 / self.world_size)
        return [(g, v) for g, v in grads_and_vars if g is not None]

    @graph_op(name="train")
    def train(self, loss, optimizer, learning_rate):
        """One optimisation step on the loss of the last ``compute_loss``
        (reference: model_base.py:97-133).  ``optimizer`` is a name from
        OPTIMIZER_CLS_NAMES (kept across calls) ; returns the optimizer object."""
        if getattr(self, "optimizer", None) is None or self.optimizer.name != optimizer.lower():
            self.optimizer = self._set_optimizer(optimizer, learning_rate)
            restored = getattr(self, "_restored_optimizer_state", None)
            if restored is not None and restored.get("name") == self.optimizer.name:
                self.optimizer.load_state(restored)
            self._restored_optimizer_state = None
        grads_and_vars = self.optimizer.compute_gradients(loss)
        if self.clip_grad_norm is not None or self.world_size > 1:
            grads_and_vars = self._clip_gradients(grads_and_vars)
        self._allreduce_gradients()
        self.optimizer.apply_gradients(grads_and_vars, learning_rate=learning_rate)
        return self.optimizer

    # -------------------------------------------------------- data parallel
    def set_data_parallel(self, world_size, group=None, broadcast=True):
        """One rank per GPU; replaces the in-graph towers of
        examples/librispeech/training/train_ctc.py:82-147."""
        import torch.distributed as dist
        self.world_size, self._group = int(world_size), group
        if world_size > 1 and broadcast:
            dist.broadcast(self.flat_params, src=0, group=group)

    def _allreduce_gradients(self):
        if self.world_size <= 1:
            return
        from ..utils.training.multi_gpu import allreduce_mean_
        allreduce_mean_(self.flat_grads, self.world_size, group=getattr(self, "_group", None))
