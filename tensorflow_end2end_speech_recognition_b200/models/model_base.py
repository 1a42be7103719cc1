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
from ..compat import graph as _graph
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


class GradsAndVars(list):
    """``[(grad, var)]`` as tf.train.Optimizer.compute_gradients returns it, plus the flat buffer the
    gradients are views of (one tower = one flat fp32 buffer with the parameter layout)."""
    flat = None
    tensor_list = None


class Optimizer(object):
    """Stand-in for tf.train.*Optimizer: ``compute_gradients`` / ``apply_gradients`` (the in-graph tower flow of
    examples/librispeech/training/train_ctc.py:82-147) -- eager on tensors, lazy on graph handles."""

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
        self._towers_out = 0         # compute_gradients calls since the last apply_gradients

    def load_state(self, state):
        """slots + step counter from a checkpoint (compat.tf.train.Saver.restore)"""
        for k in ("state0", "state1"):
            st = getattr(self, k)
            if st is not None and k in state:
                st.copy_(torch.as_tensor(state[k]).to(st.device))
        self.global_step = int(state.get("global_step", self.global_step))

    def compute_gradients(self, loss):
        """Backward pass of the ``compute_loss`` that produced ``loss`` -> [(grad, var)].
        Tower k of a step (k-th call since the last ``apply_gradients``) writes its own flat buffer;
        tower 0 is the model's ``flat_grads`` itself."""
        if _graph.is_handle(loss):
            return _graph.LazyGradsAndVars(_graph.Op(self._compute_gradients, (loss,), {}, name="gradients"))
        return self._compute_gradients(loss)

    def _compute_gradients(self, loss):
        k = self._towers_out
        self._towers_out += 1
        flat, grads, tl = self.model._tower_buffers(k)
        ctx = getattr(loss, "_b2_ctx", None)
        if ctx is None and k > 0:
            raise RuntimeError("compute_gradients: a second tower needs the loss tensor of its own compute_loss "
                               "(this model type keeps one forward context)")
        if ctx is not None or k > 0:
            self.model._backward(ctx=ctx, flat=flat, grads=grads)
        else:
            self.model._backward()
        gv = GradsAndVars((grads[v.name], v) for v in self.model.trainable_variables())
        gv.flat, gv.tensor_list = flat, tl
        return gv

    def apply_gradients(self, grads_and_vars, global_step=None, learning_rate=None):
        if _graph.is_handle(grads_and_vars) or _graph.is_handle(self.learning_rate if learning_rate is None
                                                                else learning_rate):
            lr = self.learning_rate if learning_rate is None else learning_rate
            src = grads_and_vars.op if isinstance(grads_and_vars, _graph.LazyGradsAndVars) else grads_and_vars
            return _graph.Op(lambda gv, lr_: self._apply_gradients(gv, lr_), (src, lr), {}, name="apply_gradients")
        return self._apply_gradients(grads_and_vars, self.learning_rate if learning_rate is None else learning_rate)

    def _apply_gradients(self, grads_and_vars, lr):
        self.global_step += 1
        self._towers_out = 0
        flat = getattr(grads_and_vars, "flat", None)
        g = self._g
        if flat is not None and flat.data_ptr() != self.model.flat_grads.data_ptr():
            g = ops.TensorList([flat])
        elif flat is None and grads_and_vars is not None:
            # a plain [(grad, var)] list (e.g. the per-variable form of average_gradients)
            for gr, v in grads_and_vars:
                if gr is not None and gr.data_ptr() != v.grad.data_ptr():
                    v.grad.copy_(gr)
        # parameters, gradients and optimizer state are flat buffers with one layout,
        # so the whole update is a single elementwise launch
        ops.optimizer_step_multi(self.kind, self._p, g, self._s0, self._s1, float(lr), self.global_step)
        self.model._params_version += 1
        return self


class ModelBase(object):
    def __init__(self, *args, **kwargs):
        self.clip_grad_norm = None
        self._variables = []
        self.flat_params = None
        self.flat_grads = None
        self.world_size = 1
        self._pending = []
        self._towers = {}
        self._params_version = 0        # bumped by every optimizer step (packed-weight caches key on it)
        self._on_layer_done = None
        self._on_heads_done = None
        self._comm = None

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

    def _tower_buffers(self, k):
        """(flat gradient buffer, name -> view dict, TensorList) of tower k; tower 0 = flat_grads."""
        if k == 0:
            return self.flat_grads, self.grads, self._grad_list
        if k not in self._towers:
            flat = torch.zeros_like(self.flat_grads)
            base = self.flat_grads.data_ptr()
            views = {}
            for v in self._variables:
                o = (v.grad.data_ptr() - base) // 4
                views[v.name] = flat[o:o + v.grad.numel()].view(v.grad.shape)
            self._towers[k] = (flat, views, ops.TensorList([views[v.name] for v in self._variables]))
        return self._towers[k]

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
        1/world_size of the one-rank-per-GPU tower mean is folded into the same launch."""
        if isinstance(grads_and_vars, _graph.LazyGradsAndVars):
            return _graph.LazyGradsAndVars(_graph.Op(self._clip_gradients, (grads_and_vars.op,), {},
                                                     name="clip_gradients"))
        tl = getattr(grads_and_vars, "tensor_list", None) or self._grad_list
        post = 1.0 / self.world_size if (self.world_size > 1 and self._comm is None) else 1.0
        if self.clip_grad_norm is not None or post != 1.0:
            ops.clip_by_norm_multi(tl, self.clip_grad_norm, post_scale=post)
        if isinstance(grads_and_vars, GradsAndVars):
            return grads_and_vars
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
        self.optimizer._towers_out = 0
        # bucket mode needs every gradient outside the BLSTM layers final BEFORE the encoder's backward pass starts:
        # the plain BLSTM-CTC model without a weight-decay term (added at the end of _backward) qualifies
        bucketed = self.world_size > 1 and self._comm is not None and getattr(self, "bucketed_exchange", True) \
            and type(self).__name__ == "CTC" and getattr(self, "encoder_type", "") == "blstm" \
            and not getattr(self, "weight_decay", 0)
        if bucketed:
            # clip + mean per layer bucket, issued while BPTT of the layers below still runs
            self._bucketed_exchange_begin()
            grads_and_vars = self.optimizer._compute_gradients(loss)
            self._bucketed_exchange_end()
        else:
            grads_and_vars = self.optimizer._compute_gradients(loss)
            if self.clip_grad_norm is not None or self.world_size > 1:
                grads_and_vars = self._clip_gradients(grads_and_vars)
            self._allreduce_gradients()
        self.optimizer._apply_gradients(grads_and_vars, learning_rate)
        return self.optimizer

    # -------------------------------------------------------- data parallel
    def set_data_parallel(self, world_size, group=None, broadcast=True, comm=None):
        """One rank per GPU; replaces the in-graph towers of
        examples/librispeech/training/train_ctc.py:82-147.  ``comm``: a
        ``utils.training.multi_gpu.NcclComm`` -> the gradient mean goes through ``b2_allreduce_mean``
        (NCCL bound from C); without it ``torch.distributed`` all-reduce (gloo on CPU boxes)."""
        import torch.distributed as dist
        self.world_size, self._group, self._comm = int(world_size), group, comm
        if world_size > 1 and broadcast:
            dist.broadcast(self.flat_params, src=0, group=group)

    # ---- per-layer gradient buckets (clip per tensor locally, then mean over ranks), overlapped with BPTT
    def _bucket_plan(self):
        """[(first variable index, last+1)] per BLSTM layer (variables of a layer are contiguous in the flat buffer),
        plus one bucket for everything else (heads)."""
        if getattr(self, "_buckets", None) is None:
            import re
            groups, other = {}, []
            for i, v in enumerate(self._variables):
                m = re.match(r"blstm_hidden(\d+)/", v.name)
                (groups.setdefault(int(m.group(1)), []) if m else other).append(i)
            plan = {}
            for key, idx in list(groups.items()) + [(0, other)]:
                if not idx:
                    continue
                vs = [self._variables[i] for i in idx]
                base = self.flat_grads.data_ptr()
                lo = min((v.grad.data_ptr() - base) // 4 for v in vs)
                hi = max((v.grad.data_ptr() - base) // 4 + (v.grad.numel() + 3) // 4 * 4 for v in vs)
                plan[key] = (self.flat_grads[lo:hi], ops.TensorList([v.grad for v in vs]))
            self._buckets = plan
            self._comm_stream = torch.cuda.Stream(device=self.flat_grads.device)
            self._comm_done = torch.cuda.Event()
        return self._buckets

    def _reduce_bucket(self, key, stream):
        """per-tensor clip of bucket `key`, then its mean over the ranks, on `stream`"""
        flat, tl = self._bucket_plan()[key]
        with torch.cuda.stream(stream):
            if self.clip_grad_norm is not None:
                ops.clip_by_norm_multi(tl, self.clip_grad_norm, post_scale=1.0)
            self._comm.allreduce_mean_([flat], stream=stream)
        self._reduced.add(key)

    def _bucketed_exchange_begin(self):
        """installs the hooks _backward calls: after the head GEMMs (head bucket final) and after every layer"""
        self._bucket_plan()
        self._reduced = set()

        def on_heads_done():
            # called by _backward between the head GEMMs and the encoder's backward pass: head gradients are final
            if 0 in self._buckets:
                self._comm_stream.wait_stream(torch.cuda.current_stream())
                self._reduce_bucket(0, self._comm_stream)
        self._on_heads_done = on_heads_done

        def on_layer_done(i_layer):
            # layer i_layer+1's weight gradients were enqueued on the library's side stream during this call
            j = i_layer + 1
            if j in self._buckets and j not in self._reduced:
                ops.blstm_backward_side_wait(self._comm_stream)
                self._comm_stream.wait_stream(torch.cuda.current_stream())    # bias / peephole sums of layer j
                self._reduce_bucket(j, self._comm_stream)
        self._on_layer_done = on_layer_done

    def _bucketed_exchange_end(self):
        """after the backward pass (side stream joined): the remaining buckets, then wait for all of them"""
        cur = torch.cuda.current_stream()
        for key in sorted(self._buckets, reverse=True):
            if key not in self._reduced:
                self._reduce_bucket(key, cur)
        self._comm_done.record(self._comm_stream)
        cur.wait_event(self._comm_done)
        self._on_layer_done = None
        self._on_heads_done = None

    def _allreduce_gradients(self):
        if self.world_size <= 1:
            return
        if self._comm is not None:
            self._comm.allreduce_mean_([self.flat_grads])
            return
        from ..utils.training.multi_gpu import allreduce_mean_
        allreduce_mean_(self.flat_grads, self.world_size, group=getattr(self, "_group", None))
