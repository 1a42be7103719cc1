"""Deferred evaluation for the reference's graph-style call pattern.

The reference builds TF-1.x op handles once (``loss_op, logits = model.compute_loss(
placeholders...)``, ``train_op = model.train(loss_op, ...)``, ``decode_op``, ``ler_op``) and
then evaluates subsets of them with ``sess.run(fetches, feed_dict)`` (SURVEY 3.4,
``models/test/test_ctc.py:118-175``, ``examples/timit/training/train_ctc.py:62-172``).  The
B200 models are eager; this module gives their methods a second, lazy mode: a method decorated
with ``@graph_op`` that receives a handle (placeholder or another op's output) returns a
handle instead of a value, and ``Session.run`` evaluates the handles it is asked for -- each
at most once per ``run`` call, dependencies first, exactly like a TF session executing a
sub-graph.  No arithmetic lives here.
"""
import functools

import numpy as np


class Tensor(object):
    """Base class of everything ``Session.run`` can be asked to evaluate."""
    name = None


class Placeholder(Tensor):
    def __init__(self, dtype=None, shape=None, name=None):
        self.dtype, self.shape, self.name = dtype, shape, name

    def __repr__(self):
        return "<placeholder %s>" % (self.name,)


class SparseTensor(Tensor):
    """``tf.SparseTensor(indices_pl, values_pl, dense_shape_pl)``: fed as one value (the
    ``list2sparsetensor`` triple / ``SparseTensorValue``) under the SparseTensor key itself."""

    def __init__(self, indices=None, values=None, dense_shape=None):
        self.indices, self.values, self.dense_shape = indices, values, dense_shape


class Op(Tensor):
    def __init__(self, fn, args, kwargs, name=None):
        self.fn, self.args, self.kwargs, self.name = fn, args, kwargs, name

    def __iter__(self):              # tuple-unpacking of multi-output ops: a, b = model.compute_loss(...)
        n = getattr(self, "n_out", None)
        if n is None:
            raise TypeError("op %r has a single output" % (self.name,))
        return iter([OpOutput(self, i) for i in range(n)])

    def __getitem__(self, i):
        return OpOutput(self, i)


class OpOutput(Tensor):
    def __init__(self, op, index):
        self.op, self.index = op, index
        self.name = "%s:%d" % (op.name, index)


class LazyGradsAndVars(Tensor):
    """What ``optimizer.compute_gradients(tower_loss)`` / ``model._clip_gradients(...)`` /
    ``average_gradients(...)`` hand around in graph mode (examples/librispeech/training/train_ctc.py:112-143):
    a handle on the op that, when run, yields the tower's ``[(grad, var)]`` list."""

    def __init__(self, op):
        self.op = op
        self.name = op.name


class Constant(Tensor):
    def __init__(self, value):
        self.value = value


def is_handle(x):
    if isinstance(x, Tensor):
        return True
    if isinstance(x, (list, tuple)):
        return any(is_handle(v) for v in x)
    return False


def graph_op(n_out=None, name=None):
    """Method decorator: lazy when any argument is a handle, eager otherwise."""
    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            if any(is_handle(a) for a in args) or any(is_handle(v) for v in kwargs.values()):
                op = Op(fn, args, kwargs, name=name or fn.__name__)
                if n_out:
                    op.n_out = n_out
                return op
            return fn(*args, **kwargs)
        wrapper.eager = fn
        return wrapper
    return deco


class _Run(object):
    """One ``Session.run`` call: feed values + memo of evaluated ops."""

    def __init__(self, feed_dict):
        self.feed = {}
        for k, v in (feed_dict or {}).items():
            self.feed[id(k)] = v
        self.memo = {}

    def resolve(self, x):
        if isinstance(x, Placeholder) or isinstance(x, SparseTensor):
            if id(x) not in self.feed:
                raise ValueError("You must feed a value for placeholder %r" % (getattr(x, "name", x),))
            return self.feed[id(x)]
        if isinstance(x, Constant):
            return x.value
        if isinstance(x, OpOutput):
            return self.resolve(x.op)[x.index]
        if isinstance(x, LazyGradsAndVars):
            return self.resolve(x.op)
        if isinstance(x, Op):
            if id(x) in self.feed:                       # TF allows feeding any tensor
                return self.feed[id(x)]
            if id(x) not in self.memo:
                args = [self.resolve(a) for a in x.args]
                kwargs = {k: self.resolve(v) for k, v in x.kwargs.items()}
                self.memo[id(x)] = x.fn(*args, **kwargs)
            return self.memo[id(x)]
        if isinstance(x, list):
            return [self.resolve(v) for v in x]
        if isinstance(x, tuple):
            return tuple(self.resolve(v) for v in x)
        return x


def to_host(v):
    """What ``sess.run`` hands back: numpy arrays / python scalars / SparseTensorValue / None."""
    try:
        import torch
    except ImportError:                                  # pragma: no cover
        torch = None
    if torch is not None and torch.is_tensor(v):
        a = v.detach().cpu().numpy()
        return a[()] if a.ndim == 0 else a
    if isinstance(v, tuple) and hasattr(v, "_fields"):   # namedtuples (SparseTensorValue, decoder outputs)
        return type(v)(*[to_host(f) for f in v])
    if isinstance(v, (list, tuple)):
        return type(v)(to_host(f) for f in v)
    if hasattr(v, "predicted_ids") and not isinstance(v, tuple):     # lazy decoder output
        return type("DecoderOutput", (), {f: to_host(getattr(v, f)) for f in
                                          ("logits", "predicted_ids", "decoder_output",
                                           "attention_weights", "context_vector")})()
    if isinstance(v, np.generic):
        return v
    return v
