"""The slice of the ``tensorflow`` 1.x module the reference's training / evaluation scripts and
``models/test`` touch, re-expressed over ``compat.graph``: enough for

    with tf.Graph().as_default():
        model.create_placeholders(); lr = tf.placeholder(tf.float32, name='learning_rate')
        loss_op, logits = model.compute_loss(model.inputs_pl_list[0], ...)
        train_op = model.train(loss_op, optimizer='adam', learning_rate=lr)
        ...
        with tf.Session() as sess:
            sess.run(tf.global_variables_initializer())
            _, loss = sess.run([train_op, loss_op], feed_dict={...})

to run unchanged on the B200 models.  ``compat.install()`` registers this module as
``tensorflow`` so that ``import tensorflow as tf`` in a driver script resolves here.
Nothing here computes: ops evaluate through the models' CUDA paths.
"""
import contextlib
import os
import unittest

import numpy as np

from . import graph as _g
from ..utils.io.labels.sparsetensor import SparseTensorValue  # noqa: F401  (tf.SparseTensorValue)

float32, float64, int32, int64, bool = "float32", "float64", "int32", "int64", "bool"   # noqa: A001
__version__ = "1.2.0"        # the pinned version of the reference (requirements.txt:11)

SparseTensor = _g.SparseTensor
_default_graph = None


class Graph(object):
    def __init__(self):
        self.models = []

    @contextlib.contextmanager
    def as_default(self):
        global _default_graph
        prev, _default_graph = _default_graph, self
        try:
            yield self
        finally:
            _default_graph = prev


def get_default_graph():
    global _default_graph
    if _default_graph is None:
        _default_graph = Graph()
    return _default_graph


def reset_default_graph():
    global _default_graph
    _default_graph = Graph()


def register_model(model):
    """Models announce themselves so tf.trainable_variables() / Saver can find their variables."""
    g = get_default_graph()
    if model not in g.models:
        g.models.append(model)


def placeholder(dtype, shape=None, name=None):
    return _g.Placeholder(dtype, shape, name)


@contextlib.contextmanager
def _noop_scope(*args, **kwargs):
    yield None


device = name_scope = _noop_scope


class _VarScope(object):
    def reuse_variables(self):
        pass


@contextlib.contextmanager
def variable_scope(*args, **kwargs):
    yield _VarScope()


def get_variable_scope():
    return _VarScope()


class _Dim(object):
    def __init__(self, v):
        self.value = int(v)


class _VarView(object):
    """What utils/parameter.py::count_total_parameters needs (:14-20): .name, .get_shape() -> dims with .value"""

    def __init__(self, v):
        self._v, self.name = v, v.name

    def get_shape(self):
        return [_Dim(d) for d in self._v.tensor.shape]


def trainable_variables():
    return [_VarView(v) for m in get_default_graph().models for v in m.trainable_variables()]


def Variable(initial_value, name=None, trainable=True):   # noqa: N802  (global_step only)
    return _g.Constant(initial_value)


def global_variables_initializer():
    return _g.Op(lambda: None, (), {}, name="init")       # variables are initialised at construction


def _host(v):
    return _g.to_host(v)


def expand_dims(x, axis=0):
    def fn(v):
        try:
            import torch
        except ImportError:                               # pragma: no cover
            torch = None
        if torch is not None and torch.is_tensor(v):
            # stays a device tensor and keeps the backward context of the loss it wraps:
            # ``optimizer.compute_gradients(tf.expand_dims(tower_loss, axis=0))``, train_ctc.py:104-112
            r = v.unsqueeze(axis)
            if hasattr(v, "_b2_ctx"):
                r._b2_ctx = v._b2_ctx
            return r
        return np.expand_dims(np.asarray(_host(v)), axis)
    return _g.Op(fn, (x,), {}, name="expand_dims")


def concat(values, axis=0):
    return _g.Op(lambda vs: np.concatenate([np.asarray(_host(v)) for v in vs], axis), (list(values),), {},
                 name="concat")


def reduce_mean(x, axis=None, name=None):
    return _g.Op(lambda v: np.mean(np.asarray(_host(v)), axis=axis), (x,), {}, name=name or "reduce_mean")


def edit_distance(hypothesis, truth, normalize=True):
    from ..utils.io.labels.sparsetensor import sparse_to_label_lists

    def fn(h, t):
        import torch
        B = int(np.asarray(t[2])[0])
        hl, tl = sparse_to_label_lists(h, B), sparse_to_label_lists(t, B)
        if torch.cuda.is_available():
            from .. import ops
            d = ops.edit_distance(hl, tl, "cuda:%d" % torch.cuda.current_device()).astype(np.float64)
        else:                                   # graph plumbing tests on a CPU-only box
            from ..models.ctc.ctc import _edit_distance
            d = np.asarray([_edit_distance(a, b) for a, b in zip(hl, tl)], np.float64)
        if normalize:
            d = d / np.asarray([len(b) for b in tl], np.float64)
        return d.astype(np.float32)
    return _g.Op(fn, (hypothesis, truth), {}, name="edit_distance")


class ConfigProto(object):
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


class Session(object):
    def __init__(self, target="", graph=None, config=None):
        self.graph = graph or get_default_graph()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def run(self, fetches, feed_dict=None):
        run = _g._Run(feed_dict)
        single = not isinstance(fetches, (list, tuple))
        out = [_host(run.resolve(f)) for f in ([fetches] if single else fetches)]
        return out[0] if single else out


class _Summary(object):
    @staticmethod
    def scalar(name, tensor):
        return _g.Op(lambda v: (name, float(np.asarray(_host(v)))), (tensor,), {}, name="summary/" + name)

    @staticmethod
    def merge(inputs):
        return _g.Op(lambda vs: list(vs), (list(inputs),), {}, name="summary/merge")

    class FileWriter(object):
        def __init__(self, logdir, graph=None):
            self.logdir, self.events = logdir, []

        def add_summary(self, summary, global_step=None):
            self.events.append((global_step, summary))

        def flush(self):
            pass

        def close(self):
            pass


summary = _Summary


class _CheckpointState(object):
    def __init__(self, path):
        self.model_checkpoint_path = path


class _Train(object):
    class Saver(object):
        """Checkpoints = {TF variable name: array} in one .npz (+ a 'checkpoint' index file) and, beside it, the same
        variables as a TensorFlow checkpoint bundle (``<path>.index`` + ``<path>.data-00000-of-00001``,
        utils/io/tf_checkpoint.py) with TF's slot names for the optimizer state -- the format the reference's own
        ``tf.train.Saver`` reads and writes.  ``restore`` takes either: the .npz when it exists, else a TF bundle
        (e.g. one trained with the reference), matching variables by name (a leading scope prefix is ignored)."""

        # optimizer kind -> TF slot-name suffixes of (state0, state1)
        _SLOTS = {"adagrad": ("Adagrad", None), "adadelta": ("Adadelta", "Adadelta_1"), "adam": ("Adam", "Adam_1"),
                  "rmsprop": ("RMSProp", None), "momentum": ("Momentum", None), "nestrov": ("Momentum", None),
                  "sgd": (None, None)}

        def __init__(self, max_to_keep=None, var_list=None, write_tf_bundle=True):
            self.write_tf_bundle = write_tf_bundle

        def _bundle_arrays(self, sess):
            arrays = {}
            for m in sess.graph.models:
                opt = getattr(m, "optimizer", None)
                base = m.flat_params.data_ptr()
                for v in m.trainable_variables():
                    if v.name in arrays:
                        raise ValueError("two models of this session share the variable name %r" % v.name)
                    arrays[v.name] = v.tensor.detach().cpu().numpy()
                    if opt is None:
                        continue
                    o = (v.tensor.data_ptr() - base) // 4
                    for st, suffix in zip((opt.state0, opt.state1), self._SLOTS.get(opt.kind, (None, None))):
                        if st is not None and suffix is not None:
                            arrays["%s/%s" % (v.name, suffix)] = \
                                st[o:o + v.tensor.numel()].view(v.tensor.shape).detach().cpu().numpy()
                    if opt.kind == "rmsprop":          # TF keeps a (here always zero) momentum slot as well
                        arrays[v.name + "/RMSProp_1"] = np.zeros(tuple(v.tensor.shape), np.float32)
                if opt is not None:
                    arrays["global_step"] = np.asarray(opt.global_step, np.int32)
                    if opt.kind == "adam":             # TF: beta^t after t-1 updates, i.e. beta^(global_step + 1)
                        arrays["beta1_power"] = np.asarray(0.9 ** (opt.global_step + 1), np.float32)
                        arrays["beta2_power"] = np.asarray(0.999 ** (opt.global_step + 1), np.float32)
            return arrays

        def save(self, sess, save_path, global_step=None):
            """A TF Saver stores every global variable: the trainable ones, the optimizer slots and the step
            counters.  Same here (slot-style names under ``<model index>/optimizer/``); the container is an
            .npz keyed by TF variable names, NOT the TF checkpoint format."""
            path = save_path if global_step is None else "%s-%d" % (save_path, global_step)
            arrays = {}
            for i, m in enumerate(sess.graph.models):
                for v in m.trainable_variables():
                    arrays[v.name] = v.tensor.detach().cpu().numpy()
                arrays["%d/model_step" % i] = np.asarray(getattr(m, "_step", 0), np.int64)
                opt = getattr(m, "optimizer", None)
                if opt is not None:
                    pre = "%d/optimizer/" % i
                    arrays[pre + "name"] = np.asarray(opt.name)
                    arrays[pre + "global_step"] = np.asarray(opt.global_step, np.int64)
                    for k in ("state0", "state1"):
                        st = getattr(opt, k, None)
                        if st is not None:
                            arrays[pre + k] = st.detach().cpu().numpy()
            np.savez(path + ".npz", **arrays)
            if self.write_tf_bundle:
                from ..utils.io import tf_checkpoint
                tf_checkpoint.save_tf_checkpoint(path, self._bundle_arrays(sess), write_state=False)
            with open(os.path.join(os.path.dirname(path) or ".", "checkpoint"), "w") as f:
                f.write('model_checkpoint_path: "%s"\n' % path)
            return path

        def _restore_tf_bundle(self, sess, save_path):
            """variables (and, where the optimizer exists already, its slots) from a TensorFlow checkpoint bundle"""
            import torch
            from ..utils.io import tf_checkpoint
            data = tf_checkpoint.load_tf_checkpoint(save_path)

            def find(name):
                if name in data:
                    return data[name]
                hits = [k for k in data if k.endswith("/" + name)]
                if len(hits) == 1:
                    return data[hits[0]]
                if not hits:
                    raise KeyError("variable %r is not in the checkpoint %s" % (name, save_path))
                raise KeyError("variable %r is ambiguous in %s: %s" % (name, save_path, hits))
            for m in sess.graph.models:
                for v in m.trainable_variables():
                    a = find(v.name)
                    if tuple(a.shape) != tuple(v.tensor.shape):
                        raise ValueError("%s: checkpoint shape %s, model shape %s" % (v.name, a.shape,
                                                                                     tuple(v.tensor.shape)))
                    v.tensor.copy_(torch.as_tensor(a).to(v.tensor.device))
                m._params_version = getattr(m, "_params_version", 0) + 1
                opt = getattr(m, "optimizer", None)
                if opt is None:
                    continue
                base = m.flat_params.data_ptr()
                for st, suffix in zip((opt.state0, opt.state1), self._SLOTS.get(opt.kind, (None, None))):
                    if st is None or suffix is None:
                        continue
                    for v in m.trainable_variables():
                        try:
                            a = find("%s/%s" % (v.name, suffix))
                        except KeyError:
                            continue
                        o = (v.tensor.data_ptr() - base) // 4
                        st[o:o + v.tensor.numel()].copy_(torch.as_tensor(a).reshape(-1).to(st.device))
                if "global_step" in data:
                    opt.global_step = int(data["global_step"])

        def restore(self, sess, save_path):
            import torch
            if not os.path.isfile(save_path + ".npz"):
                from ..utils.io import tf_checkpoint
                if tf_checkpoint.is_tf_checkpoint(save_path):
                    return self._restore_tf_bundle(sess, save_path)
            data = np.load(save_path + ".npz")
            for i, m in enumerate(sess.graph.models):
                for v in m.trainable_variables():
                    v.tensor.copy_(torch.as_tensor(data[v.name]).to(v.tensor.device))
                if "%d/model_step" % i in data:
                    m._step = int(data["%d/model_step" % i])
                pre = "%d/optimizer/" % i
                if pre + "name" in data:
                    state = {k: data[pre + k] for k in ("state0", "state1", "global_step") if pre + k in data}
                    state["name"] = str(data[pre + "name"])
                    opt = getattr(m, "optimizer", None)
                    if opt is not None and opt.name == state["name"]:
                        opt.load_state(state)
                    else:                       # the optimizer is created by the first train(): hand it over then
                        m._restored_optimizer_state = state

    @staticmethod
    def get_checkpoint_state(checkpoint_dir):
        idx = os.path.join(checkpoint_dir, "checkpoint")
        if not os.path.isfile(idx):
            return None
        with open(idx) as f:
            line = f.readline()
        return _CheckpointState(line.split('"')[1])


train = _Train


class _Test(object):
    TestCase = unittest.TestCase

    @staticmethod
    def main():
        unittest.main()


test = _Test
