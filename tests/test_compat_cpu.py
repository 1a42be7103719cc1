"""compat.graph / compat.tf: the lazy-handle + Session.run layer the reference's graph-style
driver scripts need, checked on CPU with a stand-in model (no CUDA involved)."""
import numpy as np
import pytest

from tensorflow_end2end_speech_recognition_b200 import compat
from tensorflow_end2end_speech_recognition_b200.compat import graph as G
from tensorflow_end2end_speech_recognition_b200.compat import tf


class Toy(object):
    def __init__(self):
        self.calls = []
        self.w = 1.0

    @G.graph_op(n_out=2, name="compute_loss")
    def compute_loss(self, x, scale):
        self.calls.append("loss")
        return float(np.sum(x)) * scale * self.w, np.asarray(x) * 2

    @G.graph_op(name="train")
    def train(self, loss, optimizer, learning_rate):
        self.calls.append("train")
        self.w -= learning_rate
        return None

    @G.graph_op(name="decoder")
    def decoder(self, logits):
        self.calls.append("decode")
        return np.asarray(logits).argmax()


def test_eager_when_no_handles():
    m = Toy()
    loss, logits = m.compute_loss(np.ones(3), 2.0)
    assert loss == 6.0 and np.all(logits == 2)


def test_session_runs_each_op_once_and_in_dependency_order():
    m = Toy()
    with tf.Graph().as_default():
        x, s, lr = tf.placeholder(tf.float32, name="x"), tf.placeholder(tf.float32), tf.placeholder(tf.float32)
        loss_op, logits = m.compute_loss(x, s)
        train_op = m.train(loss_op, "sgd", lr)
        decode_op = m.decoder(logits)
        with tf.Session() as sess:
            sess.run(tf.global_variables_initializer())
            feed = {x: np.array([1.0, 5.0, 2.0]), s: 1.0, lr: 0.5}
            _, loss = sess.run([train_op, loss_op], feed_dict=feed)
            assert loss == 8.0 and m.calls == ["loss", "train"]          # loss evaluated once, before the update
            assert sess.run(loss_op, feed_dict=feed) == 4.0              # w moved 1.0 -> 0.5
            m.calls.clear()
            assert sess.run(decode_op, feed_dict=feed) == 1
            assert m.calls == ["loss", "decode"]
            with pytest.raises(ValueError):
                sess.run(loss_op, feed_dict={x: np.zeros(2)})             # scale placeholder not fed


def test_sparse_placeholder_feeds_and_edit_distance():
    with tf.Graph().as_default():
        hyp = tf.SparseTensor(tf.placeholder(tf.int64), tf.placeholder(tf.int32), tf.placeholder(tf.int64))
        ref = tf.SparseTensor(tf.placeholder(tf.int64), tf.placeholder(tf.int32), tf.placeholder(tf.int64))
        ler = tf.reduce_mean(tf.edit_distance(hyp, ref, normalize=True))
        h = [np.array([[0, 0], [0, 1], [1, 0]], np.int64), np.array([1, 2, 3], np.int32), np.array([2, 2], np.int64)]
        r = [np.array([[0, 0], [0, 1], [1, 0], [1, 1]], np.int64), np.array([1, 3, 3, 4], np.int32),
             np.array([2, 2], np.int64)]
        with tf.Session() as sess:
            assert abs(sess.run(ler, feed_dict={hyp: h, ref: r}) - 0.5) < 1e-7


def test_install_aliases_reference_imports():
    mod = compat.install()
    try:
        import tensorflow
        assert tensorflow is mod and tensorflow.__version__ == "1.2.0"
        from utils.io.labels.sparsetensor import list2sparsetensor
        st = list2sparsetensor([[1, 2, -1], [3, -1, -1]], padded_value=-1)
        assert st[1].tolist() == [1, 2, 3]
        from models.encoders.load_encoder import load
        assert load("blstm").__name__ == "BLSTMEncoder"
    finally:
        compat.uninstall()


def test_summary_scope_and_misc_symbols():
    with tf.Graph().as_default():
        with tf.device("/gpu:0"), tf.name_scope("tower_0"), tf.variable_scope("x") as vs:
            tf.get_variable_scope().reuse_variables()
            c = tf.Variable(0, name="global_step", trainable=False)
        merged = tf.summary.merge([tf.summary.scalar("loss", c)])
        with tf.Session(config=tf.ConfigProto(allow_soft_placement=True)) as sess:
            w = tf.summary.FileWriter("/tmp/x", sess.graph)
            w.add_summary(sess.run(merged), 1)
            w.flush()
            assert w.events[0][1] == [("loss", 0.0)]
            assert sess.run(tf.reduce_mean(tf.concat([tf.expand_dims(c, 0), tf.expand_dims(c, 0)], 0))) == 0


def test_tf_surface_listed_in_integration_md_exists():
    """every tf.* symbol INTEGRATION.md promises resolves on the shim"""
    names = ["Graph", "reset_default_graph", "device", "name_scope", "variable_scope", "get_variable_scope",
             "placeholder", "float32", "int32", "int64", "SparseTensor", "SparseTensorValue", "Variable",
             "expand_dims", "concat", "reduce_mean", "edit_distance", "global_variables_initializer",
             "trainable_variables", "Session", "ConfigProto"]
    for n in names:
        assert hasattr(tf, n), n
    for n in ("scalar", "merge", "FileWriter"):
        assert hasattr(tf.summary, n)
    for n in ("Saver", "get_checkpoint_state"):
        assert hasattr(tf.train, n)
    assert hasattr(tf.test, "TestCase") and hasattr(tf.test, "main")
