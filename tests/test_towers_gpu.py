"""The reference's in-graph multi-tower step (examples/librispeech/training/train_ctc.py:82-147): per tower
``compute_loss`` -> ``optimizer.compute_gradients`` -> ``model._clip_gradients``; then ``average_gradients`` ->
``optimizer.apply_gradients`` -- eager and through compat.tf placeholders -- against the oracle's
clip-then-mean-then-update (oracle/model.py, oracle/optim.py)."""
import numpy as np
import pytest
import torch

from oracle import model as omodel
from oracle import optim as oopt

pytestmark = pytest.mark.gpu

B, T, D, H, L, C = 6, 30, 16, 32, 2, 9


def _data(seed):
    rng = np.random.RandomState(seed)
    x = rng.randn(B, T, D).astype(np.float32)
    seq = np.array([T] + [int(rng.randint(T // 2, T + 1)) for _ in range(B - 1)], np.int32)
    for b in range(B):
        x[b, seq[b]:] = 0
    labels = [list(rng.randint(0, C, size=int(rng.randint(2, 8)))) for _ in range(B)]
    return x, seq, labels


def _oracle_step(vs, shards, clip, lr):
    tr = omodel.OracleTrainer(vs, L, optimizer="sgd", learning_rate=lr, clip_grad_norm=None)
    towers, losses = [], []
    for x, seq, labels in shards:
        l, _, g = tr.loss_and_grads(x, seq, labels)
        towers.append([oopt.clip_by_norm(gi, clip) for gi in g])
        losses.append(l)
    mean = oopt.average_gradients(towers)
    return [p - lr * g for p, g in zip(tr.params, mean)], losses


def _model(cuda, clip):
    from tensorflow_end2end_speech_recognition_b200.models.ctc.ctc import CTC
    return CTC(encoder_type="blstm", input_size=D, num_units=H, num_layers=L, num_classes=C, parameter_init=0.1,
               clip_grad_norm=clip, precision="fp32", device=cuda, seed=4)


@pytest.mark.parametrize("ntower", [2, 3])
def test_eager_tower_flow(cuda, ntower):
    from tensorflow_end2end_speech_recognition_b200.utils.training.multi_gpu import average_gradients
    clip, lr = 0.5, 0.05
    model = _model(cuda, clip)
    vs = {v.name: v.tensor.cpu().numpy() for v in model.trainable_variables()}
    shards = [_data(10 + i) for i in range(ntower)]
    optimizer = model._set_optimizer("sgd", lr)
    total = []
    # all forwards first, then the backward passes: every loss carries its own context
    losses = [model.compute_loss(x, labels, seq, keep_prob=1.0)[0] for x, seq, labels in shards]
    for loss in losses:
        gv = optimizer.compute_gradients(loss)
        total.append(model._clip_gradients(gv))
    avg = average_gradients(total)
    optimizer.apply_gradients(avg, global_step=None)
    torch.cuda.synchronize()
    want, l_ref = _oracle_step(vs, shards, clip, lr)
    for l, r in zip(losses, l_ref):
        assert abs(float(l) - r) <= 2e-4 * abs(r)
    for v, w in zip(model.trainable_variables(), want):
        np.testing.assert_allclose(v.tensor.cpu().numpy(), w, rtol=2e-3, atol=2e-5, err_msg=v.name)


@pytest.mark.parametrize("loss_first", [False, True])
def test_graph_tower_flow(cuda, loss_first):
    from tensorflow_end2end_speech_recognition_b200 import compat
    tf = compat.install()
    try:
        from utils.io.labels.sparsetensor import list2sparsetensor
        from utils.training.multi_gpu import average_gradients
        clip, lr, ntower = 0.5, 0.05, 2
        model = _model(cuda, clip)
        vs = {v.name: v.tensor.cpu().numpy() for v in model.trainable_variables()}
        shards = [_data(20 + i) for i in range(ntower)]
        with tf.Graph().as_default(), tf.device("/cpu:0"):
            global_step = tf.Variable(0, name="global_step", trainable=False)
            learning_rate_pl = tf.placeholder(tf.float32, name="learning_rate")
            optimizer = model._set_optimizer("sgd", learning_rate_pl)
            total_grads_and_vars, total_losses, decode_ops, ler_ops = [], [], [], []
            with tf.variable_scope(tf.get_variable_scope()):
                for i_gpu in range(ntower):
                    with tf.device("/gpu:%d" % i_gpu), tf.name_scope("tower_gpu%d" % i_gpu) as scope:
                        model.create_placeholders()
                        tower_loss, tower_logits = model.compute_loss(
                            model.inputs_pl_list[i_gpu], model.labels_pl_list[i_gpu],
                            model.inputs_seq_len_pl_list[i_gpu], model.keep_prob_pl_list[i_gpu], scope)
                        tower_loss = tf.expand_dims(tower_loss, axis=0)
                        total_losses.append(tower_loss)
                        tf.get_variable_scope().reuse_variables()
                        tower_grads_and_vars = optimizer.compute_gradients(tower_loss)
                        tower_grads_and_vars = model._clip_gradients(tower_grads_and_vars)
                        total_grads_and_vars.append(tower_grads_and_vars)
                        decode_op_tower = model.decoder(tower_logits, model.inputs_seq_len_pl_list[i_gpu], beam_width=1)
                        decode_ops.append(decode_op_tower)
                        ler_op_tower = model.compute_ler(decode_op_tower, model.labels_pl_list[i_gpu])
                        ler_ops.append(tf.expand_dims(ler_op_tower, axis=0))
            loss_op = tf.reduce_mean(tf.concat(axis=0, values=total_losses), axis=0)
            ler_op = tf.reduce_mean(tf.concat(axis=0, values=ler_ops), axis=0)
            average_grads_and_vars = average_gradients(total_grads_and_vars)
            train_op = optimizer.apply_gradients(average_grads_and_vars, global_step=global_step)
            feed = {learning_rate_pl: lr}
            for i_gpu, (x, seq, labels) in enumerate(shards):
                pad = np.full((B, 8), -1, np.int32)
                for b, l in enumerate(labels):
                    pad[b, :len(l)] = l
                feed[model.inputs_pl_list[i_gpu]] = x
                feed[model.labels_pl_list[i_gpu]] = list2sparsetensor(pad, padded_value=-1)
                feed[model.inputs_seq_len_pl_list[i_gpu]] = seq
                feed[model.keep_prob_pl_list[i_gpu]] = 1.0
            with tf.Session(config=tf.ConfigProto(allow_soft_placement=True, log_device_placement=False)) as sess:
                sess.run(tf.global_variables_initializer())
                if loss_first:
                    loss_val, _ = sess.run([loss_op, train_op], feed_dict=feed)
                else:
                    _, loss_val = sess.run([train_op, loss_op], feed_dict=feed)
                want, l_ref = _oracle_step(vs, shards, clip, lr)
                assert abs(float(loss_val) - np.mean(l_ref)) <= 2e-4 * abs(np.mean(l_ref))
                for v, w in zip(model.trainable_variables(), want):
                    np.testing.assert_allclose(v.tensor.cpu().numpy(), w, rtol=2e-3, atol=2e-5, err_msg=v.name)
                ler = sess.run(ler_op, feed_dict=feed)
                assert 0.0 <= float(ler) <= 5.0
                assert optimizer.global_step == 1
    finally:
        compat.uninstall()
