"""The reference's graph-style flow (models/test/test_ctc.py:72-230, test_attention.py) through
compat.tf on the B200 models: placeholders -> op handles -> sess.run with feed_dict, training
until the synthetic batch is fitted, decode + LER, parameter count, checkpoint round trip."""
import os

import numpy as np
import torch
import pytest

pytestmark = pytest.mark.gpu


def test_ctc_graph_style_training(cuda, tmp_path):
    from tensorflow_end2end_speech_recognition_b200 import compat
    tf = compat.install()
    try:
        from models.ctc.ctc import CTC
        from utils.io.labels.sparsetensor import list2sparsetensor, sparsetensor2list
        rng = np.random.RandomState(0)
        B, T, D, C = 2, 40, 24, 27
        inputs = rng.randn(B, T, D).astype(np.float32)
        inputs_seq_len = np.array([T, T - 6], np.int32)
        labels = np.full((B, 8), -1, np.int32)
        # no adjacent repeats: decode_op (beam_width=20) is tf.nn.ctc_beam_search_decoder, whose default
        # merge_repeated=True collapses "a a" to "a" in the emitted path (ctc.py:344-346)
        labels[0, :8] = [3, 7, 1, 19, 4, 11, 26, 0]
        labels[1, :5] = [5, 2, 9, 2, 14]
        with tf.Graph().as_default():
            model = CTC(encoder_type="blstm", input_size=D, splice=1, num_stack=1, num_units=32, num_layers=2,
                        num_classes=C, lstm_impl="LSTMBlockCell", parameter_init=0.1, clip_grad_norm=5.0,
                        clip_activation=50, weight_decay=1e-10, time_major=True, device=cuda)
            model.create_placeholders()
            learning_rate_pl = tf.placeholder(tf.float32, name="learning_rate")
            loss_op, logits = model.compute_loss(model.inputs_pl_list[0], model.labels_pl_list[0],
                                                 model.inputs_seq_len_pl_list[0], model.keep_prob_pl_list[0])
            train_op = model.train(loss_op, optimizer="adam", learning_rate=learning_rate_pl)
            decode_op = model.decoder(logits, model.inputs_seq_len_pl_list[0], beam_width=20)
            ler_op = model.compute_ler(decode_op, model.labels_pl_list[0])
            posteriors_op = model.posteriors(logits)
            init_op = tf.global_variables_initializer()
            saver = tf.train.Saver(max_to_keep=None)
            n_params = sum(int(np.prod([d.value for d in v.get_shape()])) for v in tf.trainable_variables())
            assert n_params == sum(v.tensor.numel() for v in model.trainable_variables())
            feed_dict = {model.inputs_pl_list[0]: inputs,
                         model.labels_pl_list[0]: list2sparsetensor(labels, padded_value=-1),
                         model.inputs_seq_len_pl_list[0]: inputs_seq_len,
                         model.keep_prob_pl_list[0]: 1.0,
                         learning_rate_pl: 5e-3}
            with tf.Session() as sess:
                sess.run(init_op)
                first = None
                for step in range(500):
                    _, loss_train = sess.run([train_op, loss_op], feed_dict=feed_dict)
                    first = loss_train if first is None else first
                    if (step + 1) % 20 == 0:
                        ler_train = sess.run(ler_op, feed_dict=feed_dict)
                        if ler_train == 0:
                            break
                assert loss_train < 0.2 * first and ler_train == 0
                labels_pred_st = sess.run(decode_op, feed_dict=feed_dict)
                pred = sparsetensor2list(labels_pred_st, batch_size=B)
                assert list(pred[0]) == list(labels[0, :8]) and list(pred[1]) == list(labels[1, :5])
                # evaluation feeds carry no labels (examples/timit/metrics/ctc.py:72-81): decode / posteriors
                # depend on (inputs, inputs_seq_len, keep_prob) only
                feed_eval = {k: v for k, v in feed_dict.items() if k is not model.labels_pl_list[0]}
                pred2 = sparsetensor2list(sess.run(decode_op, feed_dict=feed_eval), batch_size=B)
                assert [list(p) for p in pred2] == [list(p) for p in pred]
                with pytest.raises(ValueError):
                    sess.run(loss_op, feed_dict=feed_eval)
                post = sess.run(posteriors_op, feed_dict=feed_eval)
                assert post.shape == (B * T, C + 1) and np.allclose(post.sum(-1), 1, atol=1e-5)
                # checkpoint round trip
                path = saver.save(sess, os.path.join(str(tmp_path), "model.ckpt"), global_step=2)
                before = sess.run(loss_op, feed_dict=feed_dict)
                model.flat_params.mul_(0.5)
                assert sess.run(loss_op, feed_dict=feed_dict) != before
                ckpt = tf.train.get_checkpoint_state(str(tmp_path))
                saver.restore(sess, ckpt.model_checkpoint_path)
                assert path == ckpt.model_checkpoint_path
                assert abs(sess.run(loss_op, feed_dict=feed_dict) - before) < 1e-6
                # the same checkpoint as a TensorFlow bundle (<path>.index / .data-00000-of-00001), TF slot names
                from tensorflow_end2end_speech_recognition_b200.utils.io import tf_checkpoint
                bundle = tf_checkpoint.load_tf_checkpoint(path)
                name0 = model.trainable_variables()[0].name
                assert name0 in bundle and name0 + "/Adam" in bundle and name0 + "/Adam_1" in bundle   # adam: m, v slots
                assert "beta1_power" in bundle and int(bundle["global_step"]) == model.optimizer.global_step
                os.remove(path + ".npz")
                state_before, step_before = model.optimizer.state0.clone(), model.optimizer.global_step
                state1_before = model.optimizer.state1.clone()
                model.flat_params.mul_(0.5)
                model.optimizer.state0.add_(1.0)
                model.optimizer.state1.mul_(3.0)
                model.optimizer.global_step = 0
                saver.restore(sess, path)                      # no .npz any more: the TF bundle is read
                assert abs(sess.run(loss_op, feed_dict=feed_dict) - before) < 1e-6
                assert torch.equal(model.optimizer.state0, state_before) and model.optimizer.global_step == step_before
                assert torch.equal(model.optimizer.state1, state1_before)
                # a bundle written under a scope prefix (as a reference-trained graph may name its variables)
                pre = os.path.join(str(tmp_path), "scoped.ckpt")
                tf_checkpoint.save_tf_checkpoint(pre, {"ctc_model/" + k: v for k, v in bundle.items()})
                model.flat_params.mul_(0.25)
                saver.restore(sess, pre)
                assert abs(sess.run(loss_op, feed_dict=feed_dict) - before) < 1e-6
    finally:
        compat.uninstall()


def test_attention_graph_style_training(cuda):
    from tensorflow_end2end_speech_recognition_b200 import compat
    tf = compat.install()
    try:
        from models.attention.attention_seq2seq import AttentionSeq2Seq
        rng = np.random.RandomState(1)
        B, T, D, V = 2, 30, 24, 26
        inputs = rng.randn(B, T, D).astype(np.float32)
        inputs_seq_len = np.array([T, T - 4], np.int32)
        sos, eos = V, V + 1
        labels = np.full((B, 9), eos, np.int32)
        labels[:, 0] = sos
        labels[0, 1:8] = rng.randint(0, V, 7)
        labels[1, 1:5] = rng.randint(0, V, 4)
        labels_seq_len = np.array([9, 6], np.int32)
        with tf.Graph().as_default():
            model = AttentionSeq2Seq(input_size=D, encoder_type="blstm", encoder_num_units=32, encoder_num_layers=2,
                                     encoder_num_proj=None, attention_type="hybrid", attention_dim=16,
                                     decoder_type="lstm", decoder_num_units=32, decoder_num_layers=1,
                                     embedding_dim=8, num_classes=V, sos_index=sos, eos_index=eos,
                                     max_decode_length=20, parameter_init=0.1, clip_grad_norm=5.0, device=cuda)
            model.create_placeholders()
            lr_pl = tf.placeholder(tf.float32, name="learning_rate")
            loss_op, logits, out_train, out_infer = model.compute_loss(
                model.inputs_pl_list[0], model.labels_pl_list[0], model.inputs_seq_len_pl_list[0],
                model.labels_seq_len_pl_list[0], model.keep_prob_encoder_pl_list[0],
                model.keep_prob_decoder_pl_list[0], model.keep_prob_embedding_pl_list[0])
            train_op = model.train(loss_op, optimizer="adam", learning_rate=lr_pl)
            decode_op_train, decode_op_infer = model.decode(out_train, out_infer)
            ler_op = model.compute_ler(model.labels_st_true_pl, model.labels_st_pred_pl)
            feed = {model.inputs_pl_list[0]: inputs, model.labels_pl_list[0]: labels,
                    model.inputs_seq_len_pl_list[0]: inputs_seq_len, model.labels_seq_len_pl_list[0]: labels_seq_len,
                    model.keep_prob_encoder_pl_list[0]: 1.0, model.keep_prob_decoder_pl_list[0]: 1.0,
                    model.keep_prob_embedding_pl_list[0]: 1.0, lr_pl: 1e-2}
            with tf.Session() as sess:
                sess.run(tf.global_variables_initializer())
                first = None
                for step in range(200):
                    _, loss = sess.run([train_op, loss_op], feed_dict=feed)
                    first = loss if first is None else first
                    if loss < 0.02:
                        break
                assert loss < 0.1 * first
                ids_train, ids_infer = sess.run([decode_op_train, decode_op_infer], feed_dict=feed)
                assert list(ids_train[0, :8]) == list(labels[0, 1:9])
                assert list(ids_infer[0, :8]) == list(labels[0, 1:9])       # greedy decode reproduces the labels
                from utils.io.labels.sparsetensor import list2sparsetensor
                t = list2sparsetensor([list(r) for r in labels[:, 1:]], padded_value=eos)
                p = list2sparsetensor([list(r) + [eos] for r in ids_infer], padded_value=eos)
                assert sess.run(ler_op, feed_dict={model.labels_st_true_pl: t, model.labels_st_pred_pl: p}) < 0.15
    finally:
        compat.uninstall()
