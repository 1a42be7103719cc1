"""Attention seq2seq / joint CTC-attention training step on the GPU (host mirrors of
models/attention/attention_seq2seq.py and joint_ctc_attention.py) vs the float64 torch oracle
(oracle/seq2seq.py, gradients from autograd): sequence-loss kernel, total loss, teacher-forced
logits, every parameter gradient (encoder through both the context path and the bridge /
final-state path), and a short adam trajectory.  fp32 tolerance 2e-4 relative on loss/logits
(north star 1e-3), gradients 1e-3 of each tensor's max."""
import numpy as np
import pytest
import torch

from oracle import seq2seq as os2s
from oracle import optim as oopt

pytestmark = pytest.mark.gpu


def make_batch(rng, B, T, D, V, Lmax):
    """V = raw class count; sos = V, eos = V + 1; labels [B, Lmax] = <SOS> chars <EOS> pad(<EOS>)"""
    x = rng.randn(B, T, D).astype(np.float32)
    seq = np.array([T] + [int(rng.randint(T // 2, T + 1)) for _ in range(B - 1)], np.int32)
    for b in range(B):
        x[b, seq[b]:] = 0
    lab_len = np.array([Lmax] + [int(rng.randint(3, Lmax + 1)) for _ in range(B - 1)], np.int32)
    labels = np.full((B, Lmax), V + 1, np.int32)
    ctc_labels = []
    for b in range(B):
        n = lab_len[b] - 2
        chars = rng.randint(0, V, n)
        labels[b, 0] = V
        labels[b, 1:1 + n] = chars
        ctc_labels.append([int(c) for c in chars[:max(1, min(n, seq[b] // 3))]])
    return x, seq, labels, lab_len, ctc_labels


def build(cuda, attention_type, joint=False, D=12, H=16, L=2, V=7, A=12, Hd=20, emb=8, precision="fp32", **kw):
    from tensorflow_end2end_speech_recognition_b200.models.attention.attention_seq2seq import AttentionSeq2Seq
    from tensorflow_end2end_speech_recognition_b200.models.attention.joint_ctc_attention import JointCTCAttention
    common = dict(input_size=D, encoder_type="blstm", encoder_num_units=H, encoder_num_layers=L,
                  encoder_num_proj=None, attention_type=attention_type, attention_dim=A, decoder_type="lstm",
                  decoder_num_units=Hd, decoder_num_layers=1, embedding_dim=emb, num_classes=V, sos_index=V,
                  eos_index=V + 1, max_decode_length=15, parameter_init=0.2, clip_grad_norm=5.0,
                  precision=precision, device=cuda, seed=5)
    common.update(kw)
    if joint:
        return JointCTCAttention(lambda_weight=0.4, **common)
    return AttentionSeq2Seq(**common)


def oracle_loss_and_grads(model, cfg, x, seq, labels, lab_len, ctc_labels=None):
    vs = {v.name: torch.tensor(v.tensor.cpu().numpy(), dtype=torch.float64, requires_grad=True)
          for v in model.trainable_variables()}
    out = os2s.seq2seq_loss(vs, cfg, torch.tensor(x, dtype=torch.float64), seq, labels, lab_len, ctc_labels)
    out["total_loss"].backward()
    grads = {n: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for n, v in vs.items()}
    return out, grads


def test_sequence_loss_kernel(cuda):
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(0)
    B, L, C = 5, 9, 13
    logits = rng.randn(B, L, C).astype(np.float32) * 2
    labels = rng.randint(0, C, (B, L + 1)).astype(np.int32)
    lens = np.array([9, 1, 4, 0, 7], np.int32)
    lt = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    ref = os2s.sequence_loss_t(lt / 2.0, labels[:, 1:], lens)
    ref.backward()
    loss, dl = ops.sequence_loss(torch.tensor(logits, device=cuda), torch.tensor(labels, device=cuda)[:, 1:],
                                 torch.tensor(lens, device=cuda), temperature=2.0)
    assert abs(float(loss) - float(ref.detach())) < 1e-5 * abs(float(ref.detach()))
    np.testing.assert_allclose(dl.cpu().numpy(), lt.grad.numpy(), rtol=1e-4, atol=1e-7)


def check_grads(model, g_ref, tol):
    for v in model.trainable_variables():
        g = g_ref[v.name]
        s = max(1e-4, np.abs(g).max())
        np.testing.assert_allclose(v.grad.cpu().numpy(), g, rtol=0, atol=tol * s, err_msg=v.name)


@pytest.mark.parametrize("attention_type", ["bahdanau_content", "hybrid", "location", "dot_product",
                                            "luong_dot", "luong_general", "luong_concat"])
def test_seq2seq_loss_and_grads(cuda, attention_type):
    rng = np.random.RandomState(1)
    B, T, D, V, Lmax = 5, 24, 12, 7, 9
    kw = dict(Hd=32) if attention_type == "luong_dot" else {}        # luong_dot needs Hd == 2H
    model = build(cuda, attention_type, **kw)
    x, seq, labels, lab_len, _ = make_batch(rng, B, T, D, V, Lmax)
    loss, logits, out_train, out_infer = model.compute_loss(x, labels, seq, lab_len, 1.0, 1.0, 1.0)
    model._backward()
    torch.cuda.synchronize()
    cfg = dict(num_layers=2, attention_type=attention_type)
    ref, g_ref = oracle_loss_and_grads(model, cfg, x, seq, labels, lab_len)
    assert abs(float(loss) - float(ref["total_loss"].detach())) <= 2e-4 * abs(float(ref["total_loss"].detach()))
    np.testing.assert_allclose(logits.cpu().numpy(), ref["decoder"]["logits"].detach().numpy(), rtol=2e-4, atol=2e-5)
    check_grads(model, g_ref, 1e-3)
    # the lazily evaluated inference decoder
    ids_train, ids_infer = model.decode(out_train, out_infer)
    assert ids_train.shape == (B, Lmax - 1) and ids_infer.shape[0] == B and ids_infer.shape[1] <= 15


def test_seq2seq_options(cuda):
    """sharpening, sigmoid smoothing, logits temperature, weight decay, no peephole"""
    rng = np.random.RandomState(2)
    B, T, D, V, Lmax = 4, 20, 12, 7, 8
    model = build(cuda, "hybrid", sharpening_factor=2.0, sigmoid_smoothing=True, logits_temperature=2.0,
                  weight_decay=1e-3, use_peephole=False)
    x, seq, labels, lab_len, _ = make_batch(rng, B, T, D, V, Lmax)
    loss, logits, _, _ = model.compute_loss(x, labels, seq, lab_len, 1.0, 1.0, 1.0)
    model._backward()
    cfg = dict(num_layers=2, attention_type="hybrid", use_peephole=False, sharpening_factor=2.0,
               sigmoid_smoothing=True, logits_temperature=2.0, weight_decay=1e-3)
    ref, g_ref = oracle_loss_and_grads(model, cfg, x, seq, labels, lab_len)
    assert abs(float(loss) - float(ref["total_loss"].detach())) <= 2e-4 * abs(float(ref["total_loss"].detach()))
    check_grads(model, g_ref, 1e-3)


@pytest.mark.parametrize("faithful", [False, True])
def test_joint_ctc_attention(cuda, faithful):
    rng = np.random.RandomState(3)
    B, T, D, V, Lmax = 4, 22, 12, 7, 8
    model = build(cuda, "bahdanau_content", joint=True, faithful_ctc_reshape=faithful)
    x, seq, labels, lab_len, ctc_labels = make_batch(rng, B, T, D, V, Lmax)
    loss, logits, ctc_logits, _, _ = model.compute_loss(x, labels, ctc_labels, seq, lab_len, 1.0, 1.0, 1.0)
    model._backward()
    cfg = dict(num_layers=2, attention_type="bahdanau_content", lambda_weight=0.4, ctc_faithful_reshape=faithful)
    ref, g_ref = oracle_loss_and_grads(model, cfg, x, seq, labels, lab_len, ctc_labels)
    assert abs(float(loss) - float(ref["total_loss"].detach())) <= 2e-4 * abs(float(ref["total_loss"].detach()))
    np.testing.assert_allclose(ctc_logits.cpu().numpy(), ref["ctc_logits"].detach().numpy(), rtol=2e-4, atol=2e-5)
    check_grads(model, g_ref, 1e-3)


def test_joint_label_longer_than_input_is_an_error(cuda):
    rng = np.random.RandomState(4)
    model = build(cuda, "bahdanau_content", joint=True)
    x, seq, labels, lab_len, ctc_labels = make_batch(rng, 3, 10, 12, 7, 6)
    ctc_labels[1] = [1] * 40
    with pytest.raises(RuntimeError):
        model.compute_loss(x, labels, ctc_labels, seq, lab_len, 1.0, 1.0, 1.0)


def test_adam_trajectory(cuda):
    rng = np.random.RandomState(5)
    B, T, D, V, Lmax = 4, 18, 12, 7, 7
    model = build(cuda, "bahdanau_content")
    names = [v.name for v in model.trainable_variables()]
    params = [v.tensor.cpu().numpy().astype(np.float64) for v in model.trainable_variables()]
    opt = oopt.Optimizer("adam", 1e-2)
    cfg = dict(num_layers=2, attention_type="bahdanau_content")
    for _ in range(3):
        x, seq, labels, lab_len, _ = make_batch(rng, B, T, D, V, Lmax)
        loss, _, _, _ = model.compute_loss(x, labels, seq, lab_len, 1.0, 1.0, 1.0)
        model.train(loss, "adam", 1e-2)
        vs = {n: torch.tensor(p, dtype=torch.float64, requires_grad=True) for n, p in zip(names, params)}
        out = os2s.seq2seq_loss(vs, cfg, torch.tensor(x, dtype=torch.float64), seq, labels, lab_len)
        out["total_loss"].backward()
        grads = [oopt.clip_by_norm(vs[n].grad.numpy() if vs[n].grad is not None else np.zeros(vs[n].shape), 5.0)
                 for n in names]
        assert abs(float(loss) - float(out["total_loss"])) <= 1e-3 * abs(float(out["total_loss"]))
        opt.step(params, grads)
    for v, p in zip(model.trainable_variables(), params):
        np.testing.assert_allclose(v.tensor.cpu().numpy(), p, rtol=0, atol=2e-3 * max(1e-2, np.abs(p).max()),
                                   err_msg=v.name)


def test_bf16_encoder_final_state_gradient(cuda):
    """tcgen05 recurrence path (H=128): the bridge gradient enters the persistent BPTT kernel
    through d_final_state.  bf16 tolerance."""
    rng = np.random.RandomState(6)
    B, T, D, V, Lmax = 6, 30, 16, 7, 8
    model = build(cuda, "bahdanau_content", D=D, H=128, L=1, Hd=32, precision="bf16")
    x, seq, labels, lab_len, _ = make_batch(rng, B, T, D, V, Lmax)
    loss, logits, _, _ = model.compute_loss(x, labels, seq, lab_len, 1.0, 1.0, 1.0)
    model._backward()
    cfg = dict(num_layers=1, attention_type="bahdanau_content")
    ref, g_ref = oracle_loss_and_grads(model, cfg, x, seq, labels, lab_len)
    assert abs(float(loss) - float(ref["total_loss"])) <= 3e-2 * abs(float(ref["total_loss"]))
    check_grads(model, g_ref, 0.1)


@pytest.mark.parametrize("attention_type", ["bahdanau_content", "luong_general"])
def test_decoder_and_embedding_dropout(cuda, attention_type):
    """keep_prob_decoder (DropoutWrapper on the cell output) and keep_prob_embedding (dropout on the embedded
    labels): same counter-hash masks applied in the oracle; loss and every gradient."""
    from tests.util_dropout import dropout_mask
    rng = np.random.RandomState(8)
    B, T, D, V, Lmax, Hd, emb = 4, 20, 12, 7, 8, 20, 8
    kd, ke = 0.8, 0.6
    model = build(cuda, attention_type)
    x, seq, labels, lab_len, _ = make_batch(rng, B, T, D, V, Lmax)
    loss, logits, _, _ = model.compute_loss(x, labels, seq, lab_len, 1.0, kd, ke)
    seed = model._step * 7 + 3
    model._backward()
    L = Lmax - 1
    dec_mask = dropout_mask(seed, L * B * Hd, kd).reshape(L, B, Hd)
    emb_mask = dropout_mask(seed + 1, B * Lmax * emb, ke).reshape(B, Lmax, emb)
    cfg = dict(num_layers=2, attention_type=attention_type, dec_mask=dec_mask, keep_prob_decoder=kd,
               emb_mask=emb_mask, keep_prob_embedding=ke)
    ref, g_ref = oracle_loss_and_grads(model, cfg, x, seq, labels, lab_len)
    assert abs(float(loss) - float(ref["total_loss"].detach())) <= 2e-4 * abs(float(ref["total_loss"].detach()))
    np.testing.assert_allclose(logits.cpu().numpy(), ref["decoder"]["logits"].detach().numpy(), rtol=2e-4, atol=2e-5)
    check_grads(model, g_ref, 1e-3)
    # evaluation mode ignores the dropout rates
    l_eval, _, _, _ = model.compute_loss(x, labels, seq, lab_len, 1.0, kd, ke, is_training=False)
    l_ref, _, _, _ = model.compute_loss(x, labels, seq, lab_len, 1.0, 1.0, 1.0, is_training=False)
    assert abs(float(l_eval) - float(l_ref)) < 1e-6


@pytest.mark.parametrize("attention_type", ["hybrid", "location"])
def test_training_with_previous_attention_weights(cuda, attention_type):
    """feed_previous_attention=True: the location term sees the previous step's weights (the behaviour the
    reference intends but never reaches).  The gradient then also flows alpha_t -> conv ->
    alpha_{t-1} across decoder steps and into the conv filter and W_filter."""
    rng = np.random.RandomState(9)
    B, T, D, V, Lmax = 4, 22, 12, 7, 8
    model = build(cuda, attention_type, feed_previous_attention=True)
    # a wider filter init than truncated_normal(0.2)/0.1 so that its gradient is not negligible
    x, seq, labels, lab_len, _ = make_batch(rng, B, T, D, V, Lmax)
    loss, logits, _, _ = model.compute_loss(x, labels, seq, lab_len, 1.0, 1.0, 1.0)
    model._backward()
    torch.cuda.synchronize()
    cfg = dict(num_layers=2, attention_type=attention_type, feed_previous_attention=True)
    ref, g_ref = oracle_loss_and_grads(model, cfg, x, seq, labels, lab_len)
    assert abs(float(loss) - float(ref["total_loss"].detach())) <= 2e-4 * abs(float(ref["total_loss"].detach()))
    np.testing.assert_allclose(logits.cpu().numpy(), ref["decoder"]["logits"].detach().numpy(), rtol=2e-4, atol=2e-5)
    check_grads(model, g_ref, 1e-3)
    att = "decoder/attention_decoder/attention_layer/"
    assert np.abs(g_ref[att + "filter"]).max() > 0 and np.abs(g_ref[att + "W_filter/weights"]).max() > 0
