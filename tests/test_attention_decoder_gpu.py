"""Attention decoder loop (bridge -> LSTM step -> attention -> attentional vector -> logits
-> arg-max -> helper) on the GPU vs oracle/attention_decoder.py.

The GPU ids are fed to the oracle as ``forced_ids`` so that a float near-tie in one arg-max
cannot derail the comparison of the remaining steps; wherever the oracle's top-2 margin is
above 1e-4 the ids themselves must agree.  Tolerance on logits / attention weights 2e-4
(north-star 1e-3 rtol fp32)."""
from collections import namedtuple

import numpy as np
import pytest
import torch

from oracle import attention_decoder as odec

pytestmark = pytest.mark.gpu

EncoderOutput = namedtuple("EncoderOutput", ["outputs", "final_state", "seq_len"])


def build(cuda, atype, B, T, H_enc, Hd, A, emb, C, peephole, seed, max_len=30, feed_prev=False,
          time_major=False):
    from tensorflow_end2end_speech_recognition_b200.models.attention.bridge import InitialStateBridge
    from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.attention_decoder import (
        AttentionDecoder, LSTMBlockCell)
    from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.attention_layer import AttentionLayer
    rng = np.random.RandomState(seed)
    E = 2 * H_enc
    enc = rng.randn(B, T, E).astype(np.float32) * 0.5
    lens = np.array([T] + [int(rng.randint(1, T + 1)) for _ in range(B - 1)], np.int32)
    for b in range(B):
        enc[b, lens[b]:] = 0
    fs = [rng.randn(B, H_enc).astype(np.float32) * 0.5 for _ in range(4)]
    t = lambda x: torch.tensor(x, device=cuda)
    enc_out = EncoderOutput(t(enc), ((t(fs[0]), t(fs[1])), (t(fs[2]), t(fs[3]))), t(lens))
    layer = AttentionLayer(atype, A, 0.3, 1.0, False)
    layer.create_variables(E, Hd, rng, cuda)
    cell = LSTMBlockCell(Hd, forget_bias=1.0, use_peephole=peephole)
    dec = AttentionDecoder(cell, 0.3, max_len, C, enc_out.outputs, enc_out.seq_len, layer,
                           time_major=time_major, feed_previous_attention=feed_prev, poll_every=4)
    dec.create_variables(emb, rng, cuda)
    dec.cell_variables["bias"] += t(rng.randn(4 * Hd).astype(np.float32) * 0.1)
    dec.variables["output_layer/biases"] += t(rng.randn(C).astype(np.float32) * 0.1)
    bridge = InitialStateBridge(enc_out, cell.state_size, 0.3)
    bridge.create_variables(4 * H_enc, rng, cuda)
    bridge.variables["bridge/biases"] += t(rng.randn(2 * Hd).astype(np.float32) * 0.1)
    embedding = t(rng.uniform(-0.5, 0.5, (C, emb)).astype(np.float32))
    n = lambda d: {k: v.cpu().numpy() for k, v in d.items()}
    p = dict(n(bridge.variables))
    p.update(n(dec.variables))
    p["cell"] = n(dec.cell_variables)
    p["attention"] = n(layer.variables)
    p["W_embedding"] = embedding.cpu().numpy()
    return dec, bridge, embedding, enc_out, p, (enc, lens, fs)


def check(out, ref, rtol=2e-4):
    g = lambda x: x.cpu().numpy()
    L = ref["logits"].shape[1]
    assert out.logits.shape[1] == L
    np.testing.assert_allclose(g(out.logits), ref["logits"], rtol=rtol, atol=2e-5)
    np.testing.assert_allclose(g(out.decoder_output), ref["decoder_output"], rtol=rtol, atol=2e-5)
    np.testing.assert_allclose(g(out.attention_weights), ref["attention_weights"], rtol=rtol, atol=1e-6)
    np.testing.assert_allclose(g(out.context_vector), ref["context_vector"], rtol=rtol, atol=2e-5)
    srt = np.sort(ref["logits"], axis=-1)
    clear = (srt[..., -1] - srt[..., -2]) > 1e-4
    assert np.array_equal(g(out.predicted_ids)[clear], ref["predicted_ids"][clear])


def test_bridge(cuda):
    dec, bridge, emb, enc_out, p, (enc, lens, fs) = build(cuda, "bahdanau_content", 5, 20, 16, 24, 16, 8, 7, True, 1)
    st = bridge()
    c0, h0 = odec.bridge_initial_state(((fs[0], fs[1]), (fs[2], fs[3])), p)
    np.testing.assert_allclose(st.c.cpu().numpy(), c0, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.h.cpu().numpy(), h0, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("atype", ["bahdanau_content", "hybrid", "luong_general", "dot_product"])
@pytest.mark.parametrize("peephole", [False, True])
def test_greedy_decode(cuda, atype, peephole):
    from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.helpers import GreedyEmbeddingHelper
    B, T, H_enc, Hd, A, emb, C = 6, 40, 16, 32, 24, 12, 8
    sos, eos = C - 2, C - 1
    dec, bridge, embedding, enc_out, p, (enc, lens, fs) = build(cuda, atype, B, T, H_enc, Hd, A, emb, C, peephole, 3)
    helper = GreedyEmbeddingHelper(embedding, torch.full((B,), sos, dtype=torch.int32, device=cuda), eos)
    st = bridge()
    out, final = dec(st, helper)
    torch.cuda.synchronize()
    ids = out.predicted_ids.cpu().numpy()
    forced = np.zeros((B, 30), np.int64)
    forced[:, :ids.shape[1]] = ids
    # rows that finished keep feeding <EOS> in the oracle (their outputs are imputed anyway)
    ref = odec.decode(p, atype, enc, lens, (st.c.cpu().numpy(), st.h.cpu().numpy()), sos=sos, eos=eos,
                      max_decode_length=30, forced_ids=_forced_with_eos(ids, eos, 30))
    check(out, ref)
    np.testing.assert_allclose(final.c.cpu().numpy(), ref["final_state"][0], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(final.h.cpu().numpy(), ref["final_state"][1], rtol=2e-4, atol=2e-5)


def _forced_with_eos(ids, eos, L):
    B, n = ids.shape
    out = np.full((B, L), eos, np.int64)
    for b in range(B):
        hit = np.where(ids[b] == eos)[0]
        k = hit[0] + 1 if len(hit) else n
        out[b, :k] = ids[b, :k]
    return out


def test_greedy_stops_at_eos_and_imputes(cuda):
    """Force <EOS> early through the output bias: loop length = 1, then with a bias that never
    selects <EOS> the loop runs to max_decode_length."""
    from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.helpers import GreedyEmbeddingHelper
    B, C = 4, 6
    dec, bridge, embedding, enc_out, p, _ = build(cuda, "luong_general", B, 12, 8, 16, 16, 8, C, False, 5, max_len=9)
    helper = GreedyEmbeddingHelper(embedding, torch.full((B,), C - 2, dtype=torch.int32, device=cuda), C - 1)
    dec.variables["output_layer/biases"][C - 1] = 50.0
    out, _ = dec(bridge(), helper)
    assert out.logits.shape[1] == 1 and bool((out.predicted_ids == C - 1).all())
    dec.variables["output_layer/biases"][C - 1] = -50.0
    out, _ = dec(bridge(), helper)
    assert out.logits.shape[1] == 9
    assert not bool((out.predicted_ids == C - 1).any())


@pytest.mark.parametrize("feed_prev", [False, True])
def test_teacher_forced(cuda, feed_prev):
    from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.helpers import TrainingHelper
    B, T, H_enc, Hd, A, emb, C = 5, 33, 16, 32, 16, 10, 9
    dec, bridge, embedding, enc_out, p, (enc, lens, fs) = build(cuda, "hybrid", B, T, H_enc, Hd, A, emb, C, True, 11,
                                                                 feed_prev=feed_prev, time_major=True)
    rng = np.random.RandomState(2)
    T_out = 12
    lab_len = np.array([T_out, 3, 7, 2, 9], np.int32)            # incl. <SOS>, <EOS>
    labels = rng.randint(0, C - 2, (B, T_out)).astype(np.int32)
    labels[:, 0] = C - 2
    for b in range(B):
        labels[b, lab_len[b] - 1:] = C - 1
    helper = TrainingHelper(embedding, torch.tensor(labels, device=cuda), torch.tensor(lab_len - 1, device=cuda))
    st = bridge()
    out, final = dec(st, helper)
    torch.cuda.synchronize()
    assert out.logits.shape[0] == T_out - 1                       # time-major
    from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.attention_decoder import AttentionDecoderOutput
    out_bm = AttentionDecoderOutput(*[x.transpose(0, 1) for x in out])
    ref = odec.decode(p, "hybrid", enc, lens, (st.c.cpu().numpy(), st.h.cpu().numpy()), labels=labels,
                      labels_seq_len=lab_len, feed_previous_attention=feed_prev)
    check(out_bm, ref)
    for b in range(B):                                            # imputed past each length
        assert float(out_bm.logits[b, lab_len[b] - 1:].abs().sum()) == 0.0
    np.testing.assert_allclose(final.h.cpu().numpy(), ref["final_state"][1], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("atype,feed_prev,lpw", [("bahdanau_content", False, 0.6), ("location", True, 0.6),
                                                 ("hybrid", True, 0.0), ("luong_general", False, 1.0)])
def test_beam_search_vs_oracle(cuda, atype, feed_prev, lpw):
    """b2_attention_decoder_beam_search (every utterance its own beam) vs the numpy restatement of
    the reference's beam_search_step, one utterance at a time.  Utterances whose search came within
    1e-3 of a score tie (fp32 vs fp64 could order candidates differently) are skipped; the others
    must give identical hypotheses, lengths and (to 1e-4) log-probs / scores for every beam."""
    B, T, H_enc, Hd, A, emb, C, W, L = 6, 30, 16, 32, 24, 12, 9, 4, 14
    sos, eos = C - 2, C - 1
    dec, bridge, embedding, enc_out, p, (enc, lens, fs) = build(cuda, atype, B, T, H_enc, Hd, A, emb, C, True, 21,
                                                                 max_len=L, feed_prev=feed_prev)
    dec.variables["output_layer/biases"][eos] += 1.0          # make <EOS> reasonably likely
    p["output_layer/biases"] = dec.variables["output_layer/biases"].cpu().numpy()
    st = bridge()
    ids, lengths, log_probs, scores = dec.beam_search(st, embedding, sos, eos, W, lpw)
    torch.cuda.synchronize()
    ids, lengths = ids.cpu().numpy(), lengths.cpu().numpy()
    log_probs, scores = log_probs.cpu().numpy(), scores.cpu().numpy()
    c0, h0 = st.c.cpu().numpy(), st.h.cpu().numpy()
    compared = 0
    for b in range(B):
        ref = odec.beam_search_decode(p, atype, enc[b], lens[b], (c0[b], h0[b]), sos, eos, W, lpw, L,
                                      feed_previous_attention=feed_prev)
        if ref["min_margin"] < 1e-3:
            continue
        compared += 1
        Lr = ref["ids"].shape[1]
        assert ids.shape[2] >= Lr
        assert np.array_equal(ids[b, :, :Lr], ref["ids"]), (b, ids[b], ref["ids"])
        assert np.all(ids[b, :, Lr:] == eos)
        assert np.array_equal(lengths[b], ref["lengths"])
        np.testing.assert_allclose(log_probs[b], ref["log_probs"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(scores[b], ref["scores"], rtol=1e-4, atol=1e-4)
    assert compared >= 3


def test_beam_width_one_equals_greedy(cuda):
    from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.helpers import GreedyEmbeddingHelper
    B, C = 5, 9
    sos, eos = C - 2, C - 1
    dec, bridge, embedding, enc_out, p, _ = build(cuda, "bahdanau_content", B, 25, 16, 32, 24, 12, C, False, 31, max_len=12)
    st = bridge()
    helper = GreedyEmbeddingHelper(embedding, torch.full((B,), sos, dtype=torch.int32, device=cuda), eos)
    out, _ = dec(st, helper)
    ids, lengths, _, _ = dec.beam_search(st, embedding, sos, eos, 1, 0.0)
    g = out.predicted_ids.cpu().numpy()
    bm = ids[:, 0].cpu().numpy()
    for b in range(B):
        hit = np.where(g[b] == eos)[0]
        n = hit[0] + 1 if len(hit) else g.shape[1]
        assert np.array_equal(bm[b, :n], g[b, :n])
