"""cfg5, attention half: location-attention decoder, 256 utterances, T=1000, greedy and beam 20
(max_decode_length 300).  Encoder 4x512 BLSTM (bf16 path), decoder LSTM 256, attention_dim 128.
Usage: bench_attention_decode.py [B] [T] [beam] [attention_type] [feed_prev]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tensorflow_end2end_speech_recognition_b200.models.attention.attention_seq2seq import AttentionSeq2Seq

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
W = int(sys.argv[3]) if len(sys.argv) > 3 else 20
atype = sys.argv[4] if len(sys.argv) > 4 else "location"
feed_prev = (sys.argv[5] == "1") if len(sys.argv) > 5 else True
V = 28
m = AttentionSeq2Seq(input_size=80, encoder_type="blstm", encoder_num_units=512, encoder_num_layers=4,
                     encoder_num_proj=None, attention_type=atype, attention_dim=128, decoder_type="lstm",
                     decoder_num_units=256, decoder_num_layers=1, embedding_dim=64, num_classes=V, sos_index=V,
                     eos_index=V + 1, max_decode_length=int(__import__("os").environ.get("MAXLEN", "300")), precision="bf16", device="cuda:0",
                     feed_previous_attention=feed_prev)
# never emit <EOS>: every decode runs the full 300 steps (worst case, shape-independent of the weights)
m.variables["decoder/attention_decoder/output_layer/biases"][V + 1] = -1e4
rng = np.random.RandomState(0)
x = torch.tensor(rng.randn(B, T, 80).astype(np.float32), device="cuda:0")
seq = torch.full((B,), T, dtype=torch.int32, device="cuda:0")


def timed(fn, n=2):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for _ in range(n):
        torch.cuda.synchronize()
        ev0.record()
        r = fn()
        ev1.record()
        torch.cuda.synchronize()
        out.append(ev0.elapsed_time(ev1))
    return out, r


def enc_only():
    return m._encode(x, seq, 1.0, is_training=False)


def greedy():
    from tensorflow_end2end_speech_recognition_b200.models.attention.decoders.helpers import GreedyEmbeddingHelper
    enc = m._encode(x, seq, 1.0, is_training=False)
    m.decoder.encoder_outputs, m.decoder.encoder_outputs_seq_len = enc.outputs, enc.seq_len
    h = GreedyEmbeddingHelper(m.variables["decoder/output_embedding/W_embedding"],
                              torch.full((B,), V, dtype=torch.int32, device="cuda:0"), V + 1)
    out, _ = m.decoder(m.bridge(enc), h)
    return out.predicted_ids


t_enc, _ = timed(enc_only)
t_g, ids = timed(greedy)
print("%s feed_prev=%s  B=%d T=%d: encoder %.1f ms; encoder+greedy(%d steps) %.1f ms -> %.0f utt/s, %.3f ms/step" %
      (atype, feed_prev, B, T, min(t_enc), ids.shape[1], min(t_g), B / (min(t_g) * 1e-3), (min(t_g) - min(t_enc)) / ids.shape[1]))
t_b, res = timed(lambda: m.decode_beam(x, seq, beam_width=W), 2)
print("  beam %d (%d rows, %d steps): %.1f ms -> %.0f utt/s, %.3f ms/step" %
      (W, B * W, res[0].shape[2], min(t_b), B / (min(t_b) * 1e-3), (min(t_b) - min(t_enc)) / res[0].shape[2]))
