"""Single-GPU step time of the non-headline BASELINE configs (parity-test cases, timed for
the record): cfg3 joint CTC-attention (4x512 BLSTM + 256 LSTM decoder, hybrid attention) and
cfg4 VGG-BLSTM 6x1024 3000-class CTC.  Usage: bench_configs.py cfg3|cfg4 [B] [T]"""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")


def timed(fn, n=3):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for _ in range(n):
        torch.cuda.synchronize()
        ev0.record()
        fn()
        ev1.record()
        torch.cuda.synchronize()
        out.append(ev0.elapsed_time(ev1))
    return out


def cfg3(B, T, precision):
    from tensorflow_end2end_speech_recognition_b200.models.attention.joint_ctc_attention import JointCTCAttention
    V, Lout = 28, 200
    m = JointCTCAttention(input_size=80, encoder_type="blstm", encoder_num_units=512, encoder_num_layers=4,
                          encoder_num_proj=None, attention_type="hybrid", attention_dim=128, decoder_type="lstm",
                          decoder_num_units=256, decoder_num_layers=1, embedding_dim=64, lambda_weight=0.5,
                          num_classes=V, sos_index=V, eos_index=V + 1, max_decode_length=300,
                          precision=precision, device="cuda:0")
    rng = np.random.RandomState(0)
    x = torch.tensor(rng.randn(B, T, 80).astype(np.float32), device="cuda:0")
    seq = np.full(B, T, np.int32)
    labels = np.full((B, Lout), V + 1, np.int32)
    labels[:, 0] = V
    labels[:, 1:Lout - 1] = rng.randint(0, V, (B, Lout - 2))
    lab_len = np.full(B, Lout, np.int32)
    ctc_labels = [list(labels[b, 1:Lout - 1]) for b in range(B)]

    def step():
        loss, *_ = m.compute_loss(x, labels, ctc_labels, seq, lab_len, 1.0, 1.0, 1.0)
        m.train(loss, "adam", 1e-3)
        step.loss = loss
    ms = timed(step)
    print("cfg3 joint B=%d T=%d L_out=%d %s: step ms %s  loss %.4f  -> %.0f frames/s" %
          (B, T, Lout, precision, ["%.1f" % v for v in ms], float(step.loss), B * T / (min(ms) * 1e-3)))

    def fwd_only():
        m.compute_loss(x, labels, ctc_labels, seq, lab_len, 1.0, 1.0, 1.0, is_training=False)
    print("  forward only ms", ["%.1f" % v for v in timed(fwd_only, 2)])

    def infer():
        _, _, _, _, oi = m.compute_loss(x, labels, ctc_labels, seq, lab_len, 1.0, 1.0, 1.0, is_training=False)
        infer.ids = oi.predicted_ids
    print("  forward + greedy decode (max 300) ms", ["%.1f" % v for v in timed(infer, 2)], tuple(infer.ids.shape))


def cfg4(B, T, precision):
    from tensorflow_end2end_speech_recognition_b200.models.ctc.ctc import CTC
    m = CTC(encoder_type="vgg_blstm", input_size=240, num_units=1024, num_layers=6, num_classes=3000,
            clip_grad_norm=5.0, precision=precision, device="cuda:0")
    rng = np.random.RandomState(0)
    x = torch.tensor(rng.randn(B, T, 240).astype(np.float32), device="cuda:0")
    seq = np.full(B, T, np.int32)
    labels = [list(rng.randint(0, 3000, int(rng.randint(30, 71)))) for _ in range(B)]

    def step():
        loss, _ = m.compute_loss(x, labels, seq, keep_prob=0.8)
        m.train(loss, "rmsprop", 1e-3)
        step.loss = loss
    for _ in range(5):          # allocator / lazy module loading settle over the first steps (host-bound 0.3-0.9 s each)
        step()
    t0 = time.time()
    ms = timed(step, 3)
    print("cfg4 vgg_blstm 6x1024 C=3001 B=%d T=%d %s: step ms %s loss %.3f -> %.0f frames/s (wall %.1fs)" %
          (B, T, precision, ["%.1f" % v for v in ms], float(step.loss), B * T / (min(ms) * 1e-3), time.time() - t0))


if __name__ == "__main__":
    which = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (8 if which == "cfg3" else 4)
    T = int(sys.argv[3]) if len(sys.argv) > 3 else (1000 if which == "cfg3" else 1500)
    prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
    (cfg3 if which == "cfg3" else cfg4)(B, T, prec)
