#!/usr/bin/env python
"""The reference's "working check" (models/test/test_ctc.py:24-240) on the B200 path: one synthetic
utterance replicated B times, trained until the greedy label error rate drops below 0.1, using the
same model-construction and train-step calls as examples/timit/training/train_ctc.py:325-338,65-144.

    python examples/train_ctc_synthetic.py [--precision bf16|fp32] [--steps 400]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorflow_end2end_speech_recognition_b200.models.ctc.ctc import CTC  # noqa: E402
from tensorflow_end2end_speech_recognition_b200.utils.io.labels.sparsetensor import (  # noqa: E402
    list2sparsetensor, sparsetensor2list)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--beam_width", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.RandomState(0)
    batch_size, max_time, input_size, num_classes = 8, 120, 120, 28
    model = CTC(encoder_type="blstm", input_size=input_size, num_units=256, num_layers=2,
                num_classes=num_classes, lstm_impl="LSTMBlockCell", use_peephole=True,
                parameter_init=0.1, clip_grad_norm=5.0, clip_activation=50, weight_decay=0.0,
                precision=args.precision)
    model.create_placeholders()
    inputs = np.repeat(rng.randn(1, max_time, input_size).astype(np.float32), batch_size, axis=0)
    labels = np.repeat(rng.randint(0, num_classes, size=(1, 25)), batch_size, axis=0)
    inputs_seq_len = np.full(batch_size, max_time, np.int32)
    labels_st = list2sparsetensor(labels, padded_value=-1)
    t0 = time.time()
    for step in range(args.steps):
        loss, logits = model.compute_loss(inputs, labels_st, inputs_seq_len, keep_prob=1.0)
        model.train(loss, optimizer="adam", learning_rate=1e-3)
        if (step + 1) % 10 == 0:
            decode_st = model.decoder(logits, inputs_seq_len, beam_width=args.beam_width)
            ler = model.compute_ler(decode_st, labels_st)
            hyp = sparsetensor2list(decode_st, batch_size)[0]
            print("Step %d: loss = %.3f / ler = %.3f (%.2f sec)  hyp[0][:12] = %s" %
                  (step + 1, float(loss), ler, time.time() - t0, list(hyp[:12])), flush=True)
            if ler < 0.1:
                print("Model is Converged.")
                break


if __name__ == "__main__":
    main()
