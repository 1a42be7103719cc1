"""Device input pipeline (b2_stack_splice behind DeviceInputPipeline) -- bit-exact against
(1) the golden vectors produced by the reference's own stack_frame / do_splice
(tests/golden/input_pipeline.npz) and (2) the closed-form oracle on ragged batches, including
the per-GPU split."""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "input_pipeline.npz")


def test_golden_vectors_bit_exact(cuda):
    from tensorflow_end2end_speech_recognition_b200.utils.io.inputs.pipeline import DeviceInputPipeline
    g = np.load(GOLD)
    for i in range(int(g["n_cases"])):
        T, nch, S, K, P = [int(v) for v in g["cfg_%d" % i]]
        pipe = DeviceInputPipeline(S, K, P, device=cuda)
        out, out_len = pipe([g["x_%d" % i]])
        want = g["spliced_%d" % i]
        assert int(out_len[0]) == want.shape[0]
        assert np.array_equal(out[0].cpu().numpy(), want), (i, T, nch, S, K, P)


@pytest.mark.parametrize("S,K,P", [(1, 1, 1), (3, 3, 1), (3, 2, 1), (2, 2, 11), (1, 1, 5), (4, 3, 3)])
def test_ragged_batch_and_sharding(cuda, S, K, P):
    from tensorflow_end2end_speech_recognition_b200.utils.io.inputs.pipeline import DeviceInputPipeline, shard_bounds
    rng = np.random.RandomState(S * 100 + K * 10 + P)
    B, nch = 7, 5
    xs = [rng.randn(int(rng.randint(1, 60)), nch * 3).astype(np.float32) for _ in range(B)]
    labels = [list(rng.randint(0, 20, int(rng.randint(1, 9)))) for _ in range(B)]
    want_in, want_lab, want_len = oin.make_batch(xs, labels, S, K, P, padded_value=-1, num_gpu=1)
    pipe = DeviceInputPipeline(S, K, P, num_gpu=3, device=cuda)
    out, out_len = pipe(xs)
    assert np.array_equal(out.cpu().numpy(), want_in[0])
    assert np.array_equal(out_len.cpu().numpy(), want_len[0])
    # per-rank shards == np.array_split of the globally padded batch
    sp_in, _, sp_len = oin.make_batch(xs, labels, S, K, P, padded_value=-1, num_gpu=3)
    assert [e - s for s, e in shard_bounds(B, 3)] == [a.shape[0] for a in sp_in]
    for r in range(3):
        o, l = pipe(xs, rank=r)
        assert np.array_equal(o.cpu().numpy(), sp_in[r]) and np.array_equal(l.cpu().numpy(), sp_len[r])


def test_rejects_bad_parameters(cuda):
    from tensorflow_end2end_speech_recognition_b200.utils.io.inputs.pipeline import DeviceInputPipeline
    with pytest.raises(ValueError):
        DeviceInputPipeline(2, 3, 1, device=cuda)
    with pytest.raises(RuntimeError):
        DeviceInputPipeline(1, 1, 3, device=cuda)([np.zeros((4, 4), np.float32)])      # 4 % 3 != 0
