"""Data-parallel host logic on CPU (gloo, world_size 2): the all-reduce form of
utils/training/multi_gpu.py::average_gradients equals the oracle's tower mean, and the
list-of-towers form keeps the reference call shape."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import optim as oopt


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensorflow_end2end_speech_recognition_b200.utils.training.multi_gpu import allreduce_mean_
    rng = np.random.RandomState(100 + rank)
    shapes = [(7, 5), (11,), (3, 4)]
    grads = [rng.randn(*s).astype(np.float32) * (10.0 if i == 0 else 1.0) for i, s in enumerate(shapes)]
    # the step order of examples/librispeech/training/train_ctc.py:116,143: clip per tower, then mean
    clipped = [oopt.clip_by_norm(g, 5.0) for g in grads]
    flat = torch.tensor(np.concatenate([g.reshape(-1) for g in clipped])) / world   # 1/N folded in
    allreduce_mean_(flat, world)
    np.save(os.path.join(outdir, "flat_%d.npy" % rank), flat.numpy())
    np.save(os.path.join(outdir, "clipped_%d.npy" % rank),
            np.concatenate([g.reshape(-1) for g in clipped]))
    dist.destroy_process_group()


def test_allreduce_mean_matches_oracle_tower_average(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    flats = [np.load(tmp_path / ("flat_%d.npy" % r)) for r in range(world)]
    towers = [[np.load(tmp_path / ("clipped_%d.npy" % r))] for r in range(world)]
    ref = oopt.average_gradients(towers)[0]
    for f in flats:                      # every rank holds the same averaged gradient
        np.testing.assert_allclose(f, ref, rtol=1e-6, atol=1e-7)


def test_sparsetensor_round_trip():
    from tensorflow_end2end_speech_recognition_b200.utils.io.labels.sparsetensor import (
        list2sparsetensor, sparse_to_label_lists, sparsetensor2list)
    labels = np.array([[1, 2, 3, -1], [4, -1, -1, -1], [5, 6, 7, 8]])
    st = list2sparsetensor(labels, -1)
    assert st[0].dtype == np.int64 and st[1].dtype == np.int32 and list(st[2]) == [3, 4]
    back = sparsetensor2list(st, 3)
    assert [list(b) for b in back] == [[1, 2, 3], [4], [5, 6, 7, 8]]
    assert sparse_to_label_lists(st, 3) == [[1, 2, 3], [4], [5, 6, 7, 8]]


def test_shard_bounds_match_array_split():
    """rank shards of the per-GPU split = np.array_split (utils/dataset/ctc.py:171-177)"""
    import numpy as np
    from tensorflow_end2end_speech_recognition_b200.utils.io.inputs.pipeline import pad_labels, shard_bounds
    for B in (1, 7, 8, 64, 65):
        for n in (1, 2, 3, 8):
            want = [len(a) for a in np.array_split(np.arange(B), n)]
            got = shard_bounds(B, n)
            assert [e - s for s, e in got] == want and got[0][0] == 0 and got[-1][1] == B
    lab = pad_labels([[1, 2, 3], [4]], padded_value=-1)
    assert lab.tolist() == [[1, 2, 3], [4, -1, -1]]
