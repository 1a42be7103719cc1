"""clip_by_norm + TF-1.x optimizer kernels vs oracle/optim.py."""
import numpy as np
import pytest
import torch

from oracle import optim as oopt

pytestmark = pytest.mark.gpu


def test_clip_by_norm_multi(cuda):
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(0)
    gs = [rng.randn(*s).astype(np.float32) * sc for s, sc in
          [((300, 70), 1.0), ((5,), 0.01), ((1024, 33), 0.2), ((1,), 100.0)]]
    ts = [torch.tensor(g, device=cuda) for g in gs]
    tl = ops.TensorList(ts)
    norms = ops.clip_by_norm_multi(tl, 5.0, post_scale=0.5)
    torch.cuda.synchronize()
    for g, t in zip(gs, ts):
        np.testing.assert_allclose(t.cpu().numpy(), 0.5 * oopt.clip_by_norm(g, 5.0), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(np.sqrt(norms.cpu().numpy()), [np.linalg.norm(g) for g in gs], rtol=1e-4)


@pytest.mark.parametrize("name", oopt.OPTIMIZERS)
def test_optimizers(cuda, name):
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(1)
    shapes = [(64, 48), (7,), (1000,)]
    w = [rng.randn(*s).astype(np.float32) for s in shapes]
    ref_w = [a.astype(np.float64).copy() for a in w]
    opt = oopt.Optimizer(name, 1e-2)
    dw = [torch.tensor(a, device=cuda) for a in w]
    init0 = {"adagrad": 0.1, "rmsprop": 1.0}.get(name, 0.0)
    s0 = [torch.full_like(t, init0) for t in dw]
    s1 = [torch.zeros_like(t) for t in dw]
    P, S0, S1 = ops.TensorList(dw), ops.TensorList(s0), ops.TensorList(s1)
    for step in range(1, 6):
        g = [rng.randn(*s).astype(np.float32) for s in shapes]
        dg = [torch.tensor(a, device=cuda) for a in g]
        ops.optimizer_step_multi(name, P, ops.TensorList(dg), S0, S1, 1e-2, step)
        opt.step(ref_w, [a.astype(np.float64) for a in g])
    torch.cuda.synchronize()
    for t, r in zip(dw, ref_w):
        np.testing.assert_allclose(t.cpu().numpy(), r, rtol=2e-4, atol=2e-5)
