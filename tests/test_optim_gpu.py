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


def test_average_gradients_list_form(cuda):
    """utils/training/multi_gpu.py:13-48 call shape: list (towers) of [(grad, var)]; towers with a None gradient are
    skipped; the mean is the hand-written b2_tower_mean kernel"""
    import torch
    from tensorflow_end2end_speech_recognition_b200.utils.training.multi_gpu import average_gradients
    g0 = [(torch.ones(3, device=cuda), "v0"), (None, "v1")]
    g1 = [(3 * torch.ones(3, device=cuda), "v0"), (torch.ones(2, device=cuda), "v1")]
    out = average_gradients([g0, g1])
    assert out[0][1] == "v0" and torch.allclose(out[0][0], 2 * torch.ones(3, device=cuda))
    assert torch.allclose(out[1][0], torch.ones(2, device=cuda))


def test_tower_mean_kernel(cuda):
    import numpy as np
    import torch
    from tensorflow_end2end_speech_recognition_b200 import ops
    rng = np.random.RandomState(0)
    for n in (1, 5, 1024, 100003):
        for k in (1, 2, 3, 8):
            hs = [rng.randn(n).astype(np.float32) for _ in range(k)]
            ds = [torch.tensor(h, device=cuda) for h in hs]
            ops.tower_mean(ds, ds[0])                       # in place on tower 0
            ref = np.mean(np.stack(hs).astype(np.float64), 0)
            np.testing.assert_allclose(ds[0].cpu().numpy(), ref, rtol=1e-6, atol=1e-6)
    # unaligned views take the scalar path
    base = torch.tensor(rng.randn(4099).astype(np.float32), device=cuda)
    a, b = base[1:2050], base[2050:4099]
    want = (a.cpu().numpy().astype(np.float64) + b.cpu().numpy()) / 2
    out = torch.empty(2049, device=cuda)
    ops.tower_mean([a.contiguous(), b.contiguous()], out)
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-6, atol=1e-6)
