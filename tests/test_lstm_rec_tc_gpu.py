"""tcgen05 persistent recurrence (lstm_rec_tc.cu) behind b2_blstm_layer_forward with
precision=bf16: encoder states vs the fp64 oracle.  Operands of the step GEMM (Wh, h)
are rounded to bf16, so the tolerance is that of a bf16-operand recurrence;
gradients go through the shared reserve layout and the fp32 BPTT."""
import os

import numpy as np
import pytest

from tests.test_lstm_gpu import compare, run_layer

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nchain", [1, 2])
@pytest.mark.parametrize("T,B,D,H", [(6, 16, 32, 32), (12, 24, 40, 64), (20, 32, 64, 256),
                                     (16, 64, 80, 512), (9, 5, 24, 128)])
def test_rec_tc_forward_backward(cuda, nchain, T, B, D, H):
    from tensorflow_end2end_speech_recognition_b200 import ops
    os.environ["B2_REC_NCHAIN"] = str(nchain)
    try:
        seq = [T] + list(np.random.RandomState(T + B).randint(max(T // 2, 1), T + 1, size=B - 1))
        # reference init scale (parameter_init 0.1, blstm.py:79-80) for the wide layers
        got, ref = run_layer(cuda, T, B, D, H, seq, ops.PREC_BF16, seed=11,
                             parameter_init=0.1 if H >= 256 else 0.3)
    finally:
        os.environ.pop("B2_REC_NCHAIN", None)
    compare(got, ref, 3e-2, 6e-2)


def test_rec_tc_matches_fp32_step_kernels_long(cuda):
    """T=300: drift of the bf16 recurrence against the fp32 CUDA-core recurrence stays bounded."""
    import torch
    from tensorflow_end2end_speech_recognition_b200 import ops
    from oracle import lstm as olstm
    T, B, D, H = 300, 32, 80, 256
    rng = np.random.RandomState(0)
    layer = olstm.init_blstm_params(D, H, 1, parameter_init=0.1, seed=0)[0]
    P = {d: {k: torch.tensor(v, device=cuda) for k, v in layer[d].items()} for d in layer}
    x = torch.tensor(rng.randn(T, B, D).astype(np.float32), device=cuda)
    seq = torch.tensor(np.sort(rng.randint(T // 2, T + 1, size=B))[::-1].astype(np.int32).copy(), device=cuda)
    ys = {}
    for name, prec in (("fp32", ops.PREC_FP32), ("bf16", ops.PREC_BF16)):
        desc = ops.lstm_desc(T, B, D, H, precision=prec, need_backward=False)
        y, fs, _ = ops.blstm_layer_forward(desc, x, seq, P["fw"], P["bw"], want_final_state=True)
        torch.cuda.synchronize()
        ys[name] = (y.cpu().numpy(), fs.cpu().numpy())
    err = np.abs(ys["fp32"][0] - ys["bf16"][0]).max()
    assert err < 2e-2, err
    assert np.abs(ys["fp32"][1] - ys["bf16"][1]).max() < 3e-2


def test_rec_tc_dropout_matches_oracle_mask(cuda):
    """DropoutWrapper(output_keep_prob) in the bf16 path: same counter-hash mask as the oracle,
    applied to the emitted output only (the bf16 shadow fed to the next layer included)."""
    from tensorflow_end2end_speech_recognition_b200 import ops
    T, B, D, H = 14, 16, 32, 64
    seq = [T] + list(np.random.RandomState(3).randint(T // 2, T + 1, size=B - 1))
    got, ref = run_layer(cuda, T, B, D, H, seq, ops.PREC_BF16, seed=21, keep_prob=0.75)
    compare(got, ref, 3e-2, 6e-2)


def test_hybrid_fallback_path_when_units_not_multiple_of_32(cuda):
    """H=48 cannot use the cluster/TMEM recurrence -> tcgen05 GEMMs + fp32 step kernels."""
    from tensorflow_end2end_speech_recognition_b200 import ops
    T, B, D, H = 10, 4, 16, 48
    got, ref = run_layer(cuda, T, B, D, H, [10, 7, 10, 3], ops.PREC_BF16, seed=23)
    compare(got, ref, 3e-2, 6e-2)


def test_single_step_and_tiny_batch(cuda):
    from tensorflow_end2end_speech_recognition_b200 import ops
    got, ref = run_layer(cuda, 1, 1, 32, 32, [1], ops.PREC_BF16, seed=25)
    compare(got, ref, 3e-2, 6e-2)
    got, ref = run_layer(cuda, 2, 3, 32, 32, [2, 1, 2], ops.PREC_BF16, seed=26)
    compare(got, ref, 3e-2, 6e-2)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
def test_wide_layer_per_step_gemm_path(cuda, precision, tol):
    """H > 512 (config 4 uses 1024): recurrent product as one split-K skinny GEMM per direction and frame
    in front of the gate-math step kernel, forward and BPTT."""
    from tensorflow_end2end_speech_recognition_b200 import ops
    T, B, D, H = 7, 5, 24, 576
    prec = ops.PREC_FP32 if precision == "fp32" else ops.PREC_BF16
    got, ref = run_layer(cuda, T, B, D, H, [7, 4, 7, 2, 6], prec, seed=29, parameter_init=0.05)
    compare(got, ref, tol, 2 * tol)


@pytest.mark.parametrize("T,B,D,H", [(9, 5, 24, 640), (14, 32, 40, 1024), (6, 17, 16, 768), (1, 3, 16, 896)])
def test_grid_resident_wide_layer(cuda, T, B, D, H):
    """512 < H <= 1024 (config 4: 6 x 1024), precision bf16: the forward recurrence is ONE cooperative launch
    (lstm_wide.cu: recurrent weights register-resident as mma.sync fragments, h exchanged through L2 behind a grid
    barrier per step) feeding the fp32 reserve of the BPTT path; forward + gradients vs the fp64 oracle at the bf16
    tolerance, ragged lengths, final state; and against the per-frame fallback (B2_WIDE_REC=0)."""
    import torch
    from tensorflow_end2end_speech_recognition_b200 import ops
    seq = [T] + list(np.random.RandomState(T + B).randint(max(T // 2, 1), T + 1, size=B - 1))
    got, ref = run_layer(cuda, T, B, D, H, seq, ops.PREC_BF16, seed=31, parameter_init=0.05)
    compare(got, ref, 3e-2, 6e-2)
    os.environ["B2_WIDE_REC"] = "0"
    try:
        old, _ = run_layer(cuda, T, B, D, H, seq, ops.PREC_BF16, seed=31, parameter_init=0.05)
    finally:
        os.environ.pop("B2_WIDE_REC", None)
    # the fallback multiplies h by fp32 recurrent weights, the resident kernel by their bf16 rounding
    assert np.abs(got[0] - old[0]).max() < 2e-2, np.abs(got[0] - old[0]).max()


def test_forward_chunk_chain_bit_identical(cuda):
    """Layer l+1's gate GEMM issued chunk by chunk behind the progress counters of layer l's still-running recurrence
    (lstm_tc.cu, FwdChain) must give exactly the logits of the unchunked schedule: same GEMM, same K order per output
    element, only the launch schedule differs.  T = 96 -> 8 chunks of 12 frames, ragged lengths, 4 layers."""
    import torch
    from tensorflow_end2end_speech_recognition_b200.models.ctc.ctc import CTC
    rng = np.random.RandomState(3)
    B, T, D, H, L, C = 24, 96, 40, 128, 4, 12
    x = rng.randn(B, T, D).astype(np.float32)
    seq = rng.randint(T // 2, T + 1, size=B).astype(np.int32)
    seq[0] = T
    labels = [list(rng.randint(0, C - 1, size=int(rng.randint(3, 9)))) for _ in range(B)]
    outs = {}
    for chunks in ("0", "8", "5"):
        os.environ["B2_FWD_CHUNKS"] = chunks
        try:
            model = CTC(encoder_type="blstm", input_size=D, num_units=H, num_layers=L, num_classes=C,
                        parameter_init=0.1, clip_grad_norm=5.0, precision="bf16", device=cuda, seed=4)
            for _ in range(2):          # second pass: the chain state of the first one is stale, must not match
                loss, logits = model.compute_loss(x, labels, seq, keep_prob=1.0)
                model._backward()
            torch.cuda.synchronize()
            outs[chunks] = (float(loss), logits.cpu().numpy().copy(), model.flat_grads.cpu().numpy().copy())
        finally:
            os.environ.pop("B2_FWD_CHUNKS", None)
    for chunks in ("8", "5"):
        assert outs[chunks][0] == outs["0"][0]
        assert np.array_equal(outs[chunks][1], outs["0"][1])
        # weight gradients are accumulated with split-K atomics: equal up to summation order
        g0 = outs["0"][2]
        assert np.abs(outs[chunks][2] - g0).max() <= 1e-5 * np.abs(g0).max()


@pytest.mark.parametrize("T", [40, 80])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
def test_encoder_stack_with_dropout_matches_oracle_masks(cuda, precision, tol, T):
    """Three BLSTM layers with DropoutWrapper(output_keep_prob = 0.8): every layer's counter-hash mask reproduced for the
    oracle (tests/util_dropout.py), encoder output and ALL gradients (incl. d(inputs)) vs fp64 autograd -- the backward
    masks travel through the stack (layer l's dy is layer l+1's dx; on the tcgen05 path layer l+1 applies layer l's mask
    in the store of its dX GEMM, BLSTMEncoder.backward)."""
    import torch
    from oracle import lstm as olstm
    from tests.util_dropout import dropout_mask
    from tensorflow_end2end_speech_recognition_b200.models.encoders.core.blstm import BLSTMEncoder
    rng = np.random.RandomState(17)
    # T = 80: the chunked dX GEMMs (16 chunks of 5 frames), each storing through the fused mask of the layer below
    B, D, H, L, keep, dseed = 6, 24, 64, 3, 0.8, 5
    enc = BLSTMEncoder(H, None, L, "LSTMBlockCell", True, 0.2, None, time_major=True, precision=precision)
    named = enc.create_variables(D, rng)
    variables = {n: torch.tensor(a, device=cuda) for n, a in named}
    grads = {n: torch.zeros_like(v) for n, v in variables.items()}
    x = rng.randn(B, T, D).astype(np.float32)
    seq = np.array([T, T - 9, T, T // 2 + 2, T - 4, T - 13], np.int32)
    for b in range(B):
        x[b, seq[b]:] = 0
    dy = rng.randn(T, B, 2 * H).astype(np.float32)
    xd, sd = torch.tensor(x, device=cuda), torch.tensor(seq, device=cuda)
    y, _ = enc(xd, sd, keep, True, variables=variables, dropout_seed=dseed)
    dx = enc.backward(torch.tensor(dy, device=cuda), variables, grads, need_dx=True)
    torch.cuda.synchronize()
    # oracle with the same masks
    vs = {n: torch.tensor(np.asarray(a, np.float64), requires_grad=True) for n, a in named}
    layers = []
    for i in range(1, L + 1):
        layers.append({d: {k: vs["blstm_hidden%d/%s/lstm_cell/%s" % (i, d, k)]
                           for k in ("kernel", "bias", "w_i_diag", "w_f_diag", "w_o_diag")} for d in ("fw", "bw")})
    masks = [torch.tensor(dropout_mask(dseed * 131 + i, T * B * 2 * H, keep).reshape(T, B, 2 * H).astype(np.float64))
             for i in range(1, L + 1)]
    xt = torch.tensor(x.astype(np.float64), requires_grad=True)
    y_ref, _ = olstm.blstm_forward(xt, seq, layers, keep_prob=keep, dropout_masks=masks)
    (y_ref * torch.tensor(dy.astype(np.float64))).sum().backward()
    yr = y_ref.detach().numpy()
    assert np.abs(y.cpu().numpy() - yr).max() <= tol * max(1.0, np.abs(yr).max())
    g = xt.grad.numpy().transpose(1, 0, 2)
    assert np.abs(dx.cpu().numpy() - g).max() <= 2 * tol * max(1e-3, np.abs(g).max()), "d(inputs)"
    for n in vs:
        g = vs[n].grad.numpy()
        assert np.abs(grads[n].cpu().numpy() - g).max() <= 2 * tol * max(1e-3, np.abs(g).max()), n
