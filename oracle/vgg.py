"""VGG front-end of the VGG-BLSTM encoder, CPU restatement in torch (float64, autograd-able).
TEST INFRASTRUCTURE.

Follows ``models/encoders/core/vgg_blstm.py:93-177``: reshape ``[B,T,D]`` ->
``[B*T, num_channels, splice*num_stack, 3]`` (:108-110), VGG1 = conv3x3(3->64)+ReLU, dropout,
conv3x3(64->64)+ReLU, max_pool 2x2/2 SAME, dropout (:113-134); VGG2 the same with 64->128->128
(:136-157); flatten (:160-161); fully connected 256 + ReLU, dropout (:165-174); reshape back to
``[B,T,256]`` (:177).  ``conv_layer`` / ``max_pool``: ``models/encoders/core/cnn_util.py:52-84,
13-29`` (``tf.nn.conv2d`` NHWC, filter ``[H,W,C_in,C_out]``, stride 1, SAME, bias, ReLU;
``tf.nn.max_pool`` ksize 2, stride 2, SAME).

TF-upstream facts restated: SAME padding for a 3x3 stride-1 conv is one zero row/column on every
side; for the 2x2 stride-2 pool the output size is ceil(n/2) and the (at most one) padded
row/column is appended at the end and never wins the max.

params (numpy / torch, TF names): VGG{1,2}/conv{1,2}/weight [3,3,Cin,Cout], .../bias [Cout],
bridge/weights [H4*W4*128, 256], bridge/biases [256].
masks: optional list of 5 {0,1} arrays (NHWC shapes of the five dropout sites, the last one
[N,256]) applied as ``x * mask / keep_prob``.
"""
import numpy as np
import torch
import torch.nn.functional as F

CONVS = ("VGG1/conv1", "VGG1/conv2", "VGG2/conv1", "VGG2/conv2")


def _conv_relu(x_nhwc, w_hwio, b):
    x = x_nhwc.permute(0, 3, 1, 2)
    w = w_hwio.permute(3, 2, 0, 1)
    y = F.conv2d(x, w, b, stride=1, padding=1)
    return torch.relu(y).permute(0, 2, 3, 1)


def _pool(x_nhwc):
    x = x_nhwc.permute(0, 3, 1, 2)
    y = F.max_pool2d(x, kernel_size=2, stride=2, ceil_mode=True)
    return y.permute(0, 2, 3, 1)


def vgg_frontend(inputs_btd, p, num_channels, width, keep_prob=1.0, masks=None):
    """inputs [B,T,D] torch -> [B,T,256]"""
    B, T, D = inputs_btd.shape
    assert D == num_channels * width * 3
    t = lambda v: v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v), dtype=inputs_btd.dtype)
    x = inputs_btd.reshape(B * T, num_channels, width, 3)

    def drop(x, i):
        if masks is None or keep_prob >= 1.0:
            return x
        return x * t(masks[i]).reshape(x.shape) / keep_prob
    x = drop(_conv_relu(x, t(p["VGG1/conv1/weight"]), t(p["VGG1/conv1/bias"])), 0)
    x = drop(_pool(_conv_relu(x, t(p["VGG1/conv2/weight"]), t(p["VGG1/conv2/bias"]))), 1)
    x = drop(_conv_relu(x, t(p["VGG2/conv1/weight"]), t(p["VGG2/conv1/bias"])), 2)
    x = drop(_pool(_conv_relu(x, t(p["VGG2/conv2/weight"]), t(p["VGG2/conv2/bias"]))), 3)
    x = x.reshape(B * T, -1)
    x = drop(torch.relu(x @ t(p["bridge/weights"]) + t(p["bridge/biases"])), 4)
    return x.reshape(B, T, 256)


def output_geometry(num_channels, width):
    h2, w2 = (num_channels + 1) // 2, (width + 1) // 2
    return (h2 + 1) // 2, (w2 + 1) // 2


def conv_relu_numpy(x_nhwc, w_hwio, b):
    """Literal loops (float64), independent of torch; cross-checks ``_conv_relu`` on tiny shapes."""
    x = np.asarray(x_nhwc, np.float64)
    w = np.asarray(w_hwio, np.float64)
    N, H, W, C = x.shape
    xp = np.zeros((N, H + 2, W + 2, C))
    xp[:, 1:H + 1, 1:W + 1] = x
    y = np.zeros((N, H, W, w.shape[3]))
    for dh in range(3):
        for dw in range(3):
            y += np.einsum("nhwc,co->nhwo", xp[:, dh:dh + H, dw:dw + W], w[dh, dw])
    return np.maximum(y + np.asarray(b, np.float64), 0.0)


def max_pool_numpy(x_nhwc):
    x = np.asarray(x_nhwc, np.float64)
    N, H, W, C = x.shape
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    xp = np.full((N, 2 * H2, 2 * W2, C), -np.inf)
    xp[:, :H, :W] = x
    return xp.reshape(N, H2, 2, W2, 2, C).max(axis=(2, 4))
