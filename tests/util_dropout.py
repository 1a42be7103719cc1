"""numpy restatement of csrc/common.cuh::dropout_keep so the oracle can apply the
same DropoutWrapper mask as the kernels."""
import numpy as np

M64 = (1 << 64) - 1


def dropout_mask(seed, n, keep_prob):
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = ((z >> np.uint64(32)) >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return (u < np.float32(keep_prob)).astype(np.float64)
