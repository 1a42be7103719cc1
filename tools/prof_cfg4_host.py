"""Host-side profile of one config-4 training step (where does the CPU block?)."""
import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from tensorflow_end2end_speech_recognition_b200.models.ctc.ctc import CTC
B, T = 32, 1500
m = CTC(encoder_type="vgg_blstm", input_size=240, num_units=1024, num_layers=6, num_classes=3000,
        clip_grad_norm=5.0, precision="bf16", device="cuda:0")
rng = np.random.RandomState(0)
x = torch.tensor(rng.randn(B, T, 240).astype(np.float32), device="cuda:0")
seq = np.full(B, T, np.int32)
labels = [list(rng.randint(0, 3000, int(rng.randint(30, 71)))) for _ in range(B)]


def step():
    loss, _ = m.compute_loss(x, labels, seq, keep_prob=0.8)
    m.train(loss, "rmsprop", 1e-3)
    return loss


for i in range(6):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    st = torch.cuda.memory_stats()
    print("step %d: host %.1f ms, total %.1f ms | allocated %.1f GB, reserved %.1f GB, peak reserved %.1f GB, retries %d, "
          "cudaMalloc calls %d" % (i, (t1 - t0) * 1e3, (t2 - t0) * 1e3, st["allocated_bytes.all.current"] / 2**30,
          st["reserved_bytes.all.current"] / 2**30, st["reserved_bytes.all.peak"] / 2**30, st["num_alloc_retries"],
          st["num_device_alloc"]), flush=True)
torch.cuda.synchronize()
t0 = time.perf_counter(); l = step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue %.1f ms, until GPU idle %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
