import os, sys
os.environ["B2_REC_DBG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools.bench_rec import bench
from tensorflow_end2end_speech_recognition_b200 import ops
for nch in (1, 2):
    os.environ["B2_REC_NCHAIN"] = str(nch)
    print("nchain", nch, flush=True)
    bench(1000, 64, 1024, 512, ops.PREC_BF16, False, iters=1, label="dbg")
os.environ["B2_REC_NCHAIN"] = "1"
bench(1000, 16, 1024, 512, ops.PREC_BF16, False, iters=1, label="dbg B=16")
