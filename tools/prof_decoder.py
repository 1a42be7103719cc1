"""One teacher-forced decoder pass + backward at cfg3 shapes with a short label length, for an ncu launch list."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tensorflow_end2end_speech_recognition_b200.models.attention.joint_ctc_attention import JointCTCAttention

B, T, Lout, V = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 28
m = JointCTCAttention(input_size=80, encoder_type="blstm", encoder_num_units=512, encoder_num_layers=1,
                      encoder_num_proj=None, attention_type=sys.argv[4] if len(sys.argv) > 4 else "hybrid",
                      attention_dim=128, decoder_type="lstm",
                      decoder_num_units=256, decoder_num_layers=1, embedding_dim=64, lambda_weight=0.5,
                      num_classes=V, sos_index=V, eos_index=V + 1, max_decode_length=300, precision="bf16", device="cuda:0")
rng = np.random.RandomState(0)
x = torch.tensor(rng.randn(B, T, 80).astype(np.float32), device="cuda:0")
seq = np.full(B, T, np.int32)
labels = np.full((B, Lout), V + 1, np.int32)
labels[:, 0] = V
labels[:, 1:Lout - 1] = rng.randint(0, V, (B, Lout - 2))
lab_len = np.full(B, Lout, np.int32)
ctc_labels = [list(labels[b, 1:Lout - 1]) for b in range(B)]
for _ in range(2):
    loss, *_ = m.compute_loss(x, labels, ctc_labels, seq, lab_len, 1.0, 1.0, 1.0)
    m.train(loss, "adam", 1e-3)
torch.cuda.synchronize()
print("ok", float(loss))
