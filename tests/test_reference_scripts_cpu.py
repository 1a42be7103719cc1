"""SURVEY 7 step 1 / VERDICT r1 item 5: the reference's OWN training driver, unchanged, against this package's
drop-in surface.  ``examples/timit/training/train_ctc.py::do_train`` is loaded from /root/reference (never copied)
and executed with
  * ``compat.install(reference_root=...)``: ``import tensorflow`` -> compat.tf, ``models.*`` / ``utils.io.labels.
    sparsetensor`` / ``utils.training.multi_gpu`` / ``utils.evaluation.edit_distance`` -> this package, everything else
    (learning-rate controller, directory helpers, parameter counter, metrics) -> the reference's own files;
  * a synthetic ``Dataset`` with the reference constructor signature and iteration protocol
    (utils/dataset/ctc.py:32-182) registered as ``examples.timit.data.load_dataset_ctc``;
  * a model whose arithmetic is the CPU oracle (this container has no GPU; the CUDA model runs the same flow under
    the same shim in tests/test_compat_gpu.py and tests/test_towers_gpu.py) but whose graph-mode plumbing --
    placeholders, lazy op handles, ``train`` -> optimizer, ``decoder``, ``compute_ler`` -- is the package's own CTC
    class.
Skipped where /root/reference does not exist (the GPU box)."""
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


class SyntheticDataset(object):
    """Reference Dataset protocol: ``for step, (data, is_new_epoch) in enumerate(ds)``, ``ds.next()``,
    data = (inputs[num_gpu][B,T,D], labels[num_gpu][B,L] padded -1, inputs_seq_len[num_gpu][B], names)."""

    def __init__(self, data_type, label_type, batch_size, max_epoch=None, splice=1, num_stack=1, num_skip=1,
                 sort_utt=False, sort_stop_epoch=None, shuffle=False, num_gpu=1, **kw):
        self.data_type, self.label_type, self.batch_size, self.max_epoch = data_type, label_type, batch_size, max_epoch
        self.num_gpu = num_gpu
        self.rng = np.random.RandomState({"train": 0, "dev": 1}.get(data_type, 2))
        self.n_utt, self.epoch, self.cursor, self.is_new_epoch = 3 * batch_size, 0, 0, False
        self.padded_value = -1
        self.D, self.C = 12, 9
        self.utts = []
        for _ in range(self.n_utt):
            T = int(self.rng.randint(14, 22))
            L = int(self.rng.randint(2, 6))
            self.utts.append((self.rng.randn(T, self.D).astype(np.float32), self.rng.randint(0, self.C, L)))

    @property
    def epoch_detail(self):
        return self.epoch + self.cursor / float(self.n_utt)

    def __iter__(self):
        return self

    def __next__(self):
        if self.max_epoch is not None and self.epoch >= self.max_epoch:
            raise StopIteration
        idx = list(range(self.cursor, min(self.cursor + self.batch_size, self.n_utt)))
        self.cursor += len(idx)
        self.is_new_epoch = self.cursor >= self.n_utt
        if self.is_new_epoch:
            self.cursor, self.epoch = 0, self.epoch + 1
        T = max(self.utts[i][0].shape[0] for i in idx)
        Lm = max(len(self.utts[i][1]) for i in idx)
        x = np.zeros((len(idx), T, self.D), np.float32)
        y = np.full((len(idx), Lm), -1, np.int32)
        n = np.zeros(len(idx), np.int32)
        for k, i in enumerate(idx):
            f, l = self.utts[i]
            x[k, :f.shape[0]], y[k, :len(l)], n[k] = f, l, f.shape[0]
        split = lambda a: [s for s in np.array_split(a, self.num_gpu)]           # utils/dataset/ctc.py:171-177
        return (split(x), split(y), split(n), split(np.arange(len(idx)))), self.is_new_epoch

    next = __next__


def _oracle_backed_ctc():
    """the package's CTC class with the arithmetic routed to oracle/ (CPU)"""
    import torch
    from oracle import decode as odec
    from oracle import model as omodel
    from tensorflow_end2end_speech_recognition_b200.models.ctc import ctc as ctc_mod
    from tensorflow_end2end_speech_recognition_b200.utils.io.labels.sparsetensor import SparseTensorValue

    class OracleCTC(ctc_mod.CTC):
        def __init__(self, **kw):
            with mock.patch.object(ctc_mod.ops, "TensorList", lambda ts: list(ts)):
                kw.setdefault("device", "cpu")
                ctc_mod.CTC.__init__(self, **kw)
            self.steps_run = 0

        def _allocate_variables(self, named, device):
            with mock.patch("tensorflow_end2end_speech_recognition_b200.models.model_base.ops.TensorList",
                            lambda ts: list(ts)):
                ctc_mod.CTC._allocate_variables(self, named, "cpu")

        def _trainer(self):
            vs = {v.name: v.tensor.numpy() for v in self._variables}
            return omodel.OracleTrainer(vs, self.num_layers, optimizer="adam", learning_rate=1e-2,
                                        clip_grad_norm=self.clip_grad_norm, dtype=torch.float32)

        def _eval_logits(self, inputs, inputs_seq_len, keep_prob, is_training=True):
            self._last_feed = (np.asarray(inputs), np.asarray(inputs_seq_len))
            vs = {v.name: v.tensor for v in self._variables}
            with torch.no_grad():
                lens = np.asarray(inputs_seq_len)
                dummy = [[0]] * len(lens)
                _, logits, _ = omodel.ctc_model_forward(vs, torch.tensor(np.asarray(inputs)), lens, dummy,
                                                        self.num_layers)
            return logits

        def _eval_loss(self, logits, labels, inputs_seq_len, softmax_temperature=1, is_training=True):
            x, lens = self._last_feed
            lists = ctc_mod.label_lists_from(labels, x.shape[0])
            self._tr = self._trainer()
            loss, _, grads = self._tr.loss_and_grads(x, lens, lists)
            self._grads = grads
            return torch.tensor(loss)

        def train(self, loss, optimizer, learning_rate):
            from tensorflow_end2end_speech_recognition_b200.compat import graph as _g
            if _g.is_handle(loss) or _g.is_handle(learning_rate):
                return _g.Op(self.train, (loss, optimizer, learning_rate), {}, name="train")
            assert optimizer in ("adam", "rmsprop", "sgd", "momentum", "adagrad", "adadelta", "nestrov")
            from oracle import optim as oopt
            grads = [oopt.clip_by_norm(g, self.clip_grad_norm) if self.clip_grad_norm else g for g in self._grads]
            for v, g in zip(self._variables, grads):
                v.tensor -= float(learning_rate) * torch.tensor(g, dtype=torch.float32)     # plain SGD: plumbing check
            self.steps_run += 1
            return None

        def decoder(self, logits, inputs_seq_len, beam_width=1):
            from tensorflow_end2end_speech_recognition_b200.compat import graph as _g
            if _g.is_handle(logits) or _g.is_handle(inputs_seq_len):
                return _g.Op(self.decoder, (logits, inputs_seq_len, beam_width), {}, name="decoder")
            lg = np.transpose(logits.numpy(), (1, 0, 2))
            hyp = odec.greedy_decode(lg, np.asarray(inputs_seq_len), self.num_classes - 1)
            idx = [(b, j) for b, h in enumerate(hyp) for j in range(len(h))]
            val = [c for h in hyp for c in h]
            return SparseTensorValue(np.asarray(idx, np.int64).reshape(-1, 2), np.asarray(val, np.int32),
                                     np.asarray([len(hyp), max([len(h) for h in hyp] + [0])], np.int64))

        def compute_ler(self, decode_op, labels):
            from tensorflow_end2end_speech_recognition_b200.compat import graph as _g
            if _g.is_handle(decode_op) or _g.is_handle(labels):
                return _g.Op(self.compute_ler, (decode_op, labels), {}, name="compute_ler")
            from tensorflow_end2end_speech_recognition_b200.utils.io.labels.sparsetensor import sparse_to_label_lists
            B = int(decode_op.dense_shape[0])
            hyp = sparse_to_label_lists(decode_op, B)
            ref = ctc_mod.label_lists_from(labels, B)
            return odec.label_error_rate(hyp, ref)
    return OracleCTC


def test_reference_timit_train_ctc_do_train_runs_unchanged(tmp_path, capsys):
    from tensorflow_end2end_speech_recognition_b200 import compat
    stubs = {}
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "setproctitle"):
        if name not in sys.modules:
            stubs[name] = mock.MagicMock()
    ds_mod = types.ModuleType("examples.timit.data.load_dataset_ctc")
    ds_mod.Dataset = SyntheticDataset
    stubs["examples.timit.data.load_dataset_ctc"] = ds_mod
    tf = compat.install(reference_root=REF)
    try:
        with mock.patch.dict(sys.modules, stubs):
            spec = importlib.util.spec_from_file_location("ref_timit_train_ctc",
                                                          os.path.join(REF, "examples/timit/training/train_ctc.py"))
            script = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(script)              # the reference file itself, top to bottom
            assert script.CTC.__module__.startswith("tensorflow_end2end_speech_recognition_b200")
            OracleCTC = _oracle_backed_ctc()
            params = dict(label_type="phone61", batch_size=4, num_epoch=2, splice=1, num_stack=1, num_skip=1,
                          sort_stop_epoch=1, optimizer="adam", learning_rate=1e-2, dropout=0.0, beam_width=1,
                          decay_start_epoch=5, decay_rate=0.5, decay_patient_epoch=1, print_step=2,
                          eval_start_epoch=99, not_improved_patient_epoch=3)
            model = OracleCTC(encoder_type="blstm", input_size=12, num_units=8, num_layers=1, num_classes=9,
                              parameter_init=0.1, clip_grad_norm=5.0)
            model.save_path = str(tmp_path)
            before = model.flat_params.clone()
            script.do_train(model=model, params=params)   # the reference's training loop, unchanged
            out = capsys.readouterr().out
            assert model.steps_run == 6                   # 2 epochs x 3 mini-batches
            assert "Step 2" in out and "EPOCH:2" in out and "Total" in out
            assert os.path.isfile(os.path.join(str(tmp_path), "complete.txt"))
            assert float((model.flat_params - before).abs().max()) > 0
    finally:
        compat.uninstall()
        for k in ("ref_timit_train_ctc",):
            sys.modules.pop(k, None)
