"""Device-side input pipeline -- frame stacking / skipping, splicing, zero padding and the
per-GPU split of ``DatasetBase.__next__`` (``utils/dataset/ctc.py:112-182``,
``utils/io/inputs/frame_stacking.py``, ``utils/io/inputs/splicing.py``).

The raw utterances are staged once into a pinned, zero-padded host buffer (two buffers used
alternately so that the copy of batch n+1 can overlap the step of batch n), copied to the
device, and expanded there by ONE gather kernel (``b2_stack_splice``) -- the stacked / spliced
features, up to ``num_stack * splice`` times larger than the raw ones, never cross PCIe.
"""
import math

import numpy as np
import torch

from .... import _lib, ops


def shard_bounds(batch_size, num_gpu):
    """[start, end) of every rank's slice under ``np.array_split`` (utils/dataset/ctc.py:171-177):
    the first ``batch_size % num_gpu`` ranks get one utterance more."""
    q, r = divmod(batch_size, num_gpu)
    out, s = [], 0
    for i in range(num_gpu):
        n = q + (1 if i < r else 0)
        out.append((s, s + n))
        s += n
    return out


def pad_labels(label_list, padded_value=-1):
    """labels [B, max_len] padded with ``padded_value`` (utils/dataset/ctc.py:137-139,164-165)"""
    Lm = max(len(l) for l in label_list)
    out = np.full((len(label_list), Lm), padded_value, np.int32)
    for b, l in enumerate(label_list):
        out[b, :len(l)] = l
    return out


class DeviceInputPipeline(object):
    def __init__(self, num_stack=1, num_skip=1, splice=1, num_gpu=1, device="cuda:0"):
        if num_stack != 1 and num_stack < num_skip:
            raise ValueError("num_skip must be less than num_stack.")
        self.num_stack, self.num_skip, self.splice = num_stack, num_skip, splice
        self.num_gpu = num_gpu
        self.device = torch.device(device)
        self._pinned = [None, None]
        self._turn = 0

    def out_len(self, raw_len):
        return raw_len if self.num_stack == 1 else int(math.ceil(raw_len / self.num_skip))

    def _stage(self, input_list):
        B = len(input_list)
        D = input_list[0].shape[1]
        Traw = max(x.shape[0] for x in input_list)
        need = B * Traw * D
        buf = self._pinned[self._turn]
        if buf is None or buf.numel() < need:
            buf = torch.empty(max(need, 1), dtype=torch.float32, pin_memory=True)
            self._pinned[self._turn] = buf
        self._turn ^= 1
        host = buf[:need].view(B, Traw, D)
        host.zero_()
        hn = host.numpy()
        for b, x in enumerate(input_list):
            hn[b, :x.shape[0]] = x
        lens = torch.tensor([x.shape[0] for x in input_list], dtype=torch.int32)
        return host, lens, (B, Traw, D)

    def __call__(self, input_list, rank=None):
        """list of [T_b, D] float32 arrays -> (inputs [B', T', D'] cuda, inputs_seq_len [B'] cuda int32).
        ``rank``: keep only that rank's ``np.array_split`` shard (padding stays the GLOBAL max, as
        in the reference, which pads before splitting)."""
        lib = _lib.load()
        Tout = max(self.out_len(x.shape[0]) for x in input_list)
        if rank is not None:
            s, e = shard_bounds(len(input_list), self.num_gpu)[rank]
            input_list = input_list[s:e]
        host, lens, (B, Traw, D) = self._stage(input_list)
        raw = host.to(self.device, non_blocking=True)
        raw_len = lens.to(self.device, non_blocking=True)
        Dout = lib.b2_stack_splice_out_dim(D, self.num_stack, self.splice)
        out = torch.empty((B, Tout, Dout), dtype=torch.float32, device=self.device)
        out_len = torch.empty(B, dtype=torch.int32, device=self.device)
        rc = lib.b2_stack_splice(ops._ptr(raw), ops._ptr(raw_len), B, Traw, D, self.num_stack, self.num_skip,
                                 self.splice, Tout, ops._ptr(out), ops._ptr(out_len), ops._stream())
        _lib.check(rc, "b2_stack_splice")
        return out, out_len


class PinnedPrefetcher(object):
    """Double-buffered host -> device staging on a side stream.

    ``put(arrays...)`` starts the asynchronous copy of the NEXT batch (pinned host tensors) while the
    current step computes; ``get()`` makes the compute stream wait for that copy and hands out the
    device tensors.  Two device buffer sets alternate, so a batch is never overwritten while the step
    that consumes it may still be running (the event recorded by ``get`` guards the reuse)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = [None, None]
        self._ready = [None, None]       # copy finished
        self._free = [None, None]        # last consumer finished
        self._turn = 0
        self._pending = None

    def put(self, *host_tensors):
        k = self._turn
        self._turn ^= 1
        if self._free[k] is not None:
            self.stream.wait_event(self._free[k])
        with torch.cuda.stream(self.stream):
            if self._slots[k] is None or any(d.shape != h.shape or d.dtype != h.dtype
                                             for d, h in zip(self._slots[k], host_tensors)):
                self._slots[k] = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_tensors]
            for d, h in zip(self._slots[k], host_tensors):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._ready[k] = ev
        self._pending = k

    def get(self):
        k = self._pending
        assert k is not None, "PinnedPrefetcher.get() without a preceding put()"
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready[k])
        self._pending = None
        self._last = k
        return self._slots[k]

    def release(self):
        """call after the step that consumed the last ``get()`` has been enqueued"""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[self._last] = ev
