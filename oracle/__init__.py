"""CPU oracle for the BLSTM / CTC / attention acoustic-training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline), never as the thing shipped.  The product path
(``tensorflow_end2end_speech_recognition_b200``) fails loudly when its CUDA
library is missing; it never routes through this package.

PARITY STATUS: **unpinned against TensorFlow**.  The reference
(hirofumi0810/tensorflow_end2end_speech_recognition) executes every FLOP of
this path inside TensorFlow 1.x kernels (``requirements.txt:11`` pins
``tensorflow==1.2.0``), TensorFlow is neither vendored in ``/root/reference``
nor installable here, and the reference's own tests hold no numeric golden
vectors (``models/test/test_ctc.py`` asserts nothing).  The restatement
therefore follows the reference call sites plus the published TF-1.x op
semantics and is pinned as far as this environment
allows:

* ``oracle.ctc``      - cross-checked against ``torch.nn.functional.ctc_loss``
                        (an independent implementation) and brute-force path
                        enumeration on tiny lattices (tests/test_oracle_ctc.py).
* ``oracle.decode``   - pinned bit-exact against the reference's own numpy
                        decoders ``models/ctc/decoders/{greedy,beam_search}_decoder.py``
                        (the one piece of this path that imports without TF);
                        golden vectors committed under ``tests/golden/`` by
                        ``tests/golden/make_golden.py``.
* ``oracle.lstm``     - follows the only in-tree statement of the cell
                        equations, ``models/recurrent/layers/lstm.py:142-183``;
                        two independent forms (numpy loop, torch autograd) are
                        cross-checked against each other.
* ``oracle.optim``    - TF-1.x update rules as recalled; cross-checked against
                        torch.optim where the rules coincide.
"""
