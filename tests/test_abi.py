"""The C-ABI library loads without a GPU and exports every symbol the header
declares; the ctypes prototypes cover the header one to one."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b2asr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from tensorflow_end2end_speech_recognition_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    assert sorted(_lib.PROTOTYPES) == syms
    assert lib.b2_version() >= 100


def test_missing_library_fails_loudly(monkeypatch):
    from tensorflow_end2end_speech_recognition_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb2asr.so")
    try:
        _lib.load()
        assert False, "expected RuntimeError"
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)


def test_ops_refuse_cpu_tensors():
    import torch
    from tensorflow_end2end_speech_recognition_b200 import ops
    try:
        ops.softmax_rows(torch.zeros(2, 3))
        assert False
    except RuntimeError as e:
        assert "CUDA" in str(e)
