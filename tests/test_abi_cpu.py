"""The C-ABI library loads on a CPU-only box and exports every symbol include/qeb.h declares."""
import os
import re

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "qeb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qeb_\w+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "qeb_ctc_fwd" in syms and "qeb_levenshtein_batch" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol(qeb):
    lib = qeb._lib.load()
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/qeb.h but not exported"


def test_ctypes_signatures_cover_header(qeb):
    assert sorted(qeb._lib.SIGNATURES) == declared_symbols()


def test_error_reporting_without_gpu(qeb):
    # argument validation happens before any launch, so it is exercisable without a device
    lib = qeb._lib.load()
    assert lib.qeb_abi_version() == 1
    rc = lib.qeb_log_softmax_fwd(None, None, 0, 0, None)
    assert rc == -1 and b"log_softmax_fwd" in lib.qeb_last_error()
    rc = lib.qeb_levenshtein_batch(None, None, None, None, None, None, 5, 3, 10, None, None, None, None)
    assert rc == -1


def test_missing_gpu_is_loud(qeb):
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("has GPU")
    from qeb_b200.mirror import utils

    with pytest.raises(qeb.QebError):
        utils.compare_labels(["a"], ["b"])
