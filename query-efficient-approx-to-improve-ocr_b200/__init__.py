"""qeb_b200: B200-native (sm_100a) hot path of tataganesh/Query-Efficient-Approx-to-improve-OCR.

The directory is named `query-efficient-approx-to-improve-ocr_b200`; import it as `qeb_b200` through the
loader at the repository root (`import qeb_b200`). Host code is Python/PyTorch (device memory, streams,
torch.distributed); all compute goes through the C ABI of libqeb_sm100.so (include/qeb.h, csrc/*.cu).
There is no CPU fallback: importing works anywhere, calling an op without the library or a GPU raises.
"""
from . import _lib  # noqa: F401
from ._lib import QebError, launch_count, reset_launch_count  # noqa: F401

__all__ = ["QebError", "launch_count", "reset_launch_count", "build"]


def build(verbose=False, force=False):
    """Compile csrc/*.cu for sm_100a into libqeb_sm100.so (in-tree)."""
    from . import _build

    return _build.build(verbose=verbose, force=force)
