"""Drop-in for pruning/methods.topk (pruning/methods.py:5-8): keep the `num_samples` images with the
highest mean CER, ties in dictionary order (Python's sort is stable), as a dict in that order.

The ranking is one launch of the segmented top-k kernel (qeb_cer_topk_segmented, stable descending) over the whole
dataset. The kernel compares fp32 keys; the reference compares Python floats. The two orders are identical as long as no
two DIFFERENT values collapse to the same fp32 number (the reference's artifacts are rounded to 3 decimals) - that is
checked, not assumed. `facility_location` needs the apricot package and is not part of the hot path.
"""
import numpy as np

from .. import selection_utils
from ... import _lib


def topk(cer_means, num_samples):
    names = list(cer_means.keys())
    vals = np.fromiter(cer_means.values(), dtype=np.float64, count=len(names))
    v32 = vals.astype(np.float32)
    if len(np.unique(v32)) != len(np.unique(vals)):
        raise _lib.QebError("pruning.topk: two different CER values are equal in fp32; the device ranking would not match the "
                            "reference's float comparison")
    k = max(0, min(int(num_samples), len(names)))
    if k == 0:
        return {}
    idx = selection_utils.topk_cer_indices(v32, k)
    return {names[i]: cer_means[names[i]] for i in idx.tolist()}
