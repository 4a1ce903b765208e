"""Label-tracking CTC path: mirror/tracking_utils.py and mirror/label_tracking/tracking_methods.py against
fixtures written by the UNMODIFIED reference (oracle/gen_golden.py gen_tracking: tracking_utils.py:34-75,
label_tracking/tracking_methods.py:72-116)."""
import json
import os
import types

import numpy as np
import pytest
import torch

from conftest import CHAR_SET, GOLDEN


def _fixture():
    g = np.load(os.path.join(GOLDEN, "tracking.npz"))
    j = json.load(open(os.path.join(GOLDEN, "tracking.json")))
    return g, j


def _trainer(j, mode, ctc_mod, device):
    c2i = {c: i for i, c in enumerate(CHAR_SET)}
    return types.SimpleNamespace(window_size=j["window"], tracked_labels=j["tracked"], char_to_index=c2i, device=device,
                                 weightgen_method=mode, primary_loss_fn=ctc_mod.CTCLoss(),
                                 primary_loss_fn_sample_wise=ctc_mod.CTCLoss(reduction="none"))


def test_target_batches_and_decaying_weights_host_logic():
    """Pure host logic: no GPU, no library call."""
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import tracking_utils as tu
    from qeb_b200.mirror.label_tracking import tracking_methods as tm
    g, j = _fixture()
    obj = _trainer(j, "levenshtein", torch.nn, "cpu")
    tb = tu.generate_ctc_target_batches(obj, j["names"][:16])
    assert len(tb) == len(j["target_batches"])
    for (t, ts, idx), (t_ref, ts_ref, idx_ref) in zip(tb, j["target_batches"]):
        assert t.dtype == torch.int32 and ts.dtype == torch.int32
        assert t.tolist() == t_ref and ts.tolist() == ts_ref and idx == idx_ref
    gen = tm.weightgenerator_factory("decaying")(types.SimpleNamespace(window_size=j["window"], decay_factor=0.7), "cpu")
    assert np.array_equal(gen.gen_weights(obj, j["names"]).numpy(), g["decay_weights"])
    hist = {}
    tu.add_labels_to_history(types.SimpleNamespace(tracked_labels=hist), ["a", "b", "a"], ["x", "y", "z"])
    assert hist == {"a": ["x", "z"], "b": ["y"]}
    with pytest.raises(Exception):
        tm.weightgenerator_factory("self_attention")


@pytest.mark.gpu
def test_levenshtein_weights_bit_exact():
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror.label_tracking import tracking_methods as tm
    g, j = _fixture()
    gen = tm.weightgenerator_factory("levenshtein")(types.SimpleNamespace(window_size=j["window"]), "cuda")
    w = gen.gen_weights(j["tracked"], j["names"])
    assert w.is_cuda and w.dtype == torch.float32
    assert np.array_equal(w.cpu().numpy(), g["lev_weights"])          # integer distances, same float arithmetic


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["levenshtein", "decaying"])
def test_weighted_ctc_loss_matches_reference(mode):
    """Multi-target CTC over ONE log-prob tensor through the batch_index argument of the CTC kernels (no gather copies):
    value and gradient at the scores against the reference's gathered torch-CPU computation."""
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import ctc as qctc, tracking_utils as tu
    g, j = _fixture()
    obj = _trainer(j, mode, qctc, "cuda")
    scores = torch.from_numpy(g["scores"]).cuda().requires_grad_(True)
    B = scores.shape[1]
    pred_size = torch.tensor([scores.shape[0]] * B, dtype=torch.int)
    tb = tu.generate_ctc_target_batches(obj, j["names"][:B])
    lw = torch.from_numpy(g["lev_weights"]).cuda()[:B, 1:] if mode == "levenshtein" else torch.from_numpy(g["decay_weights"]).cuda()
    from qeb_b200 import _lib
    n0 = _lib.launch_count()
    loss = tu.weighted_ctc_loss(obj, scores, pred_size, tb, lw)          # every history depth in ONE CTC launch
    n_fwd = _lib.launch_count() - n0
    loss.backward()
    ref_loss, ref_grad = float(g[f"loss_{mode}"]), torch.from_numpy(g[f"grad_{mode}"])
    assert abs(float(loss) - ref_loss) <= 1e-4 * abs(ref_loss)          # north_star: CTC within 1e-3 relative
    err = float((scores.grad.cpu() - ref_grad).norm() / ref_grad.norm())
    assert err < 1e-3, err
    # the reference's loop (one CTC call per depth) gives the same value and gradient with len(tb) times the launches
    s2 = scores.detach().clone().requires_grad_(True)
    n0 = _lib.launch_count()
    loss2 = tu.weighted_ctc_loss_per_depth(obj, s2, pred_size, tb, lw)
    n_loop = _lib.launch_count() - n0
    loss2.backward()
    assert len(tb) > 1 and n_fwd <= 2 and n_loop >= len(tb) * 1 and n_loop >= 2 * n_fwd   # one CTC launch (+ one reduction) against one set per depth
    assert abs(float(loss) - float(loss2)) <= 2e-6 * abs(float(loss2))
    assert float((scores.grad - s2.grad).abs().max()) <= 1e-6 * float(s2.grad.abs().max()) + 1e-9
    # columns of images without a label at some history depth receive no gradient from that depth (scatter, not overwrite)
    assert torch.isfinite(scores.grad).all()
