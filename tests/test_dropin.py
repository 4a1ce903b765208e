"""The drop-in, dropped in: the UNMODIFIED reference trainers (train_nn_area.TrainNNPrep, train_nn_patch.TrainNNPrep) and
command lines (area_cli.py, patch_cli.py) run on the qeb mirror bound by qeb_b200.dropin, compared with the same harness on
the reference's own modules ("the end-to-end drop-in parity test").

CPU tests (reference mounted at /root/reference or shipped as baseline/_ref): the swap binds, TrainNNPrep constructs, state_dict
keys / shapes / seed-42 values equal the reference's, whole-module pickles and optimizer state round-trip, every CLI flag parses.
GPU tests: one epoch of each trainer on the device against the reference run on the CPU of the same box.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

HARNESS = os.path.join(ROOT, "tests", "dropin_harness.py")
needs_ref = pytest.mark.skipif(not refload.available(), reason="reference tree not present (/root/reference or baseline/_ref)")


def run(tmp_path, *flags, timeout=900):
    out = os.path.join(str(tmp_path), "r_" + "_".join(f.strip("-") for f in flags if f.startswith("--"))[:80] + ".json")
    r = subprocess.run([sys.executable, HARNESS, "--out", out, *flags], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, f"harness failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return json.load(open(out))


@needs_ref
def test_swap_constructs_unmodified_trainer_cpu(tmp_path):
    q = run(tmp_path, "--trainer", "area", "--impl", "qeb", "--construct-only")
    r = run(tmp_path, "--trainer", "area", "--impl", "ref", "--construct-only")
    assert q["classes"] == {"crnn": "qeb_b200.mirror.models.model_crnn.CRNN", "prep": "qeb_b200.mirror.models.model_unet.UNet",
                            "ctc": "qeb_b200.mirror.ctc.CTCLoss", "mse": "qeb_b200.mirror.train_ops.MSELoss",
                            "optimizer": "qeb_b200.mirror.train_ops.Adam", "sampler": "qeb_b200.mirror.selection_utils.TopKCERSampler"}
    assert r["classes"]["crnn"] == "models.model_crnn.CRNN" and r["classes"]["ctc"] == "torch.nn.modules.loss.CTCLoss"
    # state_dict: same keys in the same order, same shapes, and - both seeded with 42 by set_random_seeds - the same values
    for net in ("crnn", "prep"):
        assert list(q["state"][net].items()) == list(r["state"][net].items())
        for k, (shape, norm) in r[net + "_digest"].items():
            assert q[net + "_digest"][k][0] == shape and abs(q[net + "_digest"][k][1] - norm) <= 1e-6 * max(1.0, norm), k
    assert q["crnn_reload"] == {"class": "qeb_b200.mirror.models.model_crnn.CRNN", "equal": True}
    assert q["prep_reload"] == {"class": "qeb_b200.mirror.models.model_unet.UNet", "equal": True}
    assert q["hook_registered"] and q["bn_eval_after_set_bn_eval"] == [False, False] and q["optimizer_groups"] == 1


@needs_ref
@pytest.mark.parametrize("trainer,n_flags", [("area", 28), ("patch", 36)])
def test_cli_flags_parse_under_swap_cpu(tmp_path, trainer, n_flags):
    """area_cli.py:11-124 / patch_cli.py:11-155 run unmodified: every flag parses, TrainNNPrep(args) constructs on the mirror."""
    q = run(tmp_path, "--trainer", trainer, "--impl", "qeb", "--cli", "--no-train", "--n-train", "8", "--n-dev", "8")
    assert q["n_flags"] == n_flags
    assert q["classes"] == {"crnn": "qeb_b200.mirror.models.model_crnn.CRNN", "prep": "qeb_b200.mirror.models.model_unet.UNet"}
    assert q["parsed"]["minibatch_subset"] == "topKCER" and q["parsed"]["inner_limit"] == 2 and q["params_file"]


def _close(a, b, rel):
    return abs(a - b) <= rel * max(1.0, abs(b))


ZERO_GRAD = ("convo.conv5.bias", "convo.conv6.bias")   # conv bias before a train-mode BatchNorm: the true gradient is zero, Adam
                                                       # integrates each implementation's own rounding noise


def _compare_epoch(q, r, first_rel=2e-4, later_rel=2e-3):
    assert q["device"].startswith("cuda") and r["device"] == "cpu"
    assert q["launches"] > 300, "the CUDA path did not run"
    assert q["state_keys"] == r["state_keys"]
    assert len(q["phase_a_loss"]) == len(r["phase_a_loss"]) and len(q["phase_b_loss"]) == len(r["phase_b_loss"])
    # first phase-A and phase-B values: same weights, same inputs -> kernel tolerance; later ones also carry one Adam step
    # whose first update is lr * sign(g) (a sign flip of a near-zero gradient moves a weight by 2 lr)
    assert _close(q["phase_a_loss"][0], r["phase_a_loss"][0], first_rel), (q["phase_a_loss"], r["phase_a_loss"])
    assert _close(q["phase_b_loss"][0], r["phase_b_loss"][0], later_rel), (q["phase_b_loss"], r["phase_b_loss"])
    for a, b in zip(q["phase_a_loss"] + q["phase_b_loss"], r["phase_a_loss"] + r["phase_b_loss"]):
        assert _close(a, b, later_rel), (q["phase_a_loss"], r["phase_a_loss"], q["phase_b_loss"], r["phase_b_loss"])
    assert q["selected"] == r["selected"]                       # TopKCER picks on the initial (distinct) CERs
    assert list(q["cers"].keys()) == list(r["cers"].keys()) and q["all_cers"].keys() == r["all_cers"].keys()
    same = sum(q["cers"][k] == r["cers"][k] for k in r["cers"])
    assert same >= 0.9 * len(r["cers"]), (same, len(r["cers"]))  # CER of an untrained surrogate's decode: equal strings
    # checkpoint names carry the validation accuracy of the (fake) OCR, which depends on the validation order and so on how many
    # host-RNG draws the jitter made (the reference draws its noise from the host generator, the kernel from Philox)
    strip = lambda names: sorted(n.rsplit("_", 1)[0] if n.startswith("Prep_model_0") else n for n in names if n != "Prep_model_best")
    assert q["ocr_calls"] == r["ocr_calls"] and strip(q["ckpts"]) == strip(r["ckpts"]) and q["exp_files"] == r["exp_files"]
    assert q["tracked_labels"] == r["tracked_labels"]
    assert q["reloaded_class"] == "qeb_b200.mirror.models.model_unet.UNet"
    for net in ("crnn_digest", "prep_digest"):                  # after the epoch: every tensor moved the same way
        for k, (shape, norm) in r[net].items():
            assert q[net][k][0] == shape and (k in ZERO_GRAD or _close(q[net][k][1], norm, 1e-3)), (net, k, q[net][k], norm)


@needs_ref
@pytest.mark.gpu
def test_area_trainer_epoch_matches_reference(tmp_path):
    """train_nn_area.TrainNNPrep.train() (train_nn_area.py:191-413), one epoch, noise std 0 so that both runs see the same
    inputs: phases A (TopKCER subset, jitter, CRNN step), B (UNet step through the frozen-BN surrogate), C (decode, CER,
    update_cer), validation, checkpoints and JSON side files."""
    r = run(tmp_path, "--trainer", "area", "--impl", "ref")
    q = run(tmp_path, "--trainer", "area", "--impl", "qeb")
    _compare_epoch(q, r)


@needs_ref
@pytest.mark.gpu
def test_patch_trainer_epoch_matches_reference(tmp_path):
    """train_nn_patch.TrainNNPrep.train() (train_nn_patch.py:193-467): 400x512 documents through the UNet, get_text_stack crops,
    accumulated inner-loop gradients, Adam with weight decay."""
    r = run(tmp_path, "--trainer", "patch", "--impl", "ref", "--n-train", "2", "--n-dev", "1")
    q = run(tmp_path, "--trainer", "patch", "--impl", "qeb", "--n-train", "2", "--n-dev", "1")
    _compare_epoch(q, r)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("trainer", ["area", "patch"])
def test_cli_epoch_with_noise_and_label_tracking(tmp_path, trainer):
    """The reference's own command line, every flag given, jitter std 5 and --inner_limit_skip (label-tracking CTC in the first
    inner iteration): runs to completion on the device and leaves the reference's artefacts."""
    q = run(tmp_path, "--trainer", trainer, "--impl", "qeb", "--cli", "--n-train", "8" if trainer == "area" else "2", "--n-dev",
            "8" if trainer == "area" else "1")
    assert q["device"].startswith("cuda") and q["params_file"]
    assert q["cers"] and all(isinstance(v, float) and v == v for v in q["cers"].values())
