"""Run the UNMODIFIED reference trainers / CLIs on the qeb hot path.

The reference (tataganesh/Query-Efficient-Approx-to-improve-OCR) has no plugin registry: its trainers import the hot
path by name (`from models.model_crnn import CRNN`, `from utils import compare_labels, pred_to_string, ...`,
`from transform_helper import AddGaussianNoice`, `from selection_utils import datasampler_factory`,
`from torch.nn import CTCLoss, MSELoss`, `import torch.optim as optim`: train_nn_area.py:1-30, train_nn_patch.py:1-33,
train_crnn.py:1-25). `install()` rebinds exactly those names on the reference's OWN modules, before the trainer module is
imported, so that no line of the reference has to be edited:

    import qeb_b200.dropin as dropin
    dropin.install("/path/to/Query-Efficient-Approx-to-improve-OCR")     # the reference checkout goes on sys.path
    import train_nn_area                                                  # unmodified
    dropin.patch_trainer(train_nn_area)                                   # CTCLoss / MSELoss / optim.Adam names
    train_nn_area.TrainNNPrep(args).train()

or, for the command lines (`patch_cli.py:9-175`, `area_cli.py:9-140`, flags untouched):

    python -m qeb_b200.dropin /path/to/reference area_cli.py --batch_size 64 --minibatch_subset topKCER ...

Everything that is not on the hot path (datasets, OCR helpers, wandb logging, checkpointing, JSON side files, PadWhite, the
attention weight generator) stays the reference's own code. Whole-module pickles keep loading: `models.model_crnn.CRNN` /
`models.model_unet.UNet` resolve to the mirror classes, which have the reference's state_dict keys.
"""
import importlib
import os
import sys

_installed = None


def install(reference_root=None, torch_losses=True):
    """Rebind the reference's hot-path names to the qeb mirror. reference_root: the reference checkout (put first on
    sys.path); None = it is importable already. Returns the dict {module name: [rebound attributes]}.
    Idempotent. Must run before `train_nn_area` / `train_nn_patch` / `train_crnn` are imported (they bind by name)."""
    global _installed
    if _installed is not None:
        return _installed
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    for late in ("train_nn_area", "train_nn_patch", "train_crnn"):
        if late in sys.modules:
            raise RuntimeError(f"qeb_b200.dropin.install(): {late} is already imported - install() has to run first "
                               f"(the trainers bind the hot-path names at import time)")
    from .mirror import selection_utils as q_sel
    from .mirror import tracking_utils as q_track
    from .mirror import transform_helper as q_th
    from .mirror import utils as q_utils
    from .mirror.label_tracking import tracking_methods as q_tm
    from .mirror.models import model_crnn as q_crnn
    from .mirror.models import model_unet as q_unet
    from .mirror.pruning import methods as q_prune

    done = {}

    def rebind(mod_name, source, names):
        mod = importlib.import_module(mod_name)
        for n in names:
            setattr(mod, n, getattr(source, n))
        done[mod_name] = list(names)
        return mod

    rebind("models.model_crnn", q_crnn, ["CRNN", "Convolutional"])
    rebind("models.model_unet", q_unet, ["UNet"])
    rebind("utils", q_utils, ["compare_labels", "pred_to_string", "get_text_stack", "set_bn_eval", "get_char_maps"])
    rebind("transform_helper", q_th, ["AddGaussianNoice"])
    rebind("selection_utils", q_sel, ["datasampler_factory", "DataSampler", "RandomSampler", "CerRangeSampler", "TopKCERSampler",
                                      "UniformSamplerGlobal", "RandomSamplerGlobal"])
    rebind("tracking_utils", q_track, ["call_crnn", "generate_ctc_label", "generate_ctc_target_batches", "weighted_ctc_loss",
                                       "add_labels_to_history"])
    try:
        tm = importlib.import_module("label_tracking.tracking_methods")
        for n in ("LevenshteinWeightGenerator", "DecayingWeightGenerator"):
            setattr(tm, n, getattr(q_tm, n))
        ref_factory = tm.weightgenerator_factory

        def weightgenerator_factory(name):   # attention (a trained HistoryAttention model) stays the reference's class
            if name in ("levenshtein", "decaying"):
                return q_tm.weightgenerator_factory(name)
            return ref_factory(name)

        tm.weightgenerator_factory = weightgenerator_factory
        done["label_tracking.tracking_methods"] = ["LevenshteinWeightGenerator", "DecayingWeightGenerator", "weightgenerator_factory"]
    except ImportError:
        pass
    try:
        rebind("pruning.methods", q_prune, ["topk"])
    except ImportError:
        pass
    _installed = done
    return done


def patch_trainer(module, adam=True):
    """The three torch classes the trainers name explicitly (train_nn_area.py:7-8,146-154; train_nn_patch.py:8-9,143-152;
    train_crnn.py:130-134): `CTCLoss`, `MSELoss` (module globals) and `optim.Adam`. Called on the imported trainer module."""
    from .mirror import ctc as q_ctc
    from .mirror import train_ops as q_ops

    if hasattr(module, "CTCLoss"):
        module.CTCLoss = q_ctc.CTCLoss
    if hasattr(module, "MSELoss"):
        module.MSELoss = q_ops.MSELoss
    if adam and hasattr(module, "optim"):
        class _Optim:   # `optim.Adam` -> qeb Adam; every other attribute (lr_scheduler, ...) is torch.optim's
            def __init__(self, base):
                self._base = base
                self.Adam = q_ops.Adam

            def __getattr__(self, name):
                return getattr(self._base, name)

        module.optim = _Optim(module.optim)
    return module


def import_trainer(name, reference_root=None):
    """install() + import + patch_trainer() of `train_nn_area` / `train_nn_patch` / `train_crnn`."""
    install(reference_root)
    return patch_trainer(importlib.import_module(name))


def main(argv=None):
    """python -m qeb_b200.dropin <reference_root> <area_cli.py | patch_cli.py> [the CLI's own flags ...]"""
    import runpy

    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        raise SystemExit(main.__doc__)
    root, script = os.path.abspath(argv[0]), argv[1]
    install(root)
    trainer = {"area_cli.py": "train_nn_area", "patch_cli.py": "train_nn_patch"}.get(os.path.basename(script))
    if trainer:
        import_trainer(trainer)
    sys.argv = [script] + argv[2:]
    runpy.run_path(os.path.join(root, script) if not os.path.isabs(script) else script, run_name="__main__")


if __name__ == "__main__":
    main()
