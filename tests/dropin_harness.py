"""TEST INFRASTRUCTURE - runs the UNMODIFIED reference trainers end to end on a synthetic dataset with a
fake OCR engine, either on the reference's own modules (CPU torch: the integration oracle) or on the qeb mirror bound by
qeb_b200.dropin (the drop-in under test), and dumps what the run produced as JSON.

    python tests/dropin_harness.py --trainer area|patch --impl ref|qeb --out result.json [--std 0] [--tracking]

One process per run (the swap rebinds names inside the reference's modules, so the two implementations never share an
interpreter). The reference tree is /root/reference in the build container and the git-ignored copy baseline/_ref on the GPU
box (oracle/refload.py).
"""
import argparse
import json
import os
import random
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORDS = ["total", "Cash", "12.50", "TAX", "item", "Qty 2", "VISA", "change", "0.99", "Thank you", "receipt", "No.", "A", "7",
         "subtotal", "milk", "BREAD", "2 x 3.25", "card", "date", "time", "store", "amount", "due", "paid", "ref", "tel",
         "open", "close", "items"]


class FakeOCR:
    """`ocr.get_labels(imgs) -> list[str]`, `count_calls` (what the trainers use of an OCR helper, ocr_helper/tess_helper.py).
    The labels do not depend on pixel values, so both implementations see the same label stream."""

    def __init__(self):
        self.count_calls = 0

    def get_labels(self, imgs):
        n = len(imgs)
        out = [WORDS[(self.count_calls * 7 + 3 * i) % len(WORDS)] for i in range(n)]
        self.count_calls += 1
        return out


def make_area_data(base, n_train, n_dev, seed=0):
    import numpy as np
    from PIL import Image

    rng = np.random.RandomState(seed)
    names = {}
    for split, n in (("vgg_train", n_train), ("vgg_dev", n_dev)):
        d = os.path.join(base, split)
        os.makedirs(d, exist_ok=True)
        for i in range(n):
            label = WORDS[(i * 5 + (0 if split == "vgg_train" else 11)) % len(WORDS)].replace(" ", "")
            w, h = int(rng.randint(40, 128)), int(rng.randint(16, 32))
            img = (255 - (rng.rand(h, w) < 0.15) * rng.randint(100, 255, size=(h, w))).astype(np.uint8)
            fn = f"{i}_{label}_{i * 3 + 1}.png"
            Image.fromarray(img, mode="L").save(os.path.join(d, fn))
            if split == "vgg_train":   # distinct values: the order of EQUAL CERs is unspecified on the reference
                names[fn] = round((float((i * 7) % n_train) + 0.5) / n_train, 6)
    return names


def make_patch_data(base, n_train, n_dev, seed=0):
    import numpy as np
    from PIL import Image

    rng = np.random.RandomState(seed)
    cers = {}
    for split, n in (("patch_dataset_train", n_train), ("patch_dataset_dev", n_dev)):
        folder = os.path.join(base, split, "setA", "docs")
        os.makedirs(folder, exist_ok=True)
        for k in range(n):
            img = (255 - (rng.rand(400, 512) < 0.1) * rng.randint(100, 255, size=(400, 512))).astype(np.uint8)
            Image.fromarray(img, mode="L").save(os.path.join(folder, f"{k}.png"))
            boxes = []
            for j in range(6 + k):
                label = WORDS[(j * 3 + k) % len(WORDS)]
                x0, y0 = int(rng.randint(0, 380)), int(rng.randint(0, 360))
                bw, bh = int(rng.randint(20, 120)), int(rng.randint(8, 30))
                boxes.append({"label": label, "x_min": x0, "y_min": y0, "x_max": x0 + bw, "y_max": y0 + bh})
                if split == "patch_dataset_train":
                    cers[f"{j}_{label}_docs_{k}"] = round((float((j * 5 + k * 3) % 23) + 0.25 * k + 0.5) / 23.0, 6)
            json.dump(boxes, open(os.path.join(folder, f"{k}.json"), "w"))
    return cers


def area_args(base, cers_path, a):
    return types.SimpleNamespace(
        batch_size=a.batch, lr_crnn=1e-4, lr_prep=5e-5, epoch=1, warmup_epochs=0, inner_limit=2, scalar=1.0, ocr="fake", std=a.std,
        random_std=a.std > 0, inner_limit_skip=a.tracking, crnn_model=None, prep_model=None, data_base_path=base,
        exp_base_path=os.path.join(base, "exp"), random_seed=42, minibatch_subset="topKCER", minibatch_subset_prop=0.5,
        start_epoch=0, train_subset_size=None, val_subset_size=None, lr_scheduler="cosine", exp_name="dropin", exp_id="0",
        cers_ocr_path=cers_path, weightgen_method="levenshtein", window_size=2, decay_factor=0.7)


def patch_args(base, cers_path, a):
    ns = area_args(base, cers_path, a)
    del ns.batch_size, ns.lr_scheduler
    ns.__dict__.update(weight_decay=5e-4, update_CRNN=False, image_prop=None, discount_factor=1, query_dim=8, emb_dim=8,
                       attn_activation="softmax", optim_crnn_path=None, optim_prep_path=None, pruning_artifact=None,
                       minibatch_subset="rangeCER" if a.range_sampler else "topKCER")
    return ns


def digest(module):
    import torch

    return {k: [list(v.shape), float(v.detach().double().norm().cpu())] for k, v in module.state_dict().items()
            if v.dtype in (torch.float32, torch.float64)}


def construct_only(a, t, work):
    """Construction, state_dict keys, whole-module torch.save / torch.load and optimizer state round trips under the swap."""
    import torch

    res = {"impl": a.impl, "device": str(t.device),
           "classes": {"crnn": type(t.crnn_model).__module__ + "." + type(t.crnn_model).__name__,
                       "prep": type(t.prep_model).__module__ + "." + type(t.prep_model).__name__,
                       "ctc": type(t.primary_loss_fn).__module__ + "." + type(t.primary_loss_fn).__name__,
                       "mse": type(t.secondary_loss_fn).__module__ + "." + type(t.secondary_loss_fn).__name__,
                       "optimizer": type(t.optimizer_prep).__module__ + "." + type(t.optimizer_prep).__name__,
                       "sampler": type(t.sampler).__module__ + "." + type(t.sampler).__name__},
           "state": {"crnn": {k: list(v.shape) for k, v in t.crnn_model.state_dict().items()},
                     "prep": {k: list(v.shape) for k, v in t.prep_model.state_dict().items()}},
           "crnn_digest": digest(t.crnn_model), "prep_digest": digest(t.prep_model)}
    for name, m in (("crnn", t.crnn_model), ("prep", t.prep_model)):
        path = os.path.join(work, name + ".pt")
        torch.save(m, path)                                   # whole-module pickles, as the trainers write them
        back = torch.load(path, weights_only=False)
        res[name + "_reload"] = {"class": type(back).__module__ + "." + type(back).__name__,
                                 "equal": all(torch.equal(x, y) for x, y in zip(back.state_dict().values(), m.state_dict().values()))}
        m.load_state_dict(back.state_dict())
    res["hook_registered"] = len(t.crnn_model._backward_hooks) == 1
    t.crnn_model.train(); t.crnn_model.apply(sys.modules["utils"].set_bn_eval)
    res["bn_eval_after_set_bn_eval"] = [m.training for m in t.crnn_model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    sd = t.optimizer_prep.state_dict()
    t.optimizer_prep.load_state_dict(sd)
    res["optimizer_groups"] = len(sd["param_groups"])
    json.dump(res, open(a.out, "w"))
    print("dropin_harness: construct-only done")


AREA_FLAGS = ["--batch_size", "8", "--lr_crnn", "0.0001", "--scalar", "1", "--lr_prep", "0.00005", "--epoch", "1",
              "--warmup_epochs", "0", "--std", "5", "--inner_limit", "2", "--inner_limit_skip", "--ocr", "fake", "--random_std",
              "--minibatch_subset", "topKCER", "--minibatch_subset_prop", "0.5", "--start_epoch", "0", "--exp_name", "dropin",
              "--exp_id", "7", "--random_seed", "42", "--weightgen_method", "levenshtein", "--window_size", "2", "--decay_factor", "0.7",
              "--lr_scheduler", "cosine"]
PATCH_FLAGS = ["--lr_crnn", "0.0001", "--scalar", "1", "--lr_prep", "0.00005", "--epoch", "1", "--random_seed", "42", "--std", "5",
               "--inner_limit", "2", "--inner_limit_skip", "--ocr", "fake", "--random_std", "--minibatch_subset", "topKCER",
               "--minibatch_subset_prop", "0.5", "--start_epoch", "0", "--exp_name", "dropin", "--exp_id", "7", "--warmup_epochs", "0",
               "--weight_decay", "0.0005", "--window_size", "2", "--weightgen_method", "decaying", "--decay_factor", "0.7",
               "--discount_factor", "1", "--query_dim", "8", "--emb_dim", "8", "--attn_activation", "softmax"]


def run_cli(a, trainer_mod, args, work):
    """The reference's own command line under the swap: area_cli.py:11-124 / patch_cli.py:11-155 parse their flags, call
    wandb.init and TrainNNPrep(args).train() (WANDB_MODE=disabled keeps wandb offline)."""
    import runpy

    os.environ["WANDB_MODE"] = "disabled"
    root = sys.modules["properties"].__file__.rsplit(os.sep, 1)[0]
    script = "area_cli.py" if a.trainer == "area" else "patch_cli.py"
    flags = list(AREA_FLAGS if a.trainer == "area" else PATCH_FLAGS)
    flags += ["--data_base_path", args.data_base_path, "--exp_base_path", args.exp_base_path, "--cers_ocr_path", args.cers_ocr_path]
    seen = {}
    if a.no_train:
        def fake_train(self):
            seen["trainer"] = self
            return 0.0, 0
        trainer_mod.TrainNNPrep.train = fake_train
    else:
        orig = trainer_mod.TrainNNPrep.train

        def train(self):
            seen["trainer"] = self
            return orig(self)
        trainer_mod.TrainNNPrep.train = train
    if a.trainer == "patch":   # patch_cli.py reads wandb_config.json from the working directory (patch_cli.py:160-163)
        json.dump({"mode": "disabled"}, open(os.path.join(work, "wandb_config.json"), "w"))
    sys.argv = [script] + flags
    ns = runpy.run_path(os.path.join(root, script), run_name="__main__")
    t = seen["trainer"]
    parsed = vars(ns["args"])
    res = {"impl": a.impl, "parsed": {k: (v if isinstance(v, (int, float, str, bool, type(None))) else str(v)) for k, v in parsed.items()},
           "n_flags": len(parsed), "device": str(t.device),
           "classes": {"crnn": type(t.crnn_model).__module__ + "." + type(t.crnn_model).__name__,
                       "prep": type(t.prep_model).__module__ + "." + type(t.prep_model).__name__},
           "params_file": os.path.exists(os.path.join(args.exp_base_path, sys.modules["properties"].param_path)),
           "cers": t.sampler.cers if not a.no_train else None}
    json.dump(res, open(a.out, "w"))
    print(f"dropin_harness: {script} done ({len(parsed)} parsed arguments)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trainer", choices=["area", "patch"], required=True)
    ap.add_argument("--impl", choices=["ref", "qeb"], required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--std", type=int, default=0)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--n-train", type=int, default=16)
    ap.add_argument("--n-dev", type=int, default=8)
    ap.add_argument("--tracking", action="store_true", help="inner_limit_skip: the label-tracking CTC path in the first inner iteration")
    ap.add_argument("--range-sampler", action="store_true")
    ap.add_argument("--construct-only", action="store_true", help="build TrainNNPrep under the swap, save / reload the modules, no training")
    ap.add_argument("--cli", action="store_true", help="run the reference's own area_cli.py / patch_cli.py (every flag given) "
                                                       "instead of constructing TrainNNPrep directly")
    ap.add_argument("--no-train", action="store_true", help="with --cli: TrainNNPrep.train becomes a no-op (flag parsing + construction)")
    a = ap.parse_args()
    a.out = os.path.abspath(a.out)   # the run changes into its own scratch directory

    os.environ["WANDB_MODE"] = "disabled"      # before wandb is imported: the CLIs call wandb.init(project=...) themselves
    import torch
    from oracle import refload

    ref = refload.load()                       # stubs for the nine missing third-party packages + the reference on sys.path
    import wandb
    if not a.cli:
        wandb.init(mode="disabled")
    mod_name = "train_nn_area" if a.trainer == "area" else "train_nn_patch"
    if a.impl == "qeb":
        import qeb_b200  # noqa: F401
        from qeb_b200 import dropin
        trainer_mod = dropin.import_trainer(mod_name)
    else:
        if torch.cuda.is_available():          # the integration oracle is the reference on CPU torch
            torch.cuda.is_available = lambda: False
        trainer_mod = __import__(mod_name)
    trainer_mod.get_ocr_helper = lambda name, **kw: FakeOCR()     # the "cached or synthetic labels" injection point (App. D.2)

    work = tempfile.mkdtemp(prefix="qeb_dropin_")
    os.chdir(work)
    base = os.path.join(work, "data")
    if a.trainer == "area":
        cers = make_area_data(base, a.n_train, a.n_dev)
    else:
        cers = make_patch_data(base, a.n_train, a.n_dev)
    cers_path = os.path.join(work, "cers.json")
    json.dump(cers, open(cers_path, "w"))
    os.makedirs(os.path.join(base, "exp"), exist_ok=True)
    args = area_args(base, cers_path, a) if a.trainer == "area" else patch_args(base, cers_path, a)

    if a.cli:
        return run_cli(a, trainer_mod, args, work)
    random.seed(42)
    t = trainer_mod.TrainNNPrep(args)
    if a.construct_only:
        return construct_only(a, t, work)
    rec = {"phase_b_loss": [], "phase_a_loss": []}
    orig_get_loss, orig_primary = t._get_loss, t.primary_loss_fn

    def get_loss(*args_, **kw):
        out = orig_get_loss(*args_, **kw)
        rec["phase_b_loss"].append(float(out))
        return out

    class Primary:   # records every CTC value the trainer computes directly (phase A); keeps attribute access intact
        def __call__(self, *args_, **kw):
            out = orig_primary(*args_, **kw)
            rec["phase_a_loss"].append(float(out))
            return out

        def __getattr__(self, name):
            return getattr(orig_primary, name)

    t._get_loss = get_loss
    if not a.tracking:   # the tracking path type-checks the loss object (qeb gather-free path), leave it untouched there
        t.primary_loss_fn = Primary()
    best_acc, best_epoch = t.train() if a.trainer == "area" else (t.train(), 0)

    ck = sorted(os.listdir(t.ckpt_base_path))
    prep_ck = [c for c in ck if c.startswith("Prep_model_0")]
    reloaded = torch.load(os.path.join(t.ckpt_base_path, prep_ck[0]), weights_only=False) if prep_ck else None
    result = {
        "impl": a.impl, "trainer": a.trainer, "device": str(t.device),
        "classes": {"crnn": type(t.crnn_model).__module__ + "." + type(t.crnn_model).__name__,
                    "prep": type(t.prep_model).__module__ + "." + type(t.prep_model).__name__,
                    "ctc": type(orig_primary).__module__ + "." + type(orig_primary).__name__,
                    "optimizer": type(t.optimizer_prep).__module__ + "." + type(t.optimizer_prep).__name__,
                    "sampler": type(t.sampler).__module__ + "." + type(t.sampler).__name__},
        "phase_a_loss": rec["phase_a_loss"], "phase_b_loss": rec["phase_b_loss"],
        "cers": t.sampler.cers, "all_cers": t.sampler.all_cers,
        "selected": {k: v for k, v in t.selected_samples.items()},
        "tracked_labels": getattr(t, "tracked_labels", None),
        "ocr_calls": t.ocr.count_calls,
        "ckpts": ck, "exp_files": sorted(os.listdir(t.cers_base_path)) + sorted(os.listdir(t.selectedsamples_path)),
        "reloaded_class": None if reloaded is None else type(reloaded).__module__ + "." + type(reloaded).__name__,
        "crnn_digest": digest(t.crnn_model), "prep_digest": digest(t.prep_model),
        "state_keys": {"crnn": list(t.crnn_model.state_dict().keys()), "prep": list(t.prep_model.state_dict().keys())},
    }
    if a.impl == "qeb":
        result["launches"] = qeb_b200.launch_count()
    json.dump(result, open(a.out, "w"))
    print(f"dropin_harness: {a.trainer}/{a.impl} done, phase B losses {rec['phase_b_loss'][:4]} ...")


if __name__ == "__main__":
    main()
