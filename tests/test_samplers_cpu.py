"""Host-side samplers of mirror/selection_utils.py that need no kernel: UniformSamplerGlobal,
RandomSamplerGlobal (selection_utils.py:172-217), RandomSampler (:80-98) and DataSampler.update_cer (:70-77) against the
UNMODIFIED reference classes under the same seeds (reference mounted at /root/reference or shipped as baseline/_ref), and
against their definition when the reference is absent."""
import random

import numpy as np
import pytest
import torch

from oracle import refload


def _cers(n, seed):
    rng = random.Random(seed)
    return {f"{i}_{rng.randrange(10 ** 6)}_img": rng.randrange(0, 12) / rng.randrange(1, 9) for i in range(n)}


def _mirror():
    import qeb_b200  # noqa: F401
    from qeb_b200.mirror import selection_utils
    return selection_utils


@pytest.mark.skipif(not refload.available(), reason="reference not mounted")
@pytest.mark.parametrize("n,k", [(1000, 37), (64, 64), (500, 1)])
def test_global_samplers_select_the_reference_samples(n, k):
    ref, mir = refload.load().selection_utils, _mirror()
    cers = _cers(n, n + k)
    for name in ("uniformCERglobal", "randomglobal"):
        a, b = ref.datasampler_factory(name)(dict(cers), k), mir.datasampler_factory(name)(dict(cers), k)
        for s in (a, b):
            np.random.seed(5); random.seed(5)
            s.select_samples()
        assert list(a.selected_samplenames) == list(b.selected_samplenames) and len(b.selected_samplenames) <= k
        if name == "uniformCERglobal":
            assert np.array_equal(a.selected_indices, b.selected_indices)
        # query: the minibatch members that were selected globally, in minibatch order
        names = list(cers)[7:7 + 64]
        images, labels = torch.arange(64.0).view(64, 1, 1, 1), [f"l{i}" for i in range(64)]
        ia, la, xa = a.query(images, labels, -1, names)
        ib, lb, xb = b.query(images, labels, -1, names)
        assert torch.equal(xa, xb) and la == lb and torch.equal(ia, ib)
        # a second draw replaces the selection (cleared, not accumulated)
        for s in (a, b):
            np.random.seed(6); random.seed(6)
            s.select_samples()
        assert list(a.selected_samplenames) == list(b.selected_samplenames)


def test_uniform_global_sampler_draws_one_sample_per_cer_quantile():
    mir = _mirror()
    cers = _cers(997, 3)
    s = mir.datasampler_factory("uniformCERglobal")(dict(cers), 50)
    np.random.seed(1)
    s.select_samples()
    vals = np.array(list(cers.values()))
    order = np.argsort(vals)
    splits = np.array_split(order, 50)
    assert len(s.selected_indices) == 50
    for i, sp in enumerate(splits):                      # selection_utils.py:179-187: one random pick per split of the argsort
        assert s.selected_indices[i] in sp
    keys = list(cers)
    assert set(s.selected_samplenames) == {keys[i] for i in s.selected_indices}
    picked = np.sort(vals[s.selected_indices])
    assert picked[0] <= np.quantile(vals, 0.05) and picked[-1] >= np.quantile(vals, 0.95)   # covers the CER range


def test_random_samplers_and_update_cer_bookkeeping():
    mir = _mirror()
    cers = _cers(200, 9)
    s = mir.datasampler_factory("randomglobal")(dict(cers), 20)
    random.seed(2)
    s.select_samples()
    random.seed(2)
    assert list(s.selected_samplenames) == random.sample(list(cers.keys()), 20)
    r = mir.datasampler_factory("random")(dict(cers))
    torch.manual_seed(4)
    images = torch.arange(10.0).view(10, 1, 1, 1)
    sub, labels, idx = r.query(images, list("abcdefghij"), 4)
    torch.manual_seed(4)
    assert torch.equal(idx, torch.randperm(10)[:4]) and torch.equal(sub, images[idx]) and labels == [list("abcdefghij")[i] for i in idx]
    t = mir.datasampler_factory("topKCER")(dict(cers))
    names = list(cers)[:3]
    t.update_cer([0.5, 0.25, 2.0], names)
    t.update_cer([0.75], names[:1])
    assert t.cers[names[0]] == 0.75 and t.all_cers[names[0]] == [0.5, 0.75] and t.all_cers[names[2]] == [2.0]
    with pytest.raises(KeyError):
        mir.datasampler_factory("uniformEntropy")       # dead path in the reference: not mirrored
