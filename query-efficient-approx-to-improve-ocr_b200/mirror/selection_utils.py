"""Samplers of the reference's selection_utils.py with the same names, factory keys and query/update_cer API.

  DataSampler.update_cer   selection_utils.py:70-77    host dict bookkeeping (restated)
  TopKCERSampler.query     selection_utils.py:144-151  -> qeb_cer_topk_segmented
  CerRangeSampler.query    selection_utils.py:107-135  -> qeb_cer_range_segmented (torch.rand draws stay on the host
                                                          generator, so the reference's random stream is kept)
  RandomSampler / *Global  selection_utils.py:80-98, 172-217  host logic, restated (no kernel needed)
Tie order of equal CERs is lowest-index-first (= torch.argsort(stable=True)); the reference's unstable argsort
leaves it undefined.  query_segmented() scores many minibatches in one launch.
"""
import random
from abc import ABCMeta, abstractmethod

import numpy as np
import torch

from .. import _lib


def _device_of(images):
    if torch.is_tensor(images) and images.is_cuda:
        return images.device
    if not torch.cuda.is_available():
        raise _lib.QebError("the qeb samplers need a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


_WARP_SEGMENT_MAX = 1024   # the per-warp ranking kernel is quadratic in the segment size: larger segments are sorted


def topk_global_device(vals_dev, k):
    """Stable top-k of ONE device-resident fp32 vector of any size (qeb_cer_topk_global): int64 device indices."""
    n = vals_dev.numel()
    k = min(int(k), n)
    out = torch.empty(max(k, 0), dtype=torch.int64, device=vals_dev.device)
    if k <= 0:
        return out
    work = torch.empty(_lib.load().qeb_cer_topk_global_workspace_bytes(n), dtype=torch.uint8, device=vals_dev.device)
    _lib.call("qeb_cer_topk_global", vals_dev.data_ptr(), n, k, work.data_ptr(), out.data_ptr(), _lib.stream())
    return out


def _segmented(vals_list, ks, device, rands_list=None):
    """Run one launch over several segments. Returns list of int64 CPU index tensors."""
    n_seg = len(vals_list)
    if rands_list is None and any(len(v) > _WARP_SEGMENT_MAX for v in vals_list):
        # dataset-sized segments go through the sort kernel one by one, the minibatch-sized ones through the warp kernel
        small = [i for i, v in enumerate(vals_list) if len(v) <= _WARP_SEGMENT_MAX]
        res = [None] * n_seg
        if small:
            for i, r in zip(small, _segmented([vals_list[i] for i in small], [ks[i] for i in small], device)):
                res[i] = r
        for i, v in enumerate(vals_list):
            if res[i] is None:
                vd = torch.from_numpy(np.ascontiguousarray(np.asarray(v, dtype=np.float32))).pin_memory().to(device, non_blocking=True)
                res[i] = topk_global_device(vd, ks[i]).cpu()
        return res
    sizes = np.array([len(v) for v in vals_list], dtype=np.int32)
    ks = np.array(ks, dtype=np.int32)
    seg_off = np.zeros(n_seg + 1, dtype=np.int32)
    np.cumsum(sizes, out=seg_off[1:])
    picks = ks.copy() if rands_list is not None else np.minimum(ks, sizes)
    if rands_list is not None:
        picks = np.where(sizes > 0, picks, 0).astype(np.int32)
    out_off = np.zeros(n_seg + 1, dtype=np.int32)
    np.cumsum(picks, out=out_off[1:])
    total, n_out = int(seg_off[-1]), int(out_off[-1])
    if total == 0 or n_out == 0:
        return [torch.zeros(0, dtype=torch.long) for _ in range(n_seg)]
    vals = torch.from_numpy(np.concatenate([np.asarray(v, dtype=np.float32) for v in vals_list])).pin_memory().to(device, non_blocking=True)
    ints = torch.from_numpy(np.concatenate([seg_off, ks if rands_list is None else picks, out_off[:-1]])).pin_memory().to(device, non_blocking=True)
    d_off, d_k, d_oo = ints[: n_seg + 1], ints[n_seg + 1: 2 * n_seg + 1], ints[2 * n_seg + 1:]
    out = torch.empty(n_out, dtype=torch.int64, device=device)
    if rands_list is None:
        _lib.call("qeb_cer_topk_segmented", vals.data_ptr(), d_off.data_ptr(), d_k.data_ptr(), d_oo.data_ptr(), n_seg,
                  out.data_ptr(), _lib.stream())
    else:
        rands = torch.cat([r.reshape(-1).float()[: int(p)] for r, p in zip(rands_list, picks)]).pin_memory().to(device, non_blocking=True)
        work = torch.empty_like(vals)
        _lib.call("qeb_cer_range_segmented", vals.data_ptr(), d_off.data_ptr(), d_k.data_ptr(), d_oo.data_ptr(),
                  rands.data_ptr(), n_seg, work.data_ptr(), out.data_ptr(), None, _lib.stream())
    out = out.cpu()
    return [out[out_off[i]: out_off[i + 1]] for i in range(n_seg)]


def topk_cer_indices(cers_f32, k, device=None):
    """argsort(cers, descending, stable)[:k] on the device; cers = sequence of python floats / fp32 tensor."""
    v = torch.as_tensor(cers_f32, dtype=torch.float32).cpu().numpy()
    return _segmented([v], [k], device or _device_of(None))[0]


def range_cer_indices(cers_f32, rands, device=None):
    v = torch.as_tensor(cers_f32, dtype=torch.float32).cpu().numpy()
    return _segmented([v], [len(rands)], device or _device_of(None), rands_list=[torch.as_tensor(rands)])[0]


class DataSampler(metaclass=ABCMeta):
    def __init__(self, cers=dict()):
        self.cers = cers
        self.all_cers = dict()

    @abstractmethod
    def query(self):
        pass

    def update_cer(self, batch_cers, names):
        for name, cer in zip(names, batch_cers):
            if name not in self.cers:
                print(f"Sample not present - {name}")
            self.cers[name] = cer
            if name not in self.all_cers:
                self.all_cers[name] = list()
            self.all_cers[name].append(cer)

    def _gather_cers(self, names):
        # names missing from the dict are skipped, silently shifting indices, as in the reference (Appendix B.5)
        return [self.cers[name] for name in names if name in self.cers]


class RandomSampler(DataSampler):
    def __init__(self, cers=dict()):
        self.cers = cers
        self.all_cers = dict()

    def query(self, images, labels, num_samples, names=None):
        rand_indices = torch.randperm(images.shape[0])[:num_samples]
        return images[rand_indices], [labels[i] for i in rand_indices], rand_indices


class CerRangeSampler(DataSampler):
    def __init__(self, cers, discount_factor=1):
        self.cers = cers
        self.discount_factor = discount_factor
        self.all_cers = dict()

    def query(self, images, labels, num_samples, names):
        image_cers = self._gather_cers(names)
        selection_idx = torch.tensor([], dtype=torch.long)
        if len(image_cers) != 0:
            rands = torch.rand(num_samples)  # same host-generator draw as the reference
            selection_idx = _segmented([np.asarray(image_cers, dtype=np.float64).astype(np.float32)], [num_samples],
                                       _device_of(images), rands_list=[rands])[0]
        return images[selection_idx], [labels[i] for i in selection_idx], selection_idx


class TopKCERSampler(DataSampler):
    def __init__(self, cers, discount_factor=1):
        self.cers = cers
        self.discount_factor = discount_factor
        self.all_cers = dict()

    def query(self, images, labels, num_samples, names):
        image_cers = self._gather_cers(names)
        selection_idx = _segmented([np.asarray(image_cers, dtype=np.float64).astype(np.float32)], [num_samples],
                                   _device_of(images))[0]
        return images[selection_idx], [labels[i] for i in selection_idx], selection_idx

    def query_segmented(self, names_per_batch, num_samples_per_batch, device=None):
        """Selection indices for many minibatches in ONE launch (config 5: 15,625 minibatches of 64)."""
        vals = [np.asarray(self._gather_cers(n), dtype=np.float64).astype(np.float32) for n in names_per_batch]
        return _segmented(vals, num_samples_per_batch, device or _device_of(None))


class UniformSamplerGlobal(DataSampler):
    def __init__(self, cers, num_samples):
        self.cers = cers
        self.num_samples = num_samples
        self.selected_indices = np.zeros(num_samples, dtype=np.int32)
        self.selected_samplenames = dict()

    def select_samples(self):
        self.selected_samplenames.clear()
        cer_keys = list(self.cers.keys())
        cer_values = np.array(list(self.cers.values()))
        sorted_cer_indices = np.argsort(cer_values)
        for i, split in enumerate(np.array_split(sorted_cer_indices, self.num_samples)):
            self.selected_indices[i] = np.random.choice(split)
            self.selected_samplenames[cer_keys[self.selected_indices[i]]] = True

    def query(self, images, labels, num_samples=-1, names=None):
        selection_idx = torch.tensor([i for i, name in enumerate(names) if name in self.selected_samplenames]).long()
        return images[selection_idx], [labels[i] for i in selection_idx], selection_idx


class RandomSamplerGlobal(DataSampler):
    def __init__(self, cers, num_samples):
        self.cers = cers
        self.num_samples = num_samples
        self.selected_samplenames = dict()

    def select_samples(self):
        self.selected_samplenames.clear()
        for name in random.sample(list(self.cers.keys()), self.num_samples):
            self.selected_samplenames[name] = True

    def query(self, images, labels, num_samples=-1, names=None):
        selection_idx = torch.tensor([i for i, name in enumerate(names) if name in self.selected_samplenames]).long()
        return images[selection_idx], [labels[i] for i in selection_idx], selection_idx


def datasampler_factory(sampling_method):
    # "uniformEntropy" (selection_utils.py:155-169) is a dead path in the reference (constructed with the wrong
    # arity, train_nn_patch.py:76-78) and is not mirrored.
    method_mapping = {
        "random": RandomSampler,
        "topKCER": TopKCERSampler,
        "uniformCERglobal": UniformSamplerGlobal,
        "randomglobal": RandomSamplerGlobal,
        "rangeCER": CerRangeSampler,
    }
    return method_mapping[sampling_method]
