"""AddGaussianNoice of the reference's transform_helper.py:26-45 on the device, plus the batched add_noise.

Same constructor and call signature. Per image: r_std = randint(0, std+1)/100 (stochastic) or std/100, + 1e-13;
out = clamp(image - noise_coef * N(mean, r_std), 0, 1). The per-image sigma draw stays on the host generator
(torch.randint, as the reference); the normal deviates come from the in-kernel Philox stream (qeb_gauss_jitter),
seeded from the host generator, so runs are reproducible under torch.manual_seed. The reference's own noise
values (CPU mt19937 stream) are not reproduced - only their distribution; `apply_noise` restates the arithmetic
bit-exactly on a given noise tensor.
"""
import torch

from .. import _lib


def _launch(img2d, sigma, mean, coef, noise_in, seed, want_noise, out=None, seed_dev=None):
    n_img, hw = img2d.shape
    if out is None:
        out = torch.empty_like(img2d)
    noise = torch.empty_like(img2d) if want_noise else None
    if seed_dev is not None:   # Philox key read from device memory when the kernel runs (CUDA-graph safe)
        _lib.call("qeb_gauss_jitter_devseed", img2d.data_ptr(), _lib.ptr(sigma), float(mean), float(coef), seed_dev.data_ptr(),
                  int(seed), n_img, hw, out.data_ptr(), _lib.ptr(noise), _lib.stream())
    else:
        _lib.call("qeb_gauss_jitter", img2d.data_ptr(), _lib.ptr(sigma), float(mean), float(coef), _lib.ptr(noise_in),
                  int(seed), n_img, hw, out.data_ptr(), _lib.ptr(noise), _lib.stream())
    return out, noise


def _check(images):
    if not images.is_cuda or images.dtype != torch.float32:
        raise _lib.QebError("qeb jitter needs CUDA fp32 images (no CPU fallback)")


def _to_device(images):
    """The reference trainers hand the noiser CPU tensors (`img_preds.detach().cpu()`, train_nn_area.py:225,236;
    `text_crops.detach().cpu()`, train_nn_patch.py:260,270): move them to the current CUDA device for the kernel and
    remember where the result has to go back to. No CPU arithmetic happens here - without a CUDA device this raises."""
    if images.is_cuda:
        return images, None
    if not torch.cuda.is_available():
        raise _lib.QebError("qeb jitter needs a CUDA device (no CPU fallback)")
    return images.to(torch.device("cuda", torch.cuda.current_device()), torch.float32, non_blocking=True), images.device


def apply_noise(images, noise, noise_coef=1):
    """clamp(images - noise_coef*noise, 0, 1) on the device for a given noise tensor (transform_helper.py:40-41)."""
    _check(images)
    x = images.contiguous()
    z = noise.to(x.device, torch.float32).contiguous()
    out, _ = _launch(x.view(1, -1), None, 0.0, noise_coef, z.view(1, -1), 0, False)
    return out.view_as(images)


def jitter_batch(images, sigmas, mean=0.0, noise_coef=1, seed=None, return_noise=False, out=None, seed_dev=None):
    """images (N,...) CUDA fp32; sigmas (N) per-image std (host or device). One launch for the whole batch -
    the fused form of the add_noise loops at train_nn_patch.py:187-191 / train_nn_area.py:184-191.
    out: optional preallocated result (same shape, contiguous). seed_dev: optional device tensor holding one uint64 (as
    int64) Philox key that is read when the kernel RUNS (key = *seed_dev + seed): a captured CUDA graph then draws new
    noise per replay once the caller updates that tensor; `seed` is an offset in that mode (default 0)."""
    _check(images)
    x = images.contiguous()
    n = x.shape[0]
    sg = torch.as_tensor(sigmas, dtype=torch.float32)
    if not sg.is_cuda:
        sg = sg.pin_memory().to(x.device, non_blocking=True)
    if seed is None:
        seed = 0 if seed_dev is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
    if out is not None and (out.shape != images.shape or not out.is_contiguous() or out.device != x.device):
        raise _lib.QebError("jitter_batch: `out` must be a contiguous tensor of the input's shape on its device")
    out, noise = _launch(x.view(n, -1), sg.contiguous(), mean, noise_coef, None, seed, return_noise,
                         None if out is None else out.view(n, -1), seed_dev)
    out = out.view_as(images)
    return (out, noise.view_as(images)) if return_noise else out


class AddGaussianNoice(object):
    def __init__(self, std=5, mean=0, is_stochastic=False, return_noise=False):
        self.std = std
        self.mean = mean
        self.is_stochastic = is_stochastic
        self.return_noise = return_noise

    def _sigmas(self, n):
        if self.is_stochastic:
            r = torch.randint(low=0, high=self.std + 1, size=(n,)).to(torch.float64) / 100.0
        else:
            r = torch.full((n,), self.std / 100.0, dtype=torch.float64)
        return (r + 0.0000000000001).to(torch.float32)

    def __call__(self, image, noise_coef=1):
        """One image (C,H,W) like the reference call; a CPU image (what the unmodified trainers pass) is processed on the
        device and handed back on the CPU."""
        image, home = _to_device(image)
        out = jitter_batch(image.unsqueeze(0), self._sigmas(1), self.mean, noise_coef, return_noise=self.return_noise)
        if self.return_noise:
            img, noise = out[0].squeeze(0), out[1].squeeze(0)
            return (img, noise) if home is None else (img.to(home), noise.to(home))
        out = out.squeeze(0)
        return out if home is None else out.to(home)

    def batch(self, images, noise_coef=1):
        """Whole batch (N,C,H,W), one sigma per image, one launch."""
        images, home = _to_device(images)
        out = jitter_batch(images, self._sigmas(images.shape[0]), self.mean, noise_coef, return_noise=self.return_noise)
        if home is None:
            return out
        return tuple(o.to(home) for o in out) if self.return_noise else out.to(home)


def crnn_on_noised(crnn_model, imgs, noiser, noise_coef=1, seed_dev=None, out=None):
    """`scores = crnn_model(add_noise(imgs, noiser))` in one pass (CRNN.forward_jittered): the jitter is generated inside
    conv1's input load. Returns (scores, noisy_imgs) or (scores, noisy_imgs, noise) with noiser.return_noise - the noisy batch
    is still materialised for the OCR engine (train_nn_area.py:260-262, train_nn_patch.py:289-292)."""
    _check(imgs)
    return crnn_model.forward_jittered(imgs, noiser._sigmas(imgs.shape[0]), noiser.mean, noise_coef, seed_dev=seed_dev,
                                       return_noise=noiser.return_noise, out=out)


def add_noise(imgs, noiser, noise_coef=1):
    """TrainNNPrep.add_noise (train_nn_area.py:184-191 returns (imgs, noise); train_nn_patch.py:187-191 imgs only,
    selected by noiser.return_noise)."""
    return noiser.batch(imgs, noise_coef)
