"""OCR hand-off: the black-box OCR engines (Tesseract, EasyOCR, ...) run on the host and take PIL images
the reference builds with torchvision's `ToPILImage()(imgs[i])` from a CPU float tensor (ocr_helper/tess_helper.py:20-24),
after a blocking `n_text_crops.cpu()` of the fp32 batch (train_nn_patch.py:261-296, train_nn_area.py:244-262).

Here the (noisy) crops are converted to the same uint8 pixels on the device (qeb_to_uint8: `pic.mul(255).byte()`), copied
to a pinned host buffer asynchronously on a side stream - a quarter of the bytes, and the training stream is not
blocked - and handed to the OCR callable as uint8 arrays when the host asks for them:

    handoff = OcrHandoff(depth=2)
    t = handoff.submit(noisy_imgs)            # returns at once; D2H runs beside the next kernels
    ...                                        # e.g. the forward pass of the next inner iteration
    labels = handoff.labels(t, ocr)            # waits for the copy, ocr.get_labels on uint8 images

`OcrFromUint8` adapts a reference OCR helper (an object with get_labels(float tensor)) so that it receives the pixels it
would have produced itself.
"""
import numpy as np
import torch

from .. import _lib


def to_uint8(images):
    """(N,1,H,W) or (N,H,W) CUDA fp32 in [0,1] -> uint8 tensor of the same shape on the device (ToPILImage pixels)."""
    if not images.is_cuda or images.dtype != torch.float32:
        raise _lib.QebError("qeb to_uint8 needs a CUDA fp32 tensor (no CPU fallback)")
    x = images.contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    _lib.call("qeb_to_uint8", x.data_ptr(), x.numel(), out.data_ptr(), _lib.stream())
    return out


class OcrHandoff:
    """Ring of `depth` pinned host buffers; submit() = convert + asynchronous D2H on a side stream, fetch() = wait + view."""

    def __init__(self, depth=2):
        self.depth = depth
        self._slots = [None] * depth          # (pinned buffer, event, shape)
        self._next = 0
        self._stream = None

    def submit(self, images):
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=images.device)
        u8 = to_uint8(images)                                   # on the caller's stream, right behind the producer
        slot = self._next % self.depth
        self._next += 1
        buf = self._slots[slot][0] if self._slots[slot] is not None else None
        if buf is None or buf.numel() < u8.numel():
            buf = torch.empty(u8.numel(), dtype=torch.uint8).pin_memory()
        self._stream.wait_stream(torch.cuda.current_stream(images.device))
        with torch.cuda.stream(self._stream):
            buf[: u8.numel()].copy_(u8.view(-1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        u8.record_stream(self._stream)
        self._slots[slot] = (buf, ev, tuple(u8.shape))
        return self._next - 1

    def fetch(self, ticket):
        """uint8 numpy array (N,H,W) (a view of the pinned buffer: consume it before `depth` more submits)."""
        if ticket < self._next - self.depth or ticket >= self._next:
            raise _lib.QebError(f"OcrHandoff: ticket {ticket} is no longer (or not yet) in the ring")
        buf, ev, shape = self._slots[ticket % self.depth]
        ev.synchronize()
        n = int(np.prod(shape))
        arr = buf[:n].numpy().reshape(shape)
        return arr.reshape(shape[0], shape[-2], shape[-1])

    def labels(self, ticket, ocr):
        """ocr: callable uint8 array (N,H,W) -> list[str]."""
        return ocr(self.fetch(ticket))


class OcrFromUint8:
    """Adapter around a reference OCR helper (tess_helper.TessHelper, ...): its get_labels(imgs) expects a CPU float tensor
    and runs ToPILImage on every image; u8 / 255 is a float tensor whose ToPILImage pixels are exactly u8 again."""

    def __init__(self, helper):
        self.helper = helper

    def __call__(self, u8):
        return self.helper.get_labels(torch.from_numpy(np.ascontiguousarray(u8)).unsqueeze(1).float().div(255.0))
