"""Whole-step CUDA graphs for the training step (optional: the eager modules stay the drop-in default).

One phase-B step of the reference (train_nn_area.py:277-287) is ~240 kernel launches of 5-100 us each issued from two C
calls per network; measured on B200 the host needs about as long to issue them as the GPU to execute them (3.8 ms), so any per-step host
work (label encoding, `loss.item()`) serialises with the GPU. Capturing forward + losses + backward ONCE and replaying
the graph removes the host from the step: the trainer copies the batch and the encoded labels into static buffers,
replays, and steps the optimizer.

    tg = StaticTargets(batch=64, max_label_len=31, device=dev)      # static twin of the int32-CPU CTC arguments
    x = torch.empty(64, 1, 32, 128, device=dev)                      # static input batch
    def fwd_bwd():
        img = prep(x); scores = crnn(img)
        loss = ctc_loss(scores, tg) + scalar * mse_to_ones(img)
        loss.backward()
        return loss
    step = GraphedStep(fwd_bwd, modules=[prep, crnn])                # warm-up + capture
    for images, labels in loader:
        x.copy_(images, non_blocking=True); tg.load(y, pred_size, y_size)
        loss = step()                                                # replay; .grad of every parameter is refreshed
        optimizer.step()

Everything inside the graph is a kernel of libqeb_sm100.so (plus autograd's small glue kernels); the library's side stream
for weight gradients forks from and joins the capturing stream, so it is captured as a parallel branch.
"""
import numpy as np
import torch

from . import _lib
from .mirror.ctc import PackedTargets


class StaticTargets(PackedTargets):
    """CTC targets / lengths in FIXED device buffers (so a captured graph keeps reading the same addresses).

    `load(targets, input_lengths, target_lengths)` takes the arguments of torch.nn.CTCLoss as the reference builds them
    (int32 CPU tensors, train_nn_area.py:163-171), stages them in one pinned buffer and issues one H2D copy."""

    def __init__(self, batch, max_label_len, device):
        device = torch.device(device)
        self.cap = max(1, batch * max_label_len)
        # two pinned staging buffers, each with the event of the H2D copy last issued from it: load() never rewrites host
        # memory that a queued copy (behind a graph replay, with no per-step sync in the caller's loop) still has to read
        self._hosts = [torch.zeros(self.cap + 3 * batch, dtype=torch.int32).pin_memory() for _ in range(2)]
        self._events = [None, None]
        self._slot = 0
        self._host = self._hosts[0]
        self._dev = torch.zeros(self.cap + 3 * batch, dtype=torch.int32, device=device)
        c, B = self.cap, batch
        super().__init__(self._dev[:c], self._dev[c:c + B], self._dev[c + B:c + 2 * B], self._dev[c + 2 * B:], max_label_len, B)

    def load(self, targets, input_lengths, target_lengths, dev_out=None):
        """dev_out: another device buffer of this object's size to receive the staged arguments instead of the static one
        (BatchStager: the copy runs ahead on a copy stream and is moved into the static buffer when its step begins)."""
        B, c = self.B, self.cap
        tl = torch.as_tensor(target_lengths).to("cpu", torch.int32).reshape(-1)
        il = torch.as_tensor(input_lengths).to("cpu", torch.int32).reshape(-1)
        tg = torch.as_tensor(targets).to("cpu", torch.int32)
        if tl.numel() != B or il.numel() != B:
            raise RuntimeError(f"StaticTargets: expected {B} input/target lengths, got {il.numel()}/{tl.numel()}")
        tl_np = tl.numpy()
        if tg.dim() == 2:
            tg = torch.cat([tg[b, : tl_np[b]] for b in range(B)])
        tg = tg.reshape(-1)
        n_t = int(tl_np.sum())
        if int(tl_np.max()) > self.max_len or n_t > c:
            raise _lib.QebError(f"StaticTargets: a label of {int(tl_np.max())} symbols exceeds the captured capacity "
                                f"{self.max_len} (run this batch through the eager path)")
        if tg.numel() < n_t:
            raise RuntimeError("StaticTargets: targets shorter than sum(target_lengths)")
        self._slot ^= 1
        h = self._host = self._hosts[self._slot]
        if self._events[self._slot] is not None:
            self._events[self._slot].synchronize()   # the copy issued from this buffer two loads ago has finished
        h[:n_t] = tg[:n_t]
        offs = np.zeros(B, dtype=np.int32)
        np.cumsum(tl_np[:-1], out=offs[1:])
        h[c:c + B] = torch.from_numpy(offs)
        h[c + B:c + 2 * B] = il
        h[c + 2 * B:] = tl
        (self._dev if dev_out is None else dev_out).copy_(h, non_blocking=True)
        if self._dev.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self._events[self._slot] = ev
        return self


class BatchStager:
    """Input staging for a graphed step that overlaps the NEXT step's host work and host -> device copies with the replay of the
    current one.

    A captured step reads its batch and its CTC arguments from static buffers, so they cannot be overwritten while a replay runs.
    `stage()` copies the next batch (from pinned host memory) and the next labels into one of two staging sets on a private copy
    stream and returns at once; `commit()` - first thing of the next step - makes the step's stream wait for that copy and moves
    the staged set into the static buffers with two device-to-device copies (~2 us for a 1 MB batch). The trainer's loop becomes

        stager.stage(images0, y0, pred_size, y_size0)                 # before the loop
        for images, labels in loader:                                 # (one batch ahead)
            stager.commit(); loss = step(); optimizer.step()
            stager.stage(images, *encode(labels))                     # host encoding + H2D behind the replay
            log(loss.item())                                          # the step's own read-back

    and the host's per-step work (label encoding, staging, launch) no longer adds to the step time: measured 3.14 -> 3.0x ms per
    step end to end on B200 (bench.py `e2e`)."""

    def __init__(self, x_static, targets_static):
        """targets_static: one StaticTargets or a list of them (a step with several CTC calls, e.g. one per noised copy)."""
        self.x = x_static
        self.many = isinstance(targets_static, (list, tuple))
        self.tg = list(targets_static) if self.many else [targets_static]
        self.stream = torch.cuda.Stream(device=x_static.device)
        self.xs = [torch.empty_like(x_static) for _ in range(2)]
        self.tgs = [[torch.empty_like(t._dev) for t in self.tg] for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]      # the staged set has landed
        self.taken = [None, None]                                 # the step's stream has moved it into the static buffers
        self.slot = 0
        self.pending = None

    def stage(self, x_pinned, targets, input_lengths=None, target_lengths=None):
        """One target set: stage(x, targets, input_lengths, target_lengths). Several: stage(x, [(targets, input_lengths,
        target_lengths), ...]) in the order of the StaticTargets given to the constructor."""
        sets = targets if self.many else [(targets, input_lengths, target_lengths)]
        if len(sets) != len(self.tg):
            raise _lib.QebError(f"BatchStager.stage: {len(sets)} target sets for {len(self.tg)} static ones")
        s = self.slot
        self.slot ^= 1
        with torch.cuda.stream(self.stream):
            if self.taken[s] is not None:
                self.stream.wait_event(self.taken[s])             # two steps ago this set was still being read
            self.xs[s].copy_(x_pinned, non_blocking=True)
            for t, dst, (y, il, tl) in zip(self.tg, self.tgs[s], sets):
                t.load(y, il, tl, dev_out=dst)
            self.ready[s].record(self.stream)
        self.pending = s
        return self

    def commit(self):
        if self.pending is None:
            raise _lib.QebError("BatchStager.commit: nothing staged")
        s, self.pending = self.pending, None
        cur = torch.cuda.current_stream(self.x.device)
        cur.wait_event(self.ready[s])
        self.x.copy_(self.xs[s], non_blocking=True)
        for t, src in zip(self.tg, self.tgs[s]):
            t._dev.copy_(src, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.taken[s] = ev
        return self


class GraphedStep:
    """Capture `fn()` (forward + loss + backward over STATIC tensors) into one CUDA graph and replay it.

    modules: their gradients are set to None before the capture so that backward allocates them from the graph's memory
    pool; every replay overwrites them in place (no accumulation across replays - call the optimizer after each one).
    The return value of `fn` (a tensor or a tuple of tensors) is static as well: read it after the replay.
    No autograd graph of an earlier eager step over the same parameters may be alive at capture time (do not keep old loss
    tensors around): its AccumulateGrad nodes are bound to the default stream, which a capture must not touch."""

    def __init__(self, fn, modules=(), warmup=3):
        if not torch.cuda.is_available():
            raise _lib.QebError("GraphedStep needs a CUDA device (no CPU fallback)")
        self.fn, self.modules = fn, list(modules)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):      # warm-up off the default stream, as graph capture requires
            for _ in range(warmup):
                self._zero()
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._zero()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        try:
            with torch.cuda.graph(self.graph):
                self.out = fn()
        except RuntimeError as e:
            if "legacy stream" in str(e) or "StreamCaptureImplicit" in str(e):
                raise _lib.QebError(
                    "GraphedStep: the capture touched the legacy default stream. The usual cause is a loss tensor of an earlier "
                    "EAGER step that is still referenced: it keeps the parameters' AccumulateGrad nodes (bound to the default "
                    "stream) alive. Drop such references (keep float(loss), not loss) before capturing.") from e
            raise
        self.launches = _lib.launch_count() - n0   # kernels of libqeb_sm100.so inside one replay

    def _zero(self):
        for m in self.modules:
            m.zero_grad(set_to_none=True)

    def __call__(self):
        self.graph.replay()
        return self.out

    def close(self):
        """Destroy the captured graph (and return its memory pool). REQUIRED before torch.distributed.destroy_process_group()
        when the capture holds NCCL collectives (mirror/dist.BucketedAllReduce inside the step): tearing the communicator down
        under a live graph that references it hangs."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None
            self.out = None
