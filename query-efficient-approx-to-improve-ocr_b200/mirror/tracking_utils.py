"""Drop-in for the reference's tracking_utils.py (label-tracking CTC path).

  call_crnn                    tracking_utils.py:5-10
  generate_ctc_label           tracking_utils.py:34-39
  generate_ctc_target_batches  tracking_utils.py:42-56
  weighted_ctc_loss            tracking_utils.py:59-75
  add_labels_to_history        tracking_utils.py:77-81
Same names, `self`-first calling convention (the trainers call them as free functions on the trainer object) and
results. The one change in mechanism: the reference materialises `scores[:, img_indices, :]` for every history depth
(tracking_utils.py:65) before each CTC call; here the subset is an index vector handed to the CTC kernels
(qeb_ctc_fwd / qeb_ctc_bwd `batch_index`), which read those columns of the ONE log-prob tensor in place and scatter the
gradient back into it - no gather copies, no index_put in the backward.
"""
import torch

from . import ctc as qctc


def call_crnn(self, images):
    X_var = images.to(self.device)
    scores = self.crnn_model(X_var)
    out_size = torch.tensor([scores.shape[0]] * images.shape[0], dtype=torch.int)
    return scores, out_size


def generate_ctc_label(self, labels):
    y_size = torch.tensor([len(l) for l in labels], dtype=torch.int)
    conc_label = ''.join(labels)
    y = [self.char_to_index[c] for c in conc_label]
    y_var = torch.tensor(y, dtype=torch.int)
    return y_var, y_size


def generate_ctc_target_batches(self, img_names):
    """One (targets, target_sizes, image indices) triple per history depth i: the i-th most recent OCR label of every
    image whose history is at least i + 1 long."""
    target_batches = list()
    for i in range(self.window_size):
        batch_labels = list()
        img_indices = list()
        for j, name in enumerate(img_names):
            label_history = self.tracked_labels[name]
            if i < len(label_history):
                batch_labels.append(label_history[-(i + 1)])
                img_indices.append(j)
        if len(img_indices):
            target, target_size = generate_ctc_label(self, batch_labels)
            target_batches.append([target, target_size, img_indices])
    return target_batches


def _subset_ctc(loss_fn, scores, target, pred_size, target_size, img_indices):
    """loss_fn(scores[:, img_indices, :], target, pred_size[img_indices], target_size) without the gather when loss_fn
    is the qeb CTCLoss; any other loss object gets the reference's gathered call."""
    pred_size_subset = pred_size[img_indices]
    if isinstance(loss_fn, qctc.CTCLoss) and scores.is_cuda:
        return qctc.ctc_loss(scores, target, pred_size_subset, target_size, loss_fn.blank, loss_fn.reduction,
                             loss_fn.zero_infinity, batch_index=img_indices)
    return loss_fn(scores[:, img_indices, :], target, pred_size_subset, target_size)


def _all_depths_one_launch(self, scores, pred_size, target_batches, loss_weights, num_losses):
    """Every history depth in ONE CTC launch: the depths' subsets are concatenated into one batch of
    (column index, target) rows over the same log-prob tensor (qeb_ctc_fwd `batch_index`), the per-sample losses come back with
    reduction 'none' and the depths' means / weights become one coefficient per row:
      decaying:      sum_i w_i * mean_b(nll_b / max(1, len_b))      -> coef = w_i / (n_i * max(1, len_b))   (CTCLoss 'mean')
      per-sample:    sum_i mean_b(W[b, i] * nll_b)                  -> coef = W[b, i] / n_i
    The backward is one launch too (the coefficients are the per-row grad_out)."""
    dev = scores.device
    tgts, sizes, idx, coef = [], [], [], []
    for i in range(num_losses):
        target, target_size, img_indices = target_batches[i]
        n_i = len(img_indices)
        ts = torch.as_tensor(target_size).to("cpu", torch.int32).reshape(-1)
        tgts.append(torch.as_tensor(target).to("cpu", torch.int32).reshape(-1))
        sizes.append(ts)
        ii = torch.as_tensor(img_indices, dtype=torch.int64)
        idx.append(ii)
        if self.weightgen_method == "decaying":   # w_i / (n_i * max(1, len_b)): built where the weight lives, no host sync
            w_i = torch.as_tensor(loss_weights[i], dtype=torch.float32).to(dev, non_blocking=True)
            coef.append(w_i / (n_i * ts.clamp(min=1).to(torch.float32)).pin_memory().to(dev, non_blocking=True))
        else:
            lw = torch.as_tensor(loss_weights)
            coef.append(lw[ii.to(lw.device), i].to(dev, torch.float32) / n_i)
    idx = torch.cat(idx)
    nll = qctc.ctc_loss(scores, torch.cat(tgts), torch.as_tensor(pred_size)[idx], torch.cat(sizes), self.primary_loss_fn.blank,
                        "none", self.primary_loss_fn.zero_infinity, batch_index=idx)
    return (nll * torch.cat(coef)).sum()


def weighted_ctc_loss(self, scores, pred_size, target_batches, loss_weights):
    num_losses = min(len(target_batches), self.window_size)
    sample_wise = self.weightgen_method != "decaying"
    fused = isinstance(self.primary_loss_fn, qctc.CTCLoss) and scores.is_cuda and num_losses > 0 and \
        (not sample_wise or isinstance(getattr(self, "primary_loss_fn_sample_wise", None), qctc.CTCLoss)) and \
        self.primary_loss_fn.reduction == "mean"
    if fused:
        return _all_depths_one_launch(self, scores, pred_size, target_batches, loss_weights, num_losses)
    all_ctc_losses = list()
    for i in range(num_losses):
        target, target_size, img_indices = target_batches[i]
        if self.weightgen_method == "decaying":
            loss_weight = loss_weights[i]
            ctc_loss = _subset_ctc(self.primary_loss_fn, scores, target, pred_size, target_size, img_indices)
            all_ctc_losses.append(loss_weight * ctc_loss)
        else:
            loss_weights_subset = loss_weights[img_indices, i]
            ctc_losses = _subset_ctc(self.primary_loss_fn_sample_wise, scores, target, pred_size, target_size, img_indices)
            all_ctc_losses.append(torch.mean(loss_weights_subset * ctc_losses))
    return sum(all_ctc_losses)


def weighted_ctc_loss_per_depth(self, scores, pred_size, target_batches, loss_weights):
    """The reference's loop, one CTC call per history depth (kept for tests / A-B against the one-launch form)."""
    num_losses = min(len(target_batches), self.window_size)
    all_ctc_losses = list()
    for i in range(num_losses):
        target, target_size, img_indices = target_batches[i]
        if self.weightgen_method == "decaying":
            all_ctc_losses.append(loss_weights[i] * _subset_ctc(self.primary_loss_fn, scores, target, pred_size, target_size, img_indices))
        else:
            ctc_losses = _subset_ctc(self.primary_loss_fn_sample_wise, scores, target, pred_size, target_size, img_indices)
            all_ctc_losses.append(torch.mean(loss_weights[img_indices, i] * ctc_losses))
    return sum(all_ctc_losses)


def add_labels_to_history(self, image_keys, ocr_labels):
    for lbl_index, name in enumerate(image_keys):
        if name not in self.tracked_labels:
            self.tracked_labels[name] = list()
        self.tracked_labels[name].append(ocr_labels[lbl_index])
