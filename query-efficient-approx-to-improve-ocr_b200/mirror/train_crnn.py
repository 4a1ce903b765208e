"""The CRNN warm-up loop of the reference's train_crnn.py (TrainCRNN.train, :148-215) as two callables over the mirror
modules (BASELINE.json configs[0]): one training epoch (:157-166) and one validation pass (:168-183).

Same statements in the same order - `model.zero_grad(); scores, y, pred_size, y_size = _call_model(images, labels);
loss = CTCLoss()(scores, y, pred_size, y_size); loss.backward(); optimizer.step()` - with the label encoding of
`_call_model` (:137-145). The device work is the qeb CRNN (fused log-softmax head), the qeb CTC kernels and the one-launch
Adam; validation decodes from the head's arg-max path and scores all pairs of a batch with one Levenshtein launch. The
dataset / DataLoader, the StepLR scheduler (:133-135, :206) and the checkpointing (:186-190, :208-213) stay the trainer's.
"""
import torch

from . import ctc as qctc
from . import utils as qutils


def call_model(model, images, labels, char_to_index, device):
    """TrainCRNN._call_model (train_crnn.py:137-145)."""
    X_var = images.to(device)
    scores = model(X_var)
    out_size = torch.tensor([scores.shape[0]] * images.shape[0], dtype=torch.int)
    y_size = torch.tensor([len(l) for l in labels], dtype=torch.int)
    conc_label = "".join(labels)
    y = [char_to_index[c] for c in conc_label]
    y_var = torch.tensor(y, dtype=torch.int)
    return scores, y_var, out_size, y_size


def train_epoch(model, loader, optimizer, char_to_index, device, loss_function=None, log_every=0, epoch=0):
    """train_crnn.py:154-166. Returns (summed training loss, steps)."""
    loss_function = loss_function or qctc.CTCLoss()
    model.train()
    step, training_loss = 0, 0.0
    for images, labels in loader:
        model.zero_grad()
        scores, y, pred_size, y_size = call_model(model, images, labels, char_to_index, device)
        loss = loss_function(scores, y, pred_size, y_size)
        loss.backward()
        optimizer.step()
        training_loss += loss.item()
        if log_every and step % log_every == 0:
            print(f"Epoch: {epoch}, Iteration: {step} => {loss.item()}")
        step += 1
    return training_loss, step


@torch.no_grad()
def validate(model, loader, char_to_index, index_to_char, device, loss_function=None):
    """train_crnn.py:168-183. Returns (validation loss sum, exact-match count, summed CER, batches)."""
    loss_function = loss_function or qctc.CTCLoss()
    model.eval()
    validation_loss, pred_correct_count, pred_CER, label_count = 0.0, 0, 0, 0
    for images, labels in loader:
        scores, y, pred_size, y_size = call_model(model, images, labels, char_to_index, device)
        loss = loss_function(scores, y, pred_size, y_size)
        preds = qutils.pred_to_string(scores, labels, index_to_char)
        crt, cer = qutils.compare_labels(preds, list(labels))
        pred_correct_count += crt
        pred_CER += cer
        label_count += 1
        validation_loss += loss.item()
    return validation_loss, pred_correct_count, pred_CER, label_count
