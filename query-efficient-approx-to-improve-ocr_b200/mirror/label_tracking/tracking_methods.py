"""Drop-in for the loss-weight generators of label_tracking/tracking_methods.py that sit on the hot path.

  LevenshteinWeightGenerator.gen_weights  tracking_methods.py:72-101
  DecayingWeightGenerator.gen_weights     tracking_methods.py:105-116
  weightgenerator_factory                 tracking_methods.py:119-125
The reference calls Levenshtein.distance once per ordered pair of history labels of every image (window^2 C calls per
image); here all unordered pairs of the whole minibatch go through ONE launch of the batched Levenshtein kernel
(qeb_levenshtein_batch) and the integer distances are combined on the host with the reference's float arithmetic, so
the weights are bit-identical. AttentionWeightGenerator (a trained HistoryAttention model) is outside the hot path
(out of scope).
"""
import torch

from .. import utils as qutils
from ... import _lib


class LevenshteinWeightGenerator:
    def __init__(self, tracking_args, device, char_to_index=None):
        self.args = tracking_args
        self.window_size = tracking_args.window_size
        self.device = device

    def print_debug_statements(self):
        pass

    def gen_weights(self, tracked_labels, img_names):
        """(len(img_names), window_size + 1) weights: column 0 = 1 (current label), column i + 1 = 0.5 * (1 - min(mean
        distance of history label i to the other history labels, len) / len)."""
        hist_multiplier = 0.5
        histories, a, b, where = [], [], [], []
        for img_index, name in enumerate(img_names):
            if name not in tracked_labels:
                histories.append(None)
                continue
            h = tracked_labels[name][-self.window_size:][::-1]
            histories.append(h)
            for i in range(len(h)):
                for j in range(i + 1, len(h)):       # the distance is symmetric: unordered pairs only
                    a.append(h[i]); b.append(h[j]); where.append((img_index, i, j))
        dist = qutils.levenshtein_strings(a, b)[0] if a else []
        sums = {}
        for (img_index, i, j), d in zip(where, dist):
            sums[(img_index, i)] = sums.get((img_index, i), 0) + int(d)
            sums[(img_index, j)] = sums.get((img_index, j), 0) + int(d)
        loss_weights = torch.zeros(len(img_names), self.window_size + 1)
        loss_weights[:, 0] = 1
        for img_index, h in enumerate(histories):
            if h is None:
                continue
            num_elements = max((len(h) - 1), 1)
            for i in range(len(h)):
                num_chars = max(1, len(h[i]))
                dist_mean = sums.get((img_index, i), 0) / num_elements
                loss_weights[img_index][i + 1] = hist_multiplier * (1 - min(dist_mean, num_chars) / num_chars)
        return loss_weights.to(self.device)


class DecayingWeightGenerator:
    def __init__(self, tracking_args, device, char_to_index=None):
        self.decay_factor = tracking_args.decay_factor
        self.window_size = tracking_args.window_size
        self.device = device

    def print_debug_statements(self):
        pass

    def gen_weights(self, training_obj, img_names):
        return torch.tensor([self.decay_factor ** i for i in range(0, self.window_size)]).to(self.device)


def weightgenerator_factory(method):
    weight_method_mapping = {
        "levenshtein": LevenshteinWeightGenerator,
        "decaying": DecayingWeightGenerator,
    }
    if method not in weight_method_mapping:
        raise _lib.QebError(f"qeb label tracking implements the 'levenshtein' and 'decaying' weight generators; '{method}' "
                            "(a trained attention model) is outside the hot path")
    return weight_method_mapping[method]
