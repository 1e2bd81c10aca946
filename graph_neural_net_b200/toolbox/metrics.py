"""Alignment accuracy metrics (reference toolbox/metrics.py:92-141).

accuracy_max is computed on the device in the same pass as the loss (row argmax == row index);
accuracy_linear_assignment stays on the host like the reference (scipy Hungarian per graph) -- it is
a 'next' row of the scope table, not part of the CUDA hot path.
"""
import numpy as np
import torch

from .. import _ops
from .losses import _as_batch


def accuracy_max(weights, labels=None, aggregate_score=True):
    """(#rows whose arg-max column is the row index, #rows) or the per-graph accuracies."""
    if labels is not None:
        raise NotImplementedError("fgnn_b200 fuses the identity labelling only (labels=None)")
    scores, n_dev, sizes = _as_batch(weights)
    with torch.no_grad():
        _, correct = _ops.CrossEntropyIdentityFunction.apply(scores.detach(), n_dev)
    correct = correct.cpu().numpy()
    sizes = sizes.cpu().numpy().astype(np.int64)
    if aggregate_score:
        return int(correct.sum()), int(sizes.sum())
    return [c / s for c, s in zip(correct, sizes)]


def accuracy_linear_assignment(rawscores, labels=None, aggregate_score=True):
    """Hungarian matching accuracy.  log_softmax is a per-row shift, which does not change the optimal
    assignment, so the raw scores go straight to scipy after ONE device-to-host copy of the batch."""
    from scipy.optimize import linear_sum_assignment
    scores, n_dev, sizes = _as_batch(rawscores)
    host = scores.detach().cpu().numpy()
    sizes = sizes.cpu().numpy().astype(np.int64)
    acc, total, per = 0, 0, []
    for i, n in enumerate(sizes):
        label = labels[i] if labels else np.arange(n)
        _, preds = linear_sum_assignment(-host[i, :n, :n].astype(np.float64))
        hit = int(np.sum(preds == label))
        acc += hit
        total += int(n)
        per.append(hit / n)
    return (acc, total) if aggregate_score else per
