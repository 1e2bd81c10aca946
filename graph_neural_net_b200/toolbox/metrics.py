"""Alignment accuracy metrics (reference toolbox/metrics.py:92-141).

accuracy_max is computed on the device in the same pass as the loss (row argmax == row index);
accuracy_linear_assignment (the metric training_step actually calls, trainers.py:53,74) runs a batched
shortest-augmenting-path solver on the device (csrc/fgnn_lap.cu) instead of one host scipy call per graph.
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


def linear_assignment(rawscores):
    """Optimal matchings of a batch on the device (fgnn_lap_fwd): -> (col_of_row (B,N) int32 with -1 in padded rows,
    correct (B,) int32, total_cost (B,) float64).  One CTA per graph runs the shortest-augmenting-path algorithm scipy
    uses, in double precision with scipy's tie rule, on cost = -log_softmax(scores) formed in the kernel."""
    import ctypes as C
    from .. import _lib as L
    scores, n_dev, _ = _as_batch(rawscores)
    scores = L.require_cuda_f32(scores.detach(), "rawscores")
    G, N, M = scores.shape
    if N != M:
        raise L.FgnnError("rawscores must be (B,N,N)")
    cols = torch.empty((G, N), dtype=torch.int32, device=scores.device)
    correct = torch.empty(G, dtype=torch.int32, device=scores.device)
    cost = torch.empty(G, dtype=torch.float64, device=scores.device)
    L.check(L.get_lib().fgnn_lap_fwd(L.ptr(scores), L.ptr(cols), L.ptr(correct), L.ptr(cost), G, N,
                                     _ops._npg(n_dev, G), L.stream_ptr(scores.device)), "fgnn_lap_fwd")
    return cols, correct, cost


def accuracy_linear_assignment(rawscores, labels=None, aggregate_score=True):
    """Hungarian matching accuracy (reference toolbox/metrics.py:92-116) computed on the device on
    -log_softmax(scores) as the reference does; only the per-graph hit counts (B int32) travel to the host.  `labels` (a list of per-graph index arrays) compares
    the device matching with them on the host, as the reference does."""
    scores, n_dev, sizes = _as_batch(rawscores)
    cols, correct, _ = linear_assignment(rawscores)
    sizes = sizes.cpu().numpy().astype(np.int64)
    if labels:
        cols_h = cols.cpu().numpy()
        hits = np.array([int(np.sum(cols_h[i, :n] == np.asarray(labels[i])[:n])) for i, n in enumerate(sizes)])
    else:
        hits = correct.cpu().numpy().astype(np.int64)
    if aggregate_score:
        return int(hits.sum()), int(sizes.sum())
    return [h / n for h, n in zip(hits, sizes)]
