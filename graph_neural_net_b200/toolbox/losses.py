"""Row-softmax cross-entropy against the identity matching, one fused CUDA pass for the batch.

Mirror of the reference toolbox/losses.py:8-34: triplet_loss(loss_reduction='mean'|'mean_of_mean').
The reference loops over graphs in Python and launches one CrossEntropyLoss per graph; here the
whole (B,N,N) batch (dense or MaskedTensor) goes through fgnn_ce_argmax_fwd_f32, which returns the
per-graph CE sums; the reduction over graphs is a handful of scalar ops.
"""
import torch
import torch.nn as nn

from .. import _ops
from ..maskedtensors.maskedtensor import MaskedTensor


def _as_batch(raw_scores):
    """-> (scores (B,N,N) plain tensor, sizes int32 device tensor or None, float sizes (B,))."""
    if isinstance(raw_scores, MaskedTensor):
        t = raw_scores.tensor.rename(None)
        n_dev = raw_scores.sizes_i32(t.device)
        return t, n_dev, n_dev.to(torch.float32)
    if isinstance(raw_scores, (list, tuple)):
        raise TypeError("pass a (B,N,N) tensor or a MaskedTensor (build one with maskedtensor.from_list)")
    sizes = torch.full((raw_scores.shape[0],), float(raw_scores.shape[-1]), device=raw_scores.device)
    return raw_scores, None, sizes


class triplet_loss(nn.Module):
    def __init__(self, loss_reduction='mean', loss=None):
        super().__init__()
        if loss is not None and not (isinstance(loss, nn.CrossEntropyLoss) and loss.reduction == 'sum'):
            raise NotImplementedError("only CrossEntropyLoss(reduction='sum') is fused")
        if loss_reduction not in ('mean', 'mean_of_mean'):
            raise ValueError('Unknown loss_reduction parameters {}'.format(loss_reduction))
        self.loss_reduction = loss_reduction

    def forward(self, raw_scores):
        """raw_scores: (bs, n, n) tensor or MaskedTensor -> scalar."""
        scores, n_dev, sizes = _as_batch(raw_scores)
        ce, _ = _ops.CrossEntropyIdentityFunction.apply(scores, n_dev)
        if self.loss_reduction == 'mean':
            return ce.sum() / sizes.sum()
        return (ce / sizes).sum() / ce.shape[0]
