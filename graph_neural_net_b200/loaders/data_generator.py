"""Input format of the path (reference loaders/data_generator.py:118-125)."""
import torch


def adjacency_matrix_to_tensor_representation(W):
    """B[0] = W, B[1] = diag(deg)."""
    n = len(W)
    B = torch.zeros((2, n, n), dtype=W.dtype if W.is_floating_point() else torch.float32)
    B[0] = W
    B[1].diagonal().copy_(W.sum(1))
    return B


def adjacency_batch_to_tensor_representation(adj, sizes=None):
    """Device-side form of adjacency_matrix_to_tensor_representation for a whole (padded) batch: `adj` is a CUDA
    (G,N,N) uint8/bool tensor, `sizes` an optional CUDA int32 tensor of per-graph vertex counts.  Returns the
    (G,2,N,N) float32 input of the embedder (zero outside each graph's n x n block)."""
    from .. import _ops
    return _ops.features_from_adjacency(adj, sizes)
