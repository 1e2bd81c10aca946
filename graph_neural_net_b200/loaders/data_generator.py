"""Input format of the path (reference loaders/data_generator.py:118-125)."""
import torch


def adjacency_matrix_to_tensor_representation(W):
    """B[0] = W, B[1] = diag(deg)."""
    n = len(W)
    B = torch.zeros((2, n, n), dtype=W.dtype if W.is_floating_point() else torch.float32)
    B[0] = W
    B[1].diagonal().copy_(W.sum(1))
    return B
