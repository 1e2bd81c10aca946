"""Input format of the path (reference loaders/data_generator.py:118-125) and the on-device synthetic pair
generators (reference loaders/data_generator.py:39-87)."""
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


GENERATORS_ON_DEVICE = {"ErdosRenyi": 0, "Regular": 1}


def generate_pairs_on_device(generator, num_pairs, n_vertices, edge_density, noise, seed=3787, sizes=None, device="cuda"):
    """`num_pairs` (graph, noisy copy) pairs drawn on the GPU: the reference's GENERATOR_FUNCTIONS[generator](edge_density, n)
    followed by noise_erdos_renyi(noise) (reference loaders/data_generator.py:39-87), as two CUDA uint8 (G,N,N) adjacency
    batches -- feed them to Network.forward_fused_adjacency / adjacency_batch_to_tensor_representation.  `sizes`
    (CUDA int32, optional) gives ragged vertex counts inside the N = n_vertices padding.  Same seed -> same graphs;
    parity with the reference's networkx generators is distributional, not bitwise."""
    from .. import _ops
    if generator not in GENERATORS_ON_DEVICE:
        raise NotImplementedError(f"generator {generator} is not built on the device (ErdosRenyi, Regular)")
    return _ops.generate_pairs(GENERATORS_ON_DEVICE[generator], int(num_pairs), int(n_vertices), float(edge_density),
                               float(noise), int(seed), sizes, device)
