"""Declarative description of the 2-FGNN embedder as nested dicts (consumed by models.utils.Network).

Mirrors the reference's models/blocks_emb.py:9-43: node names ('in', 'mlp1', 'mlp2', 'mult', 'cat',
'mlp3', 'bm', 'block<i>', 'suffix') define the module attribute names and hence the state-dict keys
node_embedder.ne_bm_block{i}_mlp{j}.*.
"""
from .layers import MlpBlock_Real, ColumnMaxPooling, Concat, Identity, Matmul


def _mlp(c_in, c_out, depth, cst):
    return MlpBlock_Real(c_in, c_out, depth, constant_n_vertices=cst)


def block_emb(in_features, out_features, depth_of_mlp, constant_n_vertices=True):
    return {'in': Identity(),
            'mlp3': _mlp(in_features, out_features, depth_of_mlp, constant_n_vertices)}


def block(in_features, out_features, depth_of_mlp, constant_n_vertices=True):
    """x -> mlp3(cat[mlp1(x) @ mlp2(x), x])"""
    spec = {'in': Identity()}
    spec['mlp1'] = (_mlp(in_features, out_features, depth_of_mlp, constant_n_vertices), ['in'])
    spec['mlp2'] = (_mlp(in_features, out_features, depth_of_mlp, constant_n_vertices), ['in'])
    spec['mult'] = (Matmul(), ['mlp1', 'mlp2'])
    spec['cat'] = (Concat(), ['mult', 'in'])
    spec['mlp3'] = _mlp(in_features + out_features, out_features, depth_of_mlp, constant_n_vertices)
    return spec


def base_model(original_features_num, num_blocks, in_features, out_features, depth_of_mlp, block=block,
               constant_n_vertices=True):
    widths = [original_features_num] + [in_features] * (num_blocks - 1) + [out_features]
    spec = {'in': Identity()}
    for i in range(num_blocks):
        spec['block' + str(i + 1)] = block(widths[i], widths[i + 1], depth_of_mlp,
                                           constant_n_vertices=constant_n_vertices)
    return spec


def node_embedding(original_features_num, num_blocks, in_features, out_features, depth_of_mlp,
                   block=block, constant_n_vertices=True, **kwargs):
    return {'in': Identity(),
            'bm': base_model(original_features_num, num_blocks, in_features, out_features, depth_of_mlp,
                             block, constant_n_vertices=constant_n_vertices),
            'suffix': ColumnMaxPooling()}
