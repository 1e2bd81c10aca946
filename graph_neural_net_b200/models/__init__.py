"""Model factories (reference models/__init__.py:6-24)."""
from .trainers import Siamese_Node_Exp


def get_siamese_model_exp(args, config_optim):
    node_emb = args['node_emb']
    print('Fetching model %s with (total = %s ) init %s and inside %s' % (
        node_emb['type'], node_emb['num_blocks'], node_emb['block_init'], node_emb['block_inside']))
    return Siamese_Node_Exp(args['original_features_num'], node_emb, lr=config_optim['lr'],
                            scheduler_decay=config_optim['scheduler_decay'],
                            scheduler_step=config_optim['scheduler_step'])


def get_siamese_model_test(name, config=None):
    """Load a reference / Lightning checkpoint (state_dict keys are identical)."""
    import json
    import torch
    if config is None:
        split_name = name.split("/")[-4]
        with open(name.split(split_name)[0] + 'config.json') as f:
            config = json.load(f)
    model = Siamese_Node_Exp(2, dict(config['arch']['node_emb']))
    ckpt = torch.load(name, map_location='cpu')
    model.load_state_dict(ckpt.get('state_dict', ckpt))
    return model
