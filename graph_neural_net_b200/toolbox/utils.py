"""Small helpers kept from the reference's toolbox/utils.py that the hot path touches."""


def get_device(t):
    """Device index of a CUDA tensor, 'cpu' otherwise (reference toolbox/utils.py:104-107)."""
    return t.get_device() if t.is_cuda else 'cpu'


def get_lr(optimizer):
    for param_group in optimizer.param_groups:
        return param_group['lr']
