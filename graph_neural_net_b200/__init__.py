"""graph_neural_net_b200: B200-native implementation of the 2-FGNN siamese hot path of
mlelarge/graph_neural_net behind the reference's own Python API (models, maskedtensors,
toolbox.losses/metrics, loaders) -- numerics run in libfgnn_b200.so (CUDA, sm_100a) via a C ABI.
"""
import sys as _sys

from . import _lib
from ._lib import FgnnError, get_lib  # noqa: F401
from . import maskedtensors, toolbox, loaders, models  # noqa: F401

__all__ = ["models", "maskedtensors", "toolbox", "loaders", "FgnnError", "get_lib", "install_as_reference"]


def install_as_reference():
    """Alias the sub-packages under the reference's top-level import names (`models`, `toolbox`,
    `maskedtensors`, `loaders`) so unmodified reference-side scripts pick up this implementation."""
    for name in ("models", "toolbox", "maskedtensors", "loaders"):
        pkg = _sys.modules[__name__ + "." + name]
        _sys.modules[name] = pkg
        for sub, mod in list(_sys.modules.items()):
            if sub.startswith(__name__ + "." + name + "."):
                _sys.modules[sub[len(__name__) + 1:]] = mod
    _sys.modules.setdefault("maskedtensor", _sys.modules[__name__ + ".maskedtensors.maskedtensor"])
