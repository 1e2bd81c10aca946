"""Import shim that makes the UNMODIFIED reference importable in this container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py (run here, where
/root/reference exists) to generate the committed fixtures under tests/golden/.
Nothing on the GPU box imports this file: /root/reference does not exist there.

Two stand-ins are installed in sys.modules before the reference is imported:
  * pytorch_lightning  -- not installed; reference models/trainers.py:1 imports it.
    LightningModule := nn.Module with a no-op log().
  * numpy.lib.arraysetops -- removed in numpy 2.x; reference toolbox/utils.py:6 and
    toolbox/metrics.py:2 import `isin` from it (unused).
The reference source itself is untouched.
"""
import os
import sys
import types

import numpy as np
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("FGNN_REFERENCE_ROOT", "/root/reference")


def install(reference_root: str = REFERENCE_ROOT) -> str:
    if not os.path.isdir(reference_root):
        raise FileNotFoundError(f"reference tree not found at {reference_root}")
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def log(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        pl.seed_everything = lambda *a, **k: None
        sys.modules["pytorch_lightning"] = pl
    if "numpy.lib.arraysetops" not in sys.modules:
        ar = types.ModuleType("numpy.lib.arraysetops")
        ar.isin = np.isin
        sys.modules["numpy.lib.arraysetops"] = ar
    for p in (reference_root,):
        if p not in sys.path:
            sys.path.insert(0, p)
    return reference_root
