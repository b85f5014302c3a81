"""torch.load wrapper (reference: mct_quantizers/pytorch/load_model.py:23-34).  Importing this package first is
what makes pickled modules loadable: it registers the `mctq` operators their graphs refer to."""
import torch

import mct_quantizers_b200.ops  # noqa: F401


def pytorch_load_quantized_model(filepath: str, **kwargs):
    """Loads a whole pickled module (wrappers / holders are plain Python objects, not state dicts).  torch >= 2.6
    defaults `torch.load` to weights_only=True, which cannot rebuild such modules (the reference's own load tests fail
    there for that reason); unless the caller says otherwise the full unpickler is used, as on the torch versions the
    reference was written for."""
    kwargs.setdefault('weights_only', False)
    return torch.load(filepath, **kwargs)
