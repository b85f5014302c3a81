"""torch.load wrapper (reference: mct_quantizers/pytorch/load_model.py:23-34).  Importing this package first is
what makes pickled modules loadable: it registers the `mctq` operators their graphs refer to."""
import torch

import mct_quantizers_b200.ops  # noqa: F401


def pytorch_load_quantized_model(filepath: str, **kwargs):
    return torch.load(filepath, **kwargs)
