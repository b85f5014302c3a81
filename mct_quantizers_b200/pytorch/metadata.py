"""Attach / read a metadata dictionary on a torch model or an ONNX ModelProto
(reference: mct_quantizers/pytorch/metadata.py:25-126).  Bookkeeping only -- no arithmetic."""
from typing import Dict

import torch

from mct_quantizers_b200.common.constants import FOUND_ONNX, FRAMEWORK_VERSION, ONNX_VERSION
from mct_quantizers_b200.common.metadata import verify_and_init_metadata
from mct_quantizers_b200.logger import Logger


def add_metadata(model: torch.nn.Module, metadata: Dict) -> torch.nn.Module:
    """model.metadata = the verified dictionary (+ framework version); returns the model."""
    metadata = verify_and_init_metadata(metadata)
    metadata.setdefault(FRAMEWORK_VERSION, torch.__version__)
    model.metadata = metadata
    return model


def get_metadata(model: torch.nn.Module) -> Dict:
    return getattr(model, 'metadata', {})


if FOUND_ONNX:
    import onnx

    def add_onnx_metadata(model: "onnx.ModelProto", metadata: Dict):
        """Appends every entry to model.metadata_props (values must be str / bytes)."""
        metadata = verify_and_init_metadata(metadata)
        metadata.setdefault(ONNX_VERSION, onnx.__version__)
        for key, value in metadata.items():
            if not isinstance(value, (bytes, str)):
                Logger.critical(f"ONNX metadata must be of byte type, but {value} has type {type(value)}")
            prop = model.metadata_props.add()
            prop.key, prop.value = key, value
        return model

    def get_onnx_metadata(model: "onnx.ModelProto") -> Dict:
        return {prop.key: prop.value for prop in model.metadata_props}
else:
    def add_onnx_metadata(model, metadata):
        Logger.critical('Installing onnx is mandatory when using add_onnx_metadata. Could not find onnx package.')

    def get_onnx_metadata(model):
        Logger.critical('Installing onnx is mandatory when using get_onnx_metadata. Could not find onnx package.')
