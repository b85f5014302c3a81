"""Metadata bookkeeping shared by the framework front ends (reference: mct_quantizers/common/metadata.py:23-72).
Not on the hot path; kept so that `add_metadata` / `get_metadata` users and pickled models keep working."""
import sys
from typing import Any, Dict

from mct_quantizers_b200.common.constants import MCTQ_VERSION, PYTHON_VERSION
from mct_quantizers_b200.logger import Logger


def _storable(value: Any) -> bool:
    """int / float / str, lists of storable values, dicts with str keys and storable values."""
    if isinstance(value, (int, float, str)):
        return True
    if isinstance(value, list):
        return all(_storable(v) for v in value)
    if isinstance(value, dict):
        return all(isinstance(k, str) and _storable(v) for k, v in value.items())
    return False


def verify_and_init_metadata(metadata: Dict = None) -> Dict:
    """Checks the dictionary (str keys are mandatory: Logger.error raises; odd value types only warn) and adds the
    python / package version entries when absent."""
    from mct_quantizers_b200 import __version__
    if not isinstance(metadata, dict):
        Logger.error(f'metadata should be a dictionary, but got type {type(metadata)}.')
    if any(not isinstance(k, str) for k in metadata):
        Logger.error('metadata dictionary should only have string keys.')
    if any(not _storable(v) for v in metadata.values()):
        Logger.warning('metadata dictionary values should be strings, integers, floats, lists, '
                       'or dictionaries with appropriate inner values. Other types may cause issues '
                       'with saving/loading the metadata.')
    metadata.setdefault(PYTHON_VERSION, sys.version)
    metadata.setdefault(MCTQ_VERSION, __version__)
    return metadata
