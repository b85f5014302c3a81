"""Names and constants shared by the quantizers, wrapper and holders.

They are part of the public surface (MCT reads e.g. ``constants.LAYER`` and ``constants.EPS``), so every name and value
equals the reference's (mct_quantizers/common/constants.py:27-97).  Most of them are keyword / attribute names whose value
is simply the constant's own name in lower case; those are generated from one list, the irregular ones are spelled out."""
import importlib.util

# ---- constants whose value is their own name in lower case
_SELF_NAMED = (
    # frameworks probed below
    "TENSORFLOW", "TORCH", "ONNX", "ONNXRUNTIME", "ONNXRUNTIME_EXTENSIONS",
    # keys of a wrapper / holder configuration
    "IS_WEIGHTS", "IS_ACTIVATIONS", "OP_CALL_ARGS", "OP_CALL_KWARGS",
    "WEIGHTS_QUANTIZATION_METHOD", "WEIGHTS_N_BITS", "WEIGHTS_QUANTIZATION_PARAMS", "ENABLE_WEIGHTS_QUANTIZATION",
    "WEIGHTS_CHANNELS_AXIS", "WEIGHTS_PER_CHANNEL_THRESHOLD", "MIN_THRESHOLD",
    "ACTIVATION_QUANTIZATION_METHOD", "ACTIVATION_N_BITS", "ACTIVATION_QUANTIZATION_PARAMS",
    "ENABLE_ACTIVATION_QUANTIZATION",
    # class attributes written by @mark_quantizer, attribute names of the wrapper (many quantizers) and the holder (one)
    "QUANTIZATION_TARGET", "QUANTIZATION_METHOD", "ACTIVATION_QUANTIZERS", "ACTIVATION_HOLDER_QUANTIZER",
    # parameter names of the quantizer constructors
    "NUM_BITS", "SIGNED", "THRESHOLD", "PER_CHANNEL", "MIN_RANGE", "MAX_RANGE", "CHANNEL_AXIS", "INPUT_RANK", "LUT_VALUES",
    # misc
    "LAYER", "TRAINING", "POSITIONAL_WEIGHT",
    # metadata fields
    "FRAMEWORK_VERSION", "PYTHON_VERSION", "MCTQ_VERSION", "ONNX_VERSION",
)
for _name in _SELF_NAMED:
    globals()[_name] = _name.lower()
del _name

# ---- the irregular ones
WEIGHTS_QUANTIZERS = "weights_quantizer"
WEIGHTS_VALUES = "weights_value"
IS_INPUT_AS_LIST = "is_inputs_as_list"
QUANTIZER_ID = "identifier"
STEPS = "optimizer_step"
QUANTIZED_POSITIONAL_WEIGHT = "quantized_" + POSITIONAL_WEIGHT            # noqa: F821  (generated above)
ONNX_CUSTOM_OP_DOMAIN = "mct_quantizers"
EPS = 1e-8                       # added to a LUT threshold before the division
LUT_VALUES_BITWIDTH = 8          # default grid of the LUT centroids: integers of this many bits


# ---- optional frameworks
def _found(pkg):
    return importlib.util.find_spec(pkg) is not None


FOUND_TF = False                 # the Keras twin is out of scope for the B200 path
FOUND_TORCH = _found(TORCH)                                               # noqa: F821
FOUND_ONNX = _found(ONNX)                                                 # noqa: F821
FOUND_ONNXRUNTIME = _found(ONNXRUNTIME)                                   # noqa: F821
FOUND_ONNXRUNTIME_EXTENSIONS = _found(ONNXRUNTIME_EXTENSIONS)             # noqa: F821
