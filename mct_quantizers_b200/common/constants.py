"""Names and constants shared by the quantizers, wrapper and holders.
Same names and values as the reference's mct_quantizers/common/constants.py:27-97 (they are part of the
public surface: MCT reads e.g. ``constants.LAYER`` and ``constants.EPS``)."""
import importlib.util


def _found(pkg):
    return importlib.util.find_spec(pkg) is not None


TENSORFLOW, TORCH, ONNX = 'tensorflow', 'torch', 'onnx'
ONNXRUNTIME, ONNXRUNTIME_EXTENSIONS = 'onnxruntime', 'onnxruntime_extensions'
FOUND_TF = False                      # the Keras twin is out of scope for the B200 path
FOUND_TORCH = _found(TORCH)
FOUND_ONNX = _found(ONNX)
FOUND_ONNXRUNTIME = _found(ONNXRUNTIME)
FOUND_ONNXRUNTIME_EXTENSIONS = _found(ONNXRUNTIME_EXTENSIONS)

# quantization properties
IS_WEIGHTS = "is_weights"
IS_ACTIVATIONS = "is_activations"
WEIGHTS_QUANTIZERS = "weights_quantizer"
WEIGHTS_VALUES = "weights_value"
OP_CALL_ARGS = 'op_call_args'
OP_CALL_KWARGS = 'op_call_kwargs'
IS_INPUT_AS_LIST = 'is_inputs_as_list'
WEIGHTS_QUANTIZATION_METHOD = 'weights_quantization_method'
WEIGHTS_N_BITS = 'weights_n_bits'
WEIGHTS_QUANTIZATION_PARAMS = 'weights_quantization_params'
ENABLE_WEIGHTS_QUANTIZATION = 'enable_weights_quantization'
WEIGHTS_CHANNELS_AXIS = 'weights_channels_axis'
WEIGHTS_PER_CHANNEL_THRESHOLD = 'weights_per_channel_threshold'
MIN_THRESHOLD = 'min_threshold'
ACTIVATION_QUANTIZATION_METHOD = 'activation_quantization_method'
ACTIVATION_N_BITS = 'activation_n_bits'
ACTIVATION_QUANTIZATION_PARAMS = 'activation_quantization_params'
ENABLE_ACTIVATION_QUANTIZATION = 'enable_activation_quantization'

# class attributes set by @mark_quantizer
QUANTIZATION_TARGET = 'quantization_target'
QUANTIZATION_METHOD = 'quantization_method'
QUANTIZER_ID = 'identifier'

ACTIVATION_QUANTIZERS = "activation_quantizers"
ACTIVATION_HOLDER_QUANTIZER = "activation_holder_quantizer"

# quantizer signature parameter names
NUM_BITS = 'num_bits'
SIGNED = 'signed'
THRESHOLD = 'threshold'
PER_CHANNEL = 'per_channel'
MIN_RANGE = 'min_range'
MAX_RANGE = 'max_range'
CHANNEL_AXIS = 'channel_axis'
INPUT_RANK = 'input_rank'
LUT_VALUES = 'lut_values'

# values
LAYER = "layer"
STEPS = "optimizer_step"
TRAINING = "training"
EPS = 1e-8
LUT_VALUES_BITWIDTH = 8

POSITIONAL_WEIGHT = 'positional_weight'
QUANTIZED_POSITIONAL_WEIGHT = f'quantized_{POSITIONAL_WEIGHT}'

ONNX_CUSTOM_OP_DOMAIN = "mct_quantizers"

FRAMEWORK_VERSION = 'framework_version'
PYTHON_VERSION = 'python_version'
MCTQ_VERSION = "mctq_version"
ONNX_VERSION = 'onnx_version'
