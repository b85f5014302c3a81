"""Base class of the PyTorch inferable quantizers.
Reference: mct_quantizers/pytorch/quantizers/base_pytorch_inferable_quantizer.py:24-62 (state: custom-impl flag
for ONNX tracing and the output-reuse cache, including the attribute spelt `resue_outputs`).

Added here: a per-device cache of the quantizer's parameter tensors.  The reference pins parameters to the
device that was current at construction, so `module.to('cuda:3')` leaves them behind; this class hands the
kernels a copy on the *input's* device, made once."""
from abc import abstractmethod

import torch

from mct_quantizers_b200.common.base_inferable_quantizer import BaseInferableQuantizer


class BasePyTorchInferableQuantizer(BaseInferableQuantizer):
    def __init__(self):
        super(BasePyTorchInferableQuantizer, self).__init__()
        self._use_custom_impl = False       # ONNX export path (only honoured under torch.jit tracing)
        self.reuse = False
        self.enable_reuse = False           # return the first call's output on every later call
        self.quantizer_first_run = True
        self.resue_outputs = None
        self._per_device = {}

    def enable_custom_impl(self):
        self._use_custom_impl = True

    def enable_reuse_quantizer(self):
        self.enable_reuse = True
        self.quantizer_first_run = True

    def disable_reuse_quantizer(self):
        self.enable_reuse = False

    # ---- per-device parameter residency
    def _on(self, device, *tensors):
        """The given parameter tensors on `device` (identity when already there; copies are made once)."""
        out = []
        for t in tensors:
            if t.device == device:
                out.append(t)
                continue
            key = (id(t), str(device))
            ver = -1 if t.is_inference() else t._version      # inference tensors carry no version counter
            hit = self._per_device.get(key)
            if hit is None or hit[0] is not t or hit[2] != ver:      # a new tensor object, or edited in place since the copy
                hit = (t, t.to(device), ver)
                self._per_device[key] = hit
            out.append(hit[1])
        return out

    def _validated_launch_args(self, scale, zero_point):
        """(key, (scale narrowed to f32, zp, qmin, qmax)) for the per-tensor scalar kernels: validated once (the
        checks ATen makes on every call) and cached until an attribute changes."""
        import numpy as np
        qmin, qmax = int(self.min_quantized_domain), int(self.max_quantized_domain)
        if qmin > qmax:
            raise RuntimeError("`quant_min` should be less than or equal to `quant_max`.")
        if not qmin <= int(zero_point) <= qmax:
            raise RuntimeError("`zero_point` must be between `quant_min` and `quant_max`.")
        cached = ((scale, zero_point, self.min_quantized_domain, self.max_quantized_domain),
                  (float(np.float32(scale)), int(zero_point), qmin, qmax))
        self.__dict__['_launch_args'] = cached
        return cached

    def __getstate__(self):
        state = dict(self.__dict__)
        state['_per_device'] = {}           # device copies are derived data; do not pickle them
        state.pop('_direct_cache', None)    # launch constants (device pointers) of the lean LUT path
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.__dict__.setdefault('_per_device', {})

    @abstractmethod
    def __call__(self, inputs: torch.Tensor):
        raise NotImplemented(f'{self.__class__.__name__} did not implement __call__')  # pragma: no cover
