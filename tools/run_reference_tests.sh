# The reference's OWN PyTorch unit tests (unmodified) against this package through the import alias
# (mct_quantizers_b200/compat.py).  Build container: copy them to the git-ignored scratch directory first
#     mkdir -p baseline/_ref_tests && cp -r /root/reference/tests/pytorch_tests baseline/_ref_tests/
#     rm -rf baseline/_ref_tests/pytorch_tests/onnx_export_tests        (needs onnx / onnxruntime)
#     printf 'import os, sys\nsys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))\nimport mct_quantizers_b200.compat\n' > baseline/_ref_tests/conftest.py
# (test_pytorch_load_model.py imports onnx at module level and is skipped)
# then on the GPU box:   bash tools/run_reference_tests.sh > gpurun_out/reference_test_suite.txt
cd baseline/_ref_tests && python -m pytest pytorch_tests -q -p no:cacheprovider --ignore=pytorch_tests/test_pytorch_load_model.py 2>&1 | grep -v Warning | tail -40
