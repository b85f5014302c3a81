"""Drop-in import alias: make `import mct_quantizers` resolve to this package.

    import mct_quantizers_b200.compat          # or: mct_quantizers_b200.compat.install()
    import mct_quantizers                      # -> mct_quantizers_b200
    from mct_quantizers.pytorch.quantizers.weights_inferable_quantizers.weights_symmetric_inferable_quantizer import \\
        WeightsSymmetricInferableQuantizer     # the B200 class

The package mirrors the reference's module tree for the PyTorch path (mct_quantizers/__init__.py:16-34 and the module
paths its users and its own tests import), so every `mct_quantizers.<sub.module>` name is served by the module
`mct_quantizers_b200.<sub.module>` -- the SAME module object under two names, so classes, the `@mark_quantizer`
registry and `isinstance` checks stay consistent.  Consequences:

  * code written against the reference runs unmodified (its own unit tests do: see DESIGN.md);
  * models pickled by the reference (`torch.save(module)`; class paths `mct_quantizers.…`) load as B200 objects through
    `pytorch_load_quantized_model` -- the attribute layout of the quantizer / wrapper / holder classes is the same.

Sub-trees that are out of scope here (mct_quantizers.keras, onnxruntime helpers) are not aliased: importing them raises
ModuleNotFoundError.  install() refuses to run when the real reference package has already been imported.
"""
import importlib
import importlib.abc
import importlib.util
import sys

_SRC = "mct_quantizers"
_DST = "mct_quantizers_b200"


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != _SRC and not fullname.startswith(_SRC + "."):
            return None
        real = _DST + fullname[len(_SRC):]
        try:
            real_spec = importlib.util.find_spec(real)
        except (ImportError, ValueError):
            return None
        if real_spec is None:
            return None
        return importlib.util.spec_from_loader(fullname, self, is_package=real_spec.submodule_search_locations is not None)

    def create_module(self, spec):
        return importlib.import_module(_DST + spec.name[len(_SRC):])     # the real module, registered under the alias too

    def exec_module(self, module):
        pass


_finder = _AliasFinder()


def install():
    """Idempotent.  Raises ImportError if the reference package itself is already loaded in this process."""
    if _finder in sys.meta_path:
        return
    loaded = sys.modules.get(_SRC)
    if loaded is not None and not getattr(loaded, "__name__", "").startswith(_DST):
        raise ImportError("the reference package `mct_quantizers` is already imported; install the alias before it "
                          "(or do not import both in one process)")
    sys.meta_path.insert(0, _finder)


def uninstall():
    if _finder in sys.meta_path:
        sys.meta_path.remove(_finder)
    for name in [n for n in sys.modules if n == _SRC or n.startswith(_SRC + ".")]:
        mod = sys.modules[name]
        if getattr(mod, "__name__", "").startswith(_DST):
            del sys.modules[name]


install()
