"""Import shim: the package directory is `deepspeech.pytorch_b200/` (not a legal module name), so this
one-file module loads it and registers it as `deepspeech_pytorch_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "deepspeech.pytorch_b200")
_spec = importlib.util.spec_from_file_location("deepspeech_pytorch_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["deepspeech_pytorch_b200"] = _mod
_spec.loader.exec_module(_mod)
