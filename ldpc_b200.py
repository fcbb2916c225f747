"""Import shim: the package directory is named `ldpc-neuralnetwork-decoder_b200` (not a valid
Python identifier), so this module loads it under the importable name `ldpc_b200`.  After
`import ldpc_b200`, `ldpc_b200.models`, `ldpc_b200.utils`, ... are the package's modules."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ldpc-neuralnetwork-decoder_b200")
_spec = importlib.util.spec_from_file_location(
    "ldpc_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ldpc_b200"] = _mod
_spec.loader.exec_module(_mod)
