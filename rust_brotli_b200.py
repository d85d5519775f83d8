"""Import shim: the package directory is named ``rust-brotli_b200`` (not a valid Python identifier), so
``import rust_brotli_b200`` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rust-brotli_b200")
_spec = importlib.util.spec_from_file_location("rust_brotli_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rust_brotli_b200"] = _mod
_spec.loader.exec_module(_mod)
