"""Import alias: ``import b200cam`` -> the package in ``privacy-preserving-vision_b200/``.

The product directory carries the reference repository's name, which is not a valid
Python identifier; this stub loads it under the importable name ``b200cam`` and
replaces itself in ``sys.modules`` (so ``import b200cam.optics`` etc. resolve into the
real directory).
"""
import importlib.util as _ilu
import pathlib as _pl
import sys as _sys

_root = _pl.Path(__file__).resolve().parent.parent / "privacy-preserving-vision_b200"
_spec = _ilu.spec_from_file_location("b200cam", _root / "__init__.py",
                                     submodule_search_locations=[str(_root)])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["b200cam"] = _mod
_spec.loader.exec_module(_mod)
