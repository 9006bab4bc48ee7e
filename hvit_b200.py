"""Import alias: `import hvit_b200` loads the package that lives in
`speech-enhancement-via-hybrid-vision-transformer-project_b200/` (a directory name
Python cannot import directly because of the hyphens)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                     "speech-enhancement-via-hybrid-vision-transformer-project_b200")
_spec = _ilu.spec_from_file_location("hvit_b200", _os.path.join(_DIR, "__init__.py"),
                                     submodule_search_locations=[_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["hvit_b200"] = _mod
_spec.loader.exec_module(_mod)
