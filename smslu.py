"""Import shim: the package directory is named `sharedmemsparselu.jl_b200` (with a dot, after the
reference's repository name), which Python's import statement cannot spell.  `import smslu`
loads it under the module name `sharedmemsparselu_jl_b200` and re-exports its public names."""
import importlib.util as _u
import os as _os
import sys as _sys

_NAME = "sharedmemsparselu_jl_b200"
_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "sharedmemsparselu.jl_b200")
if _NAME not in _sys.modules:
    _spec = _u.spec_from_file_location(_NAME, _os.path.join(_DIR, "__init__.py"),
                                       submodule_search_locations=[_DIR])
    _mod = _u.module_from_spec(_spec)
    _sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
pkg = _sys.modules[_NAME]
globals().update({k: v for k, v in vars(pkg).items() if not k.startswith("_")})
