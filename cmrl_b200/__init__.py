"""Import alias: the package lives in `causal-mbrl_b200/` (not a valid Python identifier), so this stub points
`cmrl_b200.__path__` at that directory and runs its `__init__`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "causal-mbrl_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
