"""Import shim: the package lives in the directory `moss-ttsd_b200/` (not a legal Python identifier);
`import moss_ttsd_b200` resolves its submodules from there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "moss-ttsd_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
