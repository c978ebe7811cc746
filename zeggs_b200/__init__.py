"""Import shim: the package sources live in `ubisoft-laforge-zeroeggs_b200/` (a directory name
that is not a valid Python identifier); `import zeggs_b200` resolves to them."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "ubisoft-laforge-zeroeggs_b200")
__path__ = [_real]
_init = _os.path.join(_real, "__init__.py")
exec(compile(open(_init).read(), _init, "exec"))
