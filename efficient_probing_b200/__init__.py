"""Import alias: the package lives in ``efficient-probing_b200/`` (the layout name the build contract
fixes); a hyphen is not importable, so this stub re-points ``efficient_probing_b200`` at that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "efficient-probing_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
