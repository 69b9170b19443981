"""Import shim: the package directory is named `da-clip_b200/` (not a valid Python identifier), so this
module makes it importable as `daclip_b200` by pointing its package search path at that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "da-clip_b200")]
exec(compile(open(_os.path.join(__path__[0], "__init__.py")).read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
