"""Import alias: the product package lives in the directory `nans-clip_b200/` (the name the
build contract fixes); a hyphen is not importable, so `import nans_clip_b200` resolves here and
forwards to that directory.  No code of its own."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "nans-clip_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
