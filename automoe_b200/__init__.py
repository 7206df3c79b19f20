"""Importable alias of the `self-driving-model_b200/` package directory (a hyphen cannot
appear in a Python import statement).  All sub-modules live there."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "self-driving-model_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
