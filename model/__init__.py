"""Plugin namespace looked up by train.py (`importlib.import_module(f"model.{opt.model}")`, reference train.py:23)."""
from . import planar  # noqa: F401
