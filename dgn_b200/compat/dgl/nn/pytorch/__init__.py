from . import glob  # noqa: F401
