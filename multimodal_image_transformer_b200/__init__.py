from . import _lib  # noqa: F401
