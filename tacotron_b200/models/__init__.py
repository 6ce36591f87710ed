from . import ops, tacotron  # noqa: F401
