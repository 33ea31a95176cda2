from . import utils, verification  # noqa: F401
