from .gpv import GPV, HostTargets  # noqa: F401
from .spec import gpv_specs  # noqa: F401
