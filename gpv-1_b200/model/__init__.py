from .criterion import GPVCriterion, SetCriterion  # noqa: F401
from .gpv import GPV, HostTargets  # noqa: F401
from .matcher import HungarianMatcher  # noqa: F401
from .spec import gpv_specs  # noqa: F401
