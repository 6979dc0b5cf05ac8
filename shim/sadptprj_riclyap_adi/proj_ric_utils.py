from optconpy_b200.proj_ric_utils import *  # noqa: F401,F403
from optconpy_b200.proj_ric_utils import __all__  # noqa: F401
