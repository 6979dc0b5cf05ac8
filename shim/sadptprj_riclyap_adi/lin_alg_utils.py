from optconpy_b200.lin_alg_utils import *  # noqa: F401,F403
from optconpy_b200.lin_alg_utils import __all__  # noqa: F401
