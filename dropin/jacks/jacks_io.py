from jacks_b200.jacks_io import *  # noqa: F401,F403
from jacks_b200.jacks_io import LOG, inferJACKS  # noqa: F401
