from jacks_b200.preprocess import *  # noqa: F401,F403
from jacks_b200.preprocess import LOG  # noqa: F401
