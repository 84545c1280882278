from scipy import *  # noqa: F401,F403
from scipy import linalg  # noqa: F401
