from . import *  # noqa: F401,F403
