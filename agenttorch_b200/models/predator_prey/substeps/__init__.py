from .move import *  # noqa: F401,F403
from .eat import *  # noqa: F401,F403
from .hunt import *  # noqa: F401,F403
from .grow import *  # noqa: F401,F403
