"""ilqg_gen -- build-time generator: problem description -> reference-ABI C + sm_100a device code."""
from .problem import Problem  # noqa: F401
from .lower import lower  # noqa: F401
