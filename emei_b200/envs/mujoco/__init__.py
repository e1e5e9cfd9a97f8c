from .inverted_pendulum import (
    ReboundInvertedPendulumSwingUpEnv,
    ReboundInvertedPendulumBalancingEnv,
    BoundaryInvertedPendulumSwingUpEnv,
    BoundaryInvertedPendulumBalancingEnv,
)
from .inverted_double_pendulum import (
    ReboundInvertedDoublePendulumSwingUpEnv,
    ReboundInvertedDoublePendulumBalancingEnv,
    BoundaryInvertedDoublePendulumSwingUpEnv,
    BoundaryInvertedDoublePendulumBalancingEnv,
)
from .hopper import HopperRunningEnv
from .half_cheetah import HalfCheetahRunningEnv
