from .cartpole import (
    BaseCartPoleEnv,
    CartPoleBalancingEnv,
    CartPoleSwingUpEnv,
    ContinuousCartPoleBalancingEnv,
    ContinuousCartPoleSwingUpEnv,
)
from .charged_ball import ChargedBallCenteringEnv, ContinuousChargedBallCenteringEnv
