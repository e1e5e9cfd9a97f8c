"""Classic-control task families of the hot path: cart-pole (discrete and continuous, balancing and swing-up) and
charged-ball centering."""
from .charged_ball import ChargedBallCenteringEnv, ContinuousChargedBallCenteringEnv
from .cartpole import BaseCartPoleEnv, CartPoleBalancingEnv, CartPoleSwingUpEnv
from .cartpole import ContinuousCartPoleBalancingEnv, ContinuousCartPoleSwingUpEnv

__all__ = [
    "BaseCartPoleEnv", "CartPoleBalancingEnv", "CartPoleSwingUpEnv", "ContinuousCartPoleBalancingEnv",
    "ContinuousCartPoleSwingUpEnv", "ChargedBallCenteringEnv", "ContinuousChargedBallCenteringEnv",
]
