"""Mirror of interact_drive/car/linear_reward_car.py:12-55 of the reference."""
import numpy as np

from ...runtime import as_f32
from .car import Car


class LinearRewardCar(Car):
    """Reward = weights . features(state, control); the weights are kept L2-normalised in float32."""

    def __init__(self, env, init_state, weights, color: str = "gray", opacity: float = 1.0, friction: float = 0.2,
                 **kwargs):
        super().__init__(env, init_state, color=color, opacity=opacity, friction=friction, **kwargs)
        self.weights = weights

    def features(self, state, control):
        raise NotImplementedError

    @property
    def weights(self):
        return self.weights_f32.copy()

    @weights.setter
    def weights(self, weights):
        w = np.asarray(weights)
        self.weights_f32 = as_f32(w / np.linalg.norm(w))      # normalise, then cast (reference :34,:47)

    def reward_fn(self, state, control, weights=None):
        feats = self.features(state, control)
        w = self.weights_f32 if weights is None else as_f32(weights)
        return np.sum(w * feats, axis=-1, dtype=np.float32)
