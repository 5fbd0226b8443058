"""Drop-in for env_1024_24.py (:29-31): 1024^2 x 24 RGB env."""
from binary_hologram_reinforcement_learning_b200.envs import BinaryHologramEnvRGB as BinaryHologramEnv, RW  # noqa: F401

IPS = 1024
CH = 24
