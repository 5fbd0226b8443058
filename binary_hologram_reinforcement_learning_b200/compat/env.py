"""Drop-in for the reference module of the same name (env.py:27-29,37): 256^2 x 8 mono env."""
from binary_hologram_reinforcement_learning_b200.envs import BinaryHologramEnv, RW  # noqa: F401

IPS = 256
CH = 8
