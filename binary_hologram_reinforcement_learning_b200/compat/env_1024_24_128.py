"""Drop-in for env_1024_24_128.py: RGB env, centre 896^2 window simulated (crop_margin=64)."""
from binary_hologram_reinforcement_learning_b200.envs import BinaryHologramEnvRGBCrop as BinaryHologramEnv, RW  # noqa: F401

IPS = 1024
CH = 24
