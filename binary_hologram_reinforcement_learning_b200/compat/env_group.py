"""Drop-in for env_group.py: rank-table reward from 10 000 candidate flips scored at reset."""
from binary_hologram_reinforcement_learning_b200.envs import BinaryHologramEnvGroup as BinaryHologramEnv, RW  # noqa: F401

IPS = 256
CH = 8
