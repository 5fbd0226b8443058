"""Drop-in for env_md.py: the mono env with a MultiDiscrete([CH, IPS, IPS]) action (env_md.py:54,160)."""
import functools

from binary_hologram_reinforcement_learning_b200.envs import BinaryHologramEnv as _Env, RW  # noqa: F401

IPS = 256
CH = 8


class BinaryHologramEnv(_Env):
    def __init__(self, *args, **kw):
        kw.setdefault("action_mode", "multidiscrete")
        super().__init__(*args, **kw)
