"""Drop-in for env_05.py: the mono env with T_PSNR_DIFF = 0.5 by default (env_05.py:38)."""
from binary_hologram_reinforcement_learning_b200.envs import BinaryHologramEnv as _Env, RW  # noqa: F401

IPS = 256
CH = 8


class BinaryHologramEnv(_Env):
    def __init__(self, target_function, trainloader, max_steps=10000, T_PSNR=30, T_steps=1,
                 T_PSNR_DIFF=0.5, **kw):
        super().__init__(target_function, trainloader, max_steps, T_PSNR, T_steps, T_PSNR_DIFF, **kw)
