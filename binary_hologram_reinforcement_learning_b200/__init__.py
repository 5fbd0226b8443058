"""B200-native reward / direct-binary-search engine for binary holograms.

Hot path of songyb111-gachon/binary-hologram-reinforcement-learning (env.py,
env_1024_24*.py, env_group.py, DBS*.py) rebuilt as hand-written sm_100a CUDA
kernels behind a C ABI (include/bholo.h), with the reference's gymnasium env and
DBS entry points on top.  There is no CPU fallback.
"""
from .engine import (HoloEngine, HoloError, RULE_ENV, RULE_DBS, RULE_NEVER, RESULT_DTYPE,
                     load_library, simulate)
from .envs import (BinaryHologramEnv, BinaryHologramEnvRGB, BinaryHologramEnvRGBCrop,
                   BinaryHologramEnvGroup, RW, WL_MONO, WL_RGB)
from .vec_env import HologramVecEnv, as_sb3_vec_env
from .dbs import (optimize_with_random_pixel_flips, dbs_greedy_env, dbs_sweep, sweep_engine,
                  decile_index, OUTPUT_BINS)
from .synthetic import synthetic_problem, SyntheticLoader
from .data import ImageFolderLoader, load_image, crop_to

__all__ = [
    "HoloEngine", "HoloError", "RULE_ENV", "RULE_DBS", "RULE_NEVER", "RESULT_DTYPE", "load_library",
    "simulate", "BinaryHologramEnv", "BinaryHologramEnvRGB", "BinaryHologramEnvRGBCrop",
    "BinaryHologramEnvGroup", "RW", "WL_MONO", "WL_RGB", "HologramVecEnv", "as_sb3_vec_env",
    "optimize_with_random_pixel_flips", "dbs_greedy_env", "dbs_sweep", "sweep_engine",
    "decile_index", "OUTPUT_BINS", "synthetic_problem", "SyntheticLoader", "ImageFolderLoader",
    "load_image", "crop_to",
]
