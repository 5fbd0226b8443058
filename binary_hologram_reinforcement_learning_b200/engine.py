"""ctypes binding of the C ABI in ``include/bholo.h`` (libbholo_b200.so).

There is no CPU fallback: if the CUDA library is missing, cannot be loaded, or
no sm_100 device is present, constructing a :class:`HoloEngine` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import _build

RULE_ENV, RULE_DBS, RULE_NEVER, RULE_ALWAYS = 0, 1, 2, 3
OBS_DEVICE, OBS_PINNED_HOST, OBS_CONTEXT = 0, 1, 2
OBS_COMMITTED_ONLY, OBS_FULL, OBS_SYNC = 1, 2, 4
METHOD_ASM, METHOD_FRESNEL = 0, 1

# every symbol include/bholo.h declares
ABI_SYMBOLS = (
    "bh_abi_version", "bh_last_error", "bh_create", "bh_destroy", "bh_set_stream",
    "bh_set_target", "bh_load_state", "bh_clone_env", "bh_resync", "bh_get_metrics", "bh_eval_flips",
    "bh_step_batch", "bh_vec_step", "bh_vec_book_update", "bh_step_batch_device", "bh_rollout_device", "bh_rollout_status", "bh_eval_flips_device", "bh_max_tasks",
    "bh_commit_flip", "bh_dbs_run", "bh_dbs_run_batch", "bh_sweep_all", "bh_sweep_stats", "bh_get_recon", "bh_get_state", "bh_get_field",
    "bh_recon_batch", "bh_recon_device_block", "bh_recon_planes_written", "bh_stream_sync",
    "bh_device_ptr", "bh_host_alloc", "bh_host_free", "bh_simulate", "bh_time_eval", "bh_time_step", "bh_time_commit",
    "bh_time_propagate", "bh_time_propagate_passes", "bh_launch_count",
)


class BhResult(C.Structure):
    """Mirror of ``bh_result`` (40 bytes)."""
    _fields_ = [("psnr_after", C.c_double), ("d_sii", C.c_double), ("d_sit", C.c_double),
                ("action", C.c_int64), ("accept", C.c_int32), ("sgn", C.c_int32)]


RESULT_DTYPE = np.dtype([("psnr_after", "<f8"), ("d_sii", "<f8"), ("d_sit", "<f8"),
                         ("action", "<i8"), ("accept", "<i4"), ("sgn", "<i4")])
assert RESULT_DTYPE.itemsize == C.sizeof(BhResult) == 40


class VecBook(C.Structure):
    """Mirror of ``bh_vec_book`` (host arrays updated by ``bh_vec_step``)."""
    _fields_ = [("state", C.c_void_p), ("state_record", C.c_void_p), ("stride", C.c_int64),
                ("prev_psnr", C.c_void_p), ("init_psnr", C.c_void_p), ("steps", C.c_void_p),
                ("flips", C.c_void_p), ("t_psnr_diff", C.c_void_p), ("t_psnr", C.c_void_p),
                ("max_steps", C.c_void_p), ("reward_scale", C.c_double), ("rewards", C.c_void_p),
                ("psnr_change", C.c_void_p), ("psnr_diff", C.c_void_p),
                ("last_candidate", C.c_void_p), ("event", C.c_void_p)]


class HoloError(RuntimeError):
    pass


_LIB = None


def library_path() -> str:
    return os.environ.get("BHOLO_LIB", _build.LIB_PATH)


def load_library(build_if_missing: bool = True):
    """dlopen libbholo_b200.so and declare the signatures.  Raises if unavailable."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise HoloError(f"{path} is missing; run python -m binary_hologram_reinforcement_learning_b200._build")
        _build.build()
    lib = C.CDLL(path)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    P = C.POINTER
    sig = {
        "bh_abi_version": (i32, []),
        "bh_last_error": (C.c_char_p, [vp]),
        "bh_create": (i32, [P(vp), i32, i32, i32, i32, i32, P(dbl), dbl, dbl, i32, i32, i32]),
        "bh_destroy": (i32, [vp]),
        "bh_set_stream": (i32, [vp, vp]),
        "bh_set_target": (i32, [vp, i32, vp, i32]),
        "bh_load_state": (i32, [vp, i32, vp, i32]),
        "bh_resync": (i32, [vp, i32]),
        "bh_clone_env": (i32, [vp, i32, i32]),
        "bh_get_metrics": (i32, [vp, i32, P(dbl), P(dbl), P(dbl)]),
        "bh_eval_flips": (i32, [vp, i32, i64, vp, vp, vp]),
        "bh_step_batch": (i32, [vp, i32, vp, vp, i32, vp]),
        "bh_vec_step": (i32, [vp, i32, vp, vp, i32, vp, vp]),
        "bh_vec_book_update": (i32, [i32, vp, vp, vp, vp]),
        "bh_step_batch_device": (i32, [vp, i32, vp, vp, i32, vp]),
        "bh_rollout_device": (i32, [vp, i32, vp, vp, i64, i64, i32, i32, vp, i64, i64]),
        "bh_rollout_status": (i32, [vp]),
        "bh_eval_flips_device": (i32, [vp, i32, i32, vp, vp, vp]),
        "bh_max_tasks": (i32, [vp]),
        "bh_commit_flip": (i32, [vp, i32, i64]),
        "bh_dbs_run": (i32, [vp, i32, vp, i64, i32, i64, vp, vp, P(i64), P(dbl)]),
        "bh_dbs_run_batch": (i32, [vp, i32, vp, vp, i64, i64, vp, vp, vp, vp]),
        "bh_sweep_all": (i32, [vp, i32, vp, i32]),
        "bh_sweep_stats": (i32, [vp, i32, vp, vp, vp, vp, vp, vp]),
        "bh_get_recon": (i32, [vp, i32, vp, i32, i64]),
        "bh_get_state": (i32, [vp, i32, vp, i32]),
        "bh_recon_batch": (i32, [vp, i32, vp, vp, vp, i32, i32, i32]),
        "bh_recon_device_block": (vp, [vp, i32]),
        "bh_stream_sync": (i32, [vp]),
        "bh_recon_planes_written": (i64, [vp]),
        "bh_time_step": (i32, [vp, i32, vp, vp, i32, i32, i32, i32, P(C.c_float)]),
        "bh_time_commit": (i32, [vp, i32, vp, vp, i32, i32, P(C.c_float)]),
        "bh_get_field": (i32, [vp, i32, i32, vp, i32]),
        "bh_device_ptr": (vp, [vp, i32]),
        "bh_host_alloc": (vp, [C.c_size_t]),
        "bh_host_free": (i32, [vp]),
        "bh_simulate": (i32, [i32, vp, vp, i32, i32, i32, dbl, dbl, dbl, i32, i32, vp, i32]),
        "bh_time_eval": (i32, [vp, i32, vp, vp, i32, i32, P(C.c_float)]),
        "bh_time_propagate": (i32, [vp, i32, i32, P(C.c_float)]),
        "bh_time_propagate_passes": (i32, [vp, i32, i32, P(C.c_float)]),
        "bh_launch_count": (i64, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


class _PinnedBlock:
    """Owner of one cudaHostAlloc block; freed when the last numpy view dies."""

    def __init__(self, lib, nbytes: int):
        self.lib, self.nbytes = lib, nbytes
        self.ptr = lib.bh_host_alloc(nbytes)
        if not self.ptr:
            raise HoloError("bh_host_alloc failed")

    def __del__(self):
        try:
            self.lib.bh_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array in page-locked host memory (falls back to pageable memory without a GPU)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    try:
        blk = _PinnedBlock(load_library(), max(n, 1))
    except Exception:
        return np.empty(shape, dtype=dtype)
    buf = (C.c_char * max(n, 1)).from_address(blk.ptr)
    buf._pinned_owner = blk                      # keeps the block alive with the view
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class DeviceArray:
    """Zero-copy view of device memory owned by an engine (``__cuda_array_interface__`` v3):
    ``torch.as_tensor(view, device="cuda")`` / ``cupy.asarray(view)`` alias it without a copy."""

    def __init__(self, ptr: int, shape, typestr: str, owner=None):
        self.ptr, self.shape, self.typestr, self._owner = int(ptr), tuple(int(v) for v in shape), typestr, owner

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.typestr, "data": (self.ptr, False), "version": 3,
                "strides": None}

    @property
    def dtype(self):
        return np.dtype(self.typestr)

    def __getitem__(self, i: int) -> "DeviceArray":
        """Sub-view along the first axis."""
        i = int(i)
        if not 0 <= i < self.shape[0]:
            raise IndexError(i)
        step = int(np.prod(self.shape[1:])) * self.dtype.itemsize
        return DeviceArray(self.ptr + i * step, self.shape[1:], self.typestr, self._owner)

    def numpy(self) -> np.ndarray:
        """Copy to the host (debugging / tests)."""
        import torch
        return torch.as_tensor(self, device="cuda").cpu().numpy()


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HoloEngine:
    """Device-resident state of ``n_env`` hologram environments.

    Owns U (fields), I (reconstruction), T (target), the binary state and the
    running loss sums on one GPU; all arithmetic of the reward / DBS path runs in
    the CUDA kernels behind the C ABI.
    """

    def __init__(self, N: int, F: int, wl: Sequence[float], n_env: int = 1, device: int = 0,
                 dx: float = 7.56e-6, z: float = 2e-3, pad: int = 1, relative: bool = True,
                 method: str = "asm", stream: Optional[int] = None):
        self._h = C.c_void_p()
        self.lib = load_library()
        self.N, self.F, self.G = int(N), int(F), len(wl)
        self.n_env, self.device = int(n_env), int(device)
        self.wl, self.dx, self.z = tuple(float(w) for w in wl), float(dx), float(z)
        self.pad, self.relative, self.method = int(pad), bool(relative), method
        self.Fg = self.F // self.G
        self.num_pixels = self.F * self.N * self.N
        wl_arr = (C.c_double * self.G)(*self.wl)
        rc = self.lib.bh_create(C.byref(self._h), self.device, self.n_env, self.N, self.F, self.G,
                                wl_arr, self.dx, self.z, self.pad, int(self.relative),
                                METHOD_ASM if method == "asm" else METHOD_FRESNEL)
        if rc != 0:
            msg = self.lib.bh_last_error(None).decode()
            self._h = C.c_void_p()
            raise HoloError(f"bh_create failed ({rc}): {msg}")
        if stream is not None:
            self.set_stream(stream)

    # -- plumbing ---------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise HoloError(f"{what} failed ({rc}): {self.lib.bh_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.bh_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.bh_set_stream(self._h, C.c_void_p(int(cuda_stream))), "bh_set_stream")

    @property
    def max_tasks(self) -> int:
        return self.lib.bh_max_tasks(self._h)

    @property
    def launch_count(self) -> int:
        return int(self.lib.bh_launch_count(self._h))

    def device_ptr(self, which: str) -> int:
        idx = {"U": 0, "I": 1, "T": 2, "state": 3, "sums": 4, "h": 5, "H": 6}[which]
        return int(self.lib.bh_device_ptr(self._h, idx) or 0)

    # -- reset path -------------------------------------------------------
    def set_target(self, env: int, target: np.ndarray):
        t = np.ascontiguousarray(target, dtype=np.float32).reshape(self.G, self.N, self.N)
        self._check(self.lib.bh_set_target(self._h, env, _ptr(t), 1), "bh_set_target")

    def load_state(self, env: int, state: np.ndarray):
        s = np.ascontiguousarray(state, dtype=np.int8).reshape(self.F, self.N, self.N)
        self._check(self.lib.bh_load_state(self._h, env, _ptr(s), 1), "bh_load_state")

    def clone_env(self, src: int, dst: int):
        """Device-side copy of one env's complete state (group rollouts from one reset state)."""
        self._check(self.lib.bh_clone_env(self._h, src, dst), "bh_clone_env")

    def resync(self, env: int):
        self._check(self.lib.bh_resync(self._h, env), "bh_resync")

    def metrics(self, env: int):
        """(psnr, mse, (sum I^2, sum I*T, sum T^2))."""
        p, m = C.c_double(), C.c_double()
        s3 = (C.c_double * 3)()
        self._check(self.lib.bh_get_metrics(self._h, env, C.byref(p), C.byref(m), s3), "bh_get_metrics")
        return p.value, m.value, (s3[0], s3[1], s3[2])

    # -- incremental path -------------------------------------------------
    def eval_flips(self, actions: np.ndarray, env: int = 0,
                   env_ids: Optional[np.ndarray] = None) -> np.ndarray:
        a = np.ascontiguousarray(actions, dtype=np.int64).ravel()
        e = None if env_ids is None else np.ascontiguousarray(env_ids, dtype=np.int32).ravel()
        out = np.empty(a.shape[0], dtype=np.float64)
        self._check(self.lib.bh_eval_flips(self._h, env, a.shape[0], _ptr(e), _ptr(a), _ptr(out)),
                    "bh_eval_flips")
        return out

    def step_batch(self, actions: np.ndarray, env_ids: Optional[np.ndarray] = None,
                   rule: int = RULE_ENV, out: Optional[np.ndarray] = None) -> np.ndarray:
        a = np.ascontiguousarray(actions, dtype=np.int64).ravel()
        e = None if env_ids is None else np.ascontiguousarray(env_ids, dtype=np.int32).ravel()
        if out is None:
            out = np.empty(a.shape[0], dtype=RESULT_DTYPE)
        self._check(self.lib.bh_step_batch(self._h, a.shape[0], _ptr(e), _ptr(a), rule, _ptr(out)),
                    "bh_step_batch")
        return out

    def vec_step(self, actions: np.ndarray, env_ids: np.ndarray, rule: int, out: np.ndarray,
                 book: "VecBook"):
        """One vectorised env step incl. the host bookkeeping (arrays must be C-contiguous)."""
        rc = self.lib.bh_vec_step(self._h, actions.shape[0], env_ids.ctypes.data, actions.ctypes.data,
                                  rule, out.ctypes.data, C.addressof(book))
        if rc != 0:
            self._check(rc, "bh_vec_step")
        return out

    def step_batch_ptrs(self, n: int, ptrs, rule: int):
        """``step_batch`` on cached addresses ``(env_ids, actions, results)`` of persistent host arrays
        (``ndarray.ctypes`` costs more than a microsecond per use; this is the per-step path of an env)."""
        rc = self.lib.bh_step_batch(self._h, n, ptrs[0], ptrs[1], rule, ptrs[2])
        if rc != 0:
            self._check(rc, "bh_step_batch")

    def vec_step_ptrs(self, n: int, ptrs, rule: int, book: "VecBook"):
        """``vec_step`` on cached addresses ``(env_ids, actions, results)`` of persistent host arrays."""
        rc = self.lib.bh_vec_step(self._h, n, ptrs[0], ptrs[1], rule, ptrs[2], self._book_ref(book))
        if rc != 0:
            self._check(rc, "bh_vec_step")

    def _book_ref(self, book: "VecBook"):
        ref = getattr(book, "_cached_ref", None)
        if ref is None:
            ref = book._cached_ref = C.addressof(book)
        return ref

    def step_batch_device(self, n: int, d_env_ids: int, d_actions: int, rule: int, d_results: int):
        self._check(self.lib.bh_step_batch_device(self._h, n, C.c_void_p(d_env_ids),
                                                  C.c_void_p(d_actions), rule, C.c_void_p(d_results)),
                    "bh_step_batch_device")

    def rollout_device(self, n_env: int, d_env_ids: int, d_actions: int, steps: int, rule: int, d_results: int = 0,
                       act_strides=None, res_strides=None):
        """``steps`` sequential flips per environment, actions known in advance, in one persistent launch
        (``bh_rollout_device``).  Default layout: step-major ``[steps][n_env]`` lists (strides in elements)."""
        a_s, a_e = act_strides if act_strides is not None else (n_env, 1)
        r_s, r_e = res_strides if res_strides is not None else (n_env, 1)
        self._check(self.lib.bh_rollout_device(self._h, n_env, C.c_void_p(d_env_ids or None), C.c_void_p(d_actions),
                                               a_s, a_e, steps, rule, C.c_void_p(d_results or None), r_s, r_e),
                    "bh_rollout_device")

    def rollout_status(self):
        """Synchronise and raise if the last rollout aborted at a barrier."""
        self._check(self.lib.bh_rollout_status(self._h), "bh_rollout_status")

    def eval_flips_device(self, n: int, d_env_ids: int, d_actions: int, d_results: int, env: int = 0):
        self._check(self.lib.bh_eval_flips_device(self._h, env, n, C.c_void_p(d_env_ids or None),
                                                  C.c_void_p(d_actions), C.c_void_p(d_results)),
                    "bh_eval_flips_device")

    def commit_flip(self, env: int, action: int):
        self._check(self.lib.bh_commit_flip(self._h, env, int(action)), "bh_commit_flip")

    def dbs_run(self, order: np.ndarray, env: int = 0, k_spec: int = 0, resync_every: int = 0,
                trace: bool = False):
        """Greedy DBS over ``order``.  Returns (accepted uint8[n], psnr_trace|None, n_acc, psnr)."""
        o = np.ascontiguousarray(order, dtype=np.int64).ravel()
        n = o.shape[0]
        acc = np.zeros(n, dtype=np.uint8)
        tr = np.zeros(n, dtype=np.float64) if trace else None
        nacc, fin = C.c_int64(0), C.c_double(0.0)
        self._check(self.lib.bh_dbs_run(self._h, env, _ptr(o), n, k_spec, resync_every, _ptr(acc),
                                        _ptr(tr), C.byref(nacc), C.byref(fin)), "bh_dbs_run")
        return acc, tr, int(nacc.value), fin.value

    def dbs_run_batch(self, orders: np.ndarray, env_ids: Optional[np.ndarray] = None, resync_every: int = 0,
                      trace: bool = False):
        """Greedy DBS of several images at once: ``orders`` (E, n), one candidate order per environment.
        Returns (accepted uint8 (E, n), psnr_trace (E, n) | None, n_accepted int64 (E,), final psnr (E,))."""
        o = np.ascontiguousarray(orders, dtype=np.int64)
        E, n = o.shape
        ids = None if env_ids is None else np.ascontiguousarray(env_ids, dtype=np.int32)
        acc = np.zeros((E, n), dtype=np.uint8)
        tr = np.zeros((E, n), dtype=np.float64) if trace else None
        nacc, fin = np.zeros(E, dtype=np.int64), np.zeros(E, dtype=np.float64)
        self._check(self.lib.bh_dbs_run_batch(self._h, E, _ptr(ids), _ptr(o), n, resync_every, _ptr(acc), _ptr(tr),
                                              _ptr(nacc), _ptr(fin)), "bh_dbs_run_batch")
        return acc, tr, nacc, fin

    def sweep_all(self, env: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        """PSNR after flipping each pixel of each frame (fixed state): float64 (F, N, N)."""
        if out is None:
            out = np.empty((self.F, self.N, self.N), dtype=np.float64)
        self._check(self.lib.bh_sweep_all(self._h, env, _ptr(out), 1), "bh_sweep_all")
        return out

    def sweep_stats(self, pre_model: np.ndarray, edges: np.ndarray, env: int = 0, want_map: bool = False):
        """Exhaustive sweep + decile statistics on the device.

        Returns (attempted int64[10], improved int64[10], gains float64[10], psnr_map | None).
        """
        pre = np.ascontiguousarray(pre_model, dtype=np.float32).reshape(self.F, self.N, self.N)
        ed = np.ascontiguousarray(edges, dtype=np.float64)
        assert ed.shape == (11,)
        att, imp = np.zeros(10, dtype=np.int64), np.zeros(10, dtype=np.int64)
        gains = np.zeros(10, dtype=np.float64)
        pm = np.empty((self.F, self.N, self.N), dtype=np.float64) if want_map else None
        self._check(self.lib.bh_sweep_stats(self._h, env, _ptr(pre), _ptr(ed), _ptr(att), _ptr(imp),
                                            _ptr(gains), _ptr(pm)), "bh_sweep_stats")
        return att, imp, gains, pm

    def sweep_all_device(self, env: int, d_out: int):
        self._check(self.lib.bh_sweep_all(self._h, env, C.c_void_p(d_out), 0), "bh_sweep_all")

    # -- read-back --------------------------------------------------------
    def recon(self, env: int = 0, candidate_action: int = -1,
              out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.G, self.N, self.N), dtype=np.float32)
        self._check(self.lib.bh_get_recon(self._h, env, _ptr(out), 1, int(candidate_action)),
                    "bh_get_recon")
        return out

    def recon_batch(self, n: int, out: int, kind: int = OBS_PINNED_HOST, buffer: int = 0, flags: int = OBS_SYNC,
                    env_ids_ptr: int = 0, d_results: int = 0):
        """obs["recon_image"] of the ``n`` envs of the step just enqueued, written plane by plane into the
        observation block at address ``out`` (see ``bh_recon_batch`` in include/bholo.h)."""
        rc = self.lib.bh_recon_batch(self._h, n, env_ids_ptr or None, d_results or None, out or None,
                                     kind, buffer, flags)
        if rc != 0:
            self._check(rc, "bh_recon_batch")

    def recon_device_block(self, buffer: int = 0) -> "DeviceArray":
        """Context-owned device observation block ``buffer`` as a zero-copy (E, 1, G, N, N) float32 view."""
        ptr = self.lib.bh_recon_device_block(self._h, buffer)
        if not ptr:
            raise HoloError("bh_recon_device_block failed: " + self.lib.bh_last_error(self._h).decode())
        return DeviceArray(int(ptr), (self.n_env, 1, self.G, self.N, self.N), "<f4", owner=self)

    @property
    def recon_planes_written(self) -> int:
        return int(self.lib.bh_recon_planes_written(self._h))

    def stream_sync(self):
        self._check(self.lib.bh_stream_sync(self._h), "bh_stream_sync")

    def state(self, env: int = 0) -> np.ndarray:
        out = np.empty((self.F, self.N, self.N), dtype=np.int8)
        self._check(self.lib.bh_get_state(self._h, env, _ptr(out), 1), "bh_get_state")
        return out

    def field(self, env: int, frame: int) -> np.ndarray:
        out = np.empty((self.N, self.N), dtype=np.complex64)
        self._check(self.lib.bh_get_field(self._h, env, frame, _ptr(out), 1), "bh_get_field")
        return out

    # -- timing hooks -----------------------------------------------------
    def time_eval(self, n: int, d_env_ids: int, d_actions: int, n_sets: int, reps: int) -> float:
        ms = C.c_float(0.0)
        self._check(self.lib.bh_time_eval(self._h, n, C.c_void_p(d_env_ids or None),
                                          C.c_void_p(d_actions), n_sets, reps, C.byref(ms)),
                    "bh_time_eval")
        return float(ms.value)

    def time_step(self, n: int, d_env_ids: int, d_actions: int, n_sets: int, reps: int, rule: int = RULE_ENV,
                  with_commit: bool = True) -> float:
        ms = C.c_float(0.0)
        self._check(self.lib.bh_time_step(self._h, n, C.c_void_p(d_env_ids), C.c_void_p(d_actions), n_sets,
                                          reps, rule, int(with_commit), C.byref(ms)), "bh_time_step")
        return float(ms.value)

    def time_commit(self, n: int, d_env_ids: int, d_actions: int, n_sets: int, reps: int) -> float:
        ms = C.c_float(0.0)
        self._check(self.lib.bh_time_commit(self._h, n, C.c_void_p(d_env_ids), C.c_void_p(d_actions), n_sets,
                                            reps, C.byref(ms)), "bh_time_commit")
        return float(ms.value)

    def time_propagate_passes(self, env: int, reps: int):
        ms = (C.c_float * 4)()
        self._check(self.lib.bh_time_propagate_passes(self._h, env, reps, ms), "bh_time_propagate_passes")
        return [float(x) for x in ms]

    def time_propagate(self, env: int, reps: int) -> float:
        ms = C.c_float(0.0)
        self._check(self.lib.bh_time_propagate(self._h, env, reps, C.byref(ms)), "bh_time_propagate")
        return float(ms.value)


def simulate(field: np.ndarray, wl: float, z: float = 2e-3, dx: float = 7.56e-6, pad: int = 1,
             method: str = "asm", device: int = 0) -> np.ndarray:
    """``tt.simulate`` on host arrays: (..., N, N) real or complex -> complex64."""
    lib = load_library()
    x = np.asarray(field)
    N = x.shape[-1]
    lead = x.shape[:-2]
    cplx = np.iscomplexobj(x)
    xin = np.ascontiguousarray(x, dtype=np.complex64 if cplx else np.float32).reshape(-1, N, N)
    out = np.empty(xin.shape, dtype=np.complex64)
    rc = lib.bh_simulate(device, None, _ptr(xin), int(cplx), xin.shape[0], N, float(wl), float(dx),
                         float(z), int(pad), METHOD_ASM if method == "asm" else METHOD_FRESNEL,
                         _ptr(out), 1)
    if rc != 0:
        raise HoloError(f"bh_simulate failed ({rc}): {lib.bh_last_error(None).decode()}")
    return out.reshape(lead + (N, N))
