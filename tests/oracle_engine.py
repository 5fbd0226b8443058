"""TEST INFRASTRUCTURE ONLY: a stand-in for ``HoloEngine`` backed by the CPU oracle.

The `-m "not gpu"` suite uses it (monkeypatched into envs / vec_env / dbs) to run the Python host
logic -- rewards, bonuses, termination, auto-reset, group tables, DBS drivers, checkpoints -- on a
machine without a GPU, against ``oracle.hologram_oracle.OracleEnv``.  It is never imported by the
package: the product path has no CPU fallback (tests/test_host.py::test_no_cpu_fallback_without_gpu).
Every flip is scored by a full re-simulation of the flipped colour group in float64
(the reference's own structure, env.py:163-174), so it is only usable at toy sizes.
"""
import ctypes as C

import numpy as np

from binary_hologram_reinforcement_learning_b200 import engine as real
from oracle import hologram_oracle as O


class OracleEngine:
    def __init__(self, N, F, wl, n_env=1, device=0, dx=7.56e-6, z=2e-3, pad=1, relative=True,
                 method="asm"):
        self.cfg = O.HoloConfig(N=int(N), F=int(F), wl=tuple(wl), dx=float(dx), z=float(z), pad=int(pad),
                                relative=bool(relative), method=method)
        self.N, self.F, self.G, self.n_env = int(N), int(F), len(wl), int(n_env)
        self.dx, self.z, self.pad, self.relative = float(dx), float(z), int(pad), bool(relative)
        self.Fg = self.F // self.G
        self.num_pixels = self.F * self.N * self.N
        self.launch_count = 0
        self._state = [np.zeros((F, N, N), np.int8) for _ in range(n_env)]
        self._target = [np.zeros((self.G, N, N)) for _ in range(n_env)]
        self._means = [None] * n_env
        self._prev = [0.0] * n_env
        self._last = {}                      # env -> (group, mean_after) of its last REJECTED flip
        self.closed = False

    # -- context ---------------------------------------------------------------------------
    def close(self):
        self.closed = True

    def set_stream(self, s):
        pass

    @property
    def max_tasks(self):
        return 4096

    # -- state -----------------------------------------------------------------------------
    def set_target(self, env, target):
        self._target[env] = np.asarray(target, dtype=np.float64).reshape(self.G, self.N, self.N)

    def load_state(self, env, state):
        self._state[env] = np.asarray(state, dtype=np.int8).reshape(self.F, self.N, self.N).copy()
        self._repropagate(env)

    def _repropagate(self, env):
        self._last[env] = None
        self._means[env] = O.reconstruct(self.cfg, self._state[env])
        self._prev[env] = O.score(self.cfg, self._means[env], self._target[env])[0]

    def resync(self, env):
        self._repropagate(env)

    def clone_env(self, src, dst):
        self._state[dst] = self._state[src].copy()
        self._target[dst] = self._target[src].copy()
        self._means[dst] = self._means[src].copy()
        self._prev[dst] = self._prev[src]
        self._last[dst] = None

    def metrics(self, env):
        psnr, mse = O.score(self.cfg, self._means[env], self._target[env])
        return psnr, mse, np.array(O.loss_sums(self._means[env], self._target[env]))

    def state(self, env=0):
        return self._state[env].copy()

    # -- scoring ---------------------------------------------------------------------------
    def _flip(self, env, action):
        """(psnr_after, group, mean_after) of one candidate flip; the state is left unchanged."""
        cfg = self.cfg
        ch, r, c = cfg.decode(int(action))
        st = self._state[env]
        st[ch, r, c] = 1 - st[ch, r, c]
        g = cfg.group_of(ch)
        mean_after = O.group_mean_intensity(O.propagate_group(cfg, st[g * cfg.Fg:(g + 1) * cfg.Fg], g))
        st[ch, r, c] = 1 - st[ch, r, c]
        rec = self._means[env].copy()
        rec[g] = mean_after
        psnr = O.relative_loss(rec, self._target[env], O.get_psnr, cfg.relative)
        return psnr, g, mean_after

    def eval_flips(self, actions, env=0, env_ids=None):
        a = np.asarray(actions, dtype=np.int64).ravel()
        ids = np.full(a.shape, env) if env_ids is None else np.asarray(env_ids).ravel()
        return np.array([self._flip(int(e), int(x))[0] for e, x in zip(ids, a)], dtype=np.float64)

    def step_batch(self, actions, env_ids=None, rule=real.RULE_ENV, out=None):
        a = np.asarray(actions, dtype=np.int64).ravel()
        ids = np.arange(a.shape[0]) if env_ids is None else np.asarray(env_ids).ravel()
        if len(set(int(e) for e in ids)) != len(ids):
            raise real.HoloError("environment appears twice in one batch step")
        res = np.zeros(a.shape[0], dtype=real.RESULT_DTYPE) if out is None else out
        for i, (e, x) in enumerate(zip(ids, a)):
            e, x = int(e), int(x)
            if not 0 <= x < self.num_pixels:
                raise real.HoloError(f"action[{i}]={x} out of range")
            psnr, g, mean_after = self._flip(e, x)
            if rule == real.RULE_ENV:
                acc = not (psnr - self._prev[e] < 0.0)            # env.py:191
            elif rule == real.RULE_DBS:
                acc = psnr > self._prev[e]                        # DBS.py:273
            else:
                acc = False
            ch, r, c = self.cfg.decode(x)
            res["psnr_after"][i], res["accept"][i], res["action"][i] = psnr, int(acc), x
            res["sgn"][i] = 1 - 2 * int(self._state[e][ch, r, c])
            self._last[e] = None if acc else (g, mean_after)
            if acc:
                self._state[e][ch, r, c] = 1 - self._state[e][ch, r, c]
                self._means[e][g] = mean_after
                self._prev[e] = psnr
        self.launch_count += 2
        return res

    def recon_batch(self, n, out, kind=real.OBS_PINNED_HOST, buffer=0, flags=real.OBS_SYNC, env_ids_ptr=0,
                    d_results=0):
        """bh_recon_batch on host memory: every plane of the listed envs is rewritten (the plane-wise
        staleness bookkeeping is a device-side optimisation with the same visible result)."""
        assert kind == real.OBS_PINNED_HOST and out
        ids = (np.ctypeslib.as_array((C.c_int32 * n).from_address(env_ids_ptr)) if env_ids_ptr
               else np.arange(n))
        G, N = self.G, self.N
        blk = np.ctypeslib.as_array((C.c_float * (self.n_env * G * N * N)).from_address(out))
        blk = blk.reshape(self.n_env, G, N, N)
        for e in ids:
            e = int(e)
            rec = self._means[e].copy()
            last = self._last.get(e)
            if last is not None and not (flags & real.OBS_COMMITTED_ONLY):
                rec[last[0]] = last[1]
            blk[e] = rec

    def stream_sync(self):
        pass

    def _views(self, n, ptrs):
        ids = np.ctypeslib.as_array((C.c_int32 * n).from_address(ptrs[0]))
        acts = np.ctypeslib.as_array((C.c_int64 * n).from_address(ptrs[1]))
        res = np.frombuffer((C.c_char * (n * real.RESULT_DTYPE.itemsize)).from_address(ptrs[2]),
                            dtype=real.RESULT_DTYPE)
        return ids, acts, res

    def step_batch_ptrs(self, n, ptrs, rule):
        ids, acts, res = self._views(n, ptrs)
        self.step_batch(acts, ids, rule, out=res)

    def vec_step_ptrs(self, n, ptrs, rule, book):
        """Same contract as HoloEngine.vec_step_ptrs; the bookkeeping runs in the real C library."""
        ids, acts, res = self._views(n, ptrs)
        self.step_batch(acts, ids, rule, out=res)
        rc = real.load_library().bh_vec_book_update(n, ptrs[0], ptrs[1], ptrs[2], C.addressof(book))
        assert rc == 0

    def commit_flip(self, env, action):
        psnr, g, mean_after = self._flip(env, action)
        ch, r, c = self.cfg.decode(int(action))
        self._state[env][ch, r, c] = 1 - self._state[env][ch, r, c]
        self._means[env][g] = mean_after
        self._prev[env] = psnr

    def recon(self, env=0, candidate_action=-1, out=None):
        rec = self._means[env].copy()
        if candidate_action is not None and int(candidate_action) >= 0:
            _, g, mean_after = self._flip(env, int(candidate_action))
            rec[g] = mean_after
        if out is not None:
            out[...] = rec.reshape(out.shape)
            return out
        return rec.astype(np.float32)

    # -- DBS ---------------------------------------------------------------------------------
    def dbs_run(self, order, env=0, k_spec=0, resync_every=0, trace=False):
        o = np.asarray(order, dtype=np.int64).ravel()
        st, acc, tr = O.dbs_greedy(self.cfg, self._state[env], self._target[env], o)
        self._state[env] = st.astype(np.int8)
        self._repropagate(env)
        return acc.astype(np.uint8), (tr if trace else None), int(acc.sum()), self._prev[env]

    def sweep_all(self, env=0, out=None):
        psnr = self.eval_flips(np.arange(self.num_pixels), env=env).reshape(self.F, self.N, self.N)
        if out is not None:
            out[...] = psnr
            return out
        return psnr

    def sweep_stats(self, pre_model, edges, env=0, want_map=False):
        """dbs-1024-1024-24-6464.py:371-395 over every pixel: (attempted, improved, gains, psnr map)."""
        pm = self.sweep_all(env)
        pre = np.asarray(pre_model, dtype=np.float32).reshape(self.F, self.N, self.N)
        att, imp, gains = np.zeros(10, np.int64), np.zeros(10, np.int64), np.zeros(10)
        p0 = self._prev[env]
        for idx in np.ndindex(pre.shape):
            b = O.decile_of(float(pre[idx]))
            if b >= 0:
                att[b] += 1
                if pm[idx] > p0:
                    imp[b] += 1
                    gains[b] += pm[idx] - p0
        return att, imp, gains, (pm if want_map else None)


def pinned_stub(shape, dtype):
    return np.empty(shape, dtype=dtype)
