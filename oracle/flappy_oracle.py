"""ctypes wrapper around oracle/libflappy_oracle.so -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module (see flappy_oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libflappy_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "flappy_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libflappy_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.fo_assets_load.argtypes = [C.c_char_p, C.c_size_t]
        L.fo_assets_load.restype = C.c_int
        L.fo_env_create.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        L.fo_env_create.restype = C.c_void_p
        L.fo_env_destroy.argtypes = [C.c_void_p]
        L.fo_env_step.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)]
        L.fo_env_step.restype = C.c_int
        L.fo_env_export_state.argtypes = [C.c_void_p, C.c_void_p]
        L.fo_env_import_state.argtypes = [C.c_void_p, C.c_void_p]
        L.fo_env_render.argtypes = [C.c_void_p, C.c_void_p]
        L.fo_preprocess.argtypes = [C.c_void_p, C.c_void_p]
        L.fo_env_obs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fo_resize_tables.argtypes = [C.c_void_p]
        L.fo_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fo_stream_word.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32]
        L.fo_stream_word.restype = C.c_uint32
        L.fo_batch_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.fo_batch_step.restype = C.c_int
        _lib = L
    return _lib


_assets_loaded = False


def load_assets(blob: bytes | None = None):
    global _assets_loaded
    if blob is None:
        if _assets_loaded:
            return
        import sys
        sys.path.insert(0, os.path.dirname(_HERE))
        from dqnflappybird_b200.assets import load_blob
        blob = load_blob()
    rc = lib().fo_assets_load(blob, len(blob))
    if rc != 0:
        raise RuntimeError(f"fo_assets_load failed: {rc}")
    _assets_loaded = True


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    lib().fo_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def stream_word(seed: int, purpose: int, env: int, n: int) -> int:
    return int(lib().fo_stream_word(seed, purpose, env, n))


def resize_tables() -> np.ndarray:
    """i32[6][80]: sx, a0, a1 (game x / cv2 rows), sy, b0, b1 (game y / cv2 cols)."""
    t = np.zeros((6, 80), np.int32)
    lib().fo_resize_tables(t.ctypes.data)
    return t


def preprocess(frame: np.ndarray) -> np.ndarray:
    """u8[288][512][3] array3d frame -> u8[80][80] (FlappyBirdDQN.py:31-34)."""
    f = np.ascontiguousarray(frame, np.uint8)
    assert f.shape == (288, 512, 3)
    out = np.empty((80, 80), np.uint8)
    lib().fo_preprocess(f.ctypes.data, out.ctypes.data)
    return out


class OracleEnvs:
    """N independent reference envs (one reference process each).

    gaps: optional u8[N][G] scripted gap indices (replay mode); otherwise the
    Philox(seed, env) stream with CPython's randint(0,7) rejection rule.
    """

    def __init__(self, n: int, seed: int = 0, gaps: np.ndarray | None = None, first_env_id: int = 0):
        load_assets()
        self.n = n
        L = lib()
        self._h = []
        for k in range(n):
            if gaps is not None:
                g = np.ascontiguousarray(gaps[k], np.uint8)
                h = L.fo_env_create(seed, first_env_id + k, g.ctypes.data, g.size)
            else:
                h = L.fo_env_create(seed, first_env_id + k, None, 0)
            if not h:
                raise RuntimeError("fo_env_create failed")
            self._h.append(h)
        self._arr = (C.c_void_p * n)(*self._h)

    def __del__(self):
        try:
            for h in self._h:
                lib().fo_env_destroy(h)
        except Exception:
            pass

    def step(self, actions, want_obs: bool = True, threads: int = 1):
        a = np.ascontiguousarray(actions, np.uint8)
        assert a.shape == (self.n,)
        r = np.empty(self.n, np.float32); t = np.empty(self.n, np.uint8); s = np.empty(self.n, np.int32)
        obs = np.empty((self.n, 80, 80), np.uint8) if want_obs else None
        rc = lib().fo_batch_step(self._arr, self.n, a.ctypes.data, r.ctypes.data, t.ctypes.data, s.ctypes.data,
                                 obs.ctypes.data if want_obs else None, threads)
        if rc != 0:
            raise ValueError("Multiple input actions!")
        return obs, r, t, s

    def run(self, k: int, actions, want_obs=None):
        """T consecutive steps of env k in one C call: (reward[T], terminal[T], score[T], state[T][16], obs[n_wanted][80][80])"""
        a = np.ascontiguousarray(actions, np.uint8)
        T = len(a)
        w = np.ascontiguousarray(want_obs, np.uint8) if want_obs is not None else np.zeros(T, np.uint8)
        r = np.empty(T, np.float32); t = np.empty(T, np.uint8); s = np.empty(T, np.int32); st = np.empty((T, 16), np.int32)
        obs = np.empty((int(w.sum()), 80, 80), np.uint8)
        L = lib()
        L.fo_env_run.argtypes = [C.c_void_p] + [C.c_int] + [C.c_void_p] * 7
        rc = L.fo_env_run(self._h[k], T, a.ctypes.data, w.ctypes.data, r.ctypes.data, t.ctypes.data, s.ctypes.data, st.ctypes.data,
                          obs.ctypes.data)
        if rc != 0:
            raise ValueError("Multiple input actions!")
        return r, t, s, st, obs

    def export_state(self) -> np.ndarray:
        out = np.zeros((self.n, 16), np.int32)
        for k, h in enumerate(self._h):
            lib().fo_env_export_state(h, out[k].ctypes.data)
        return out

    def import_state(self, state: np.ndarray):
        st = np.ascontiguousarray(state, np.int32)
        assert st.shape == (self.n, 16)
        for k, h in enumerate(self._h):
            lib().fo_env_import_state(h, st[k].ctypes.data)

    def obs_all(self) -> np.ndarray:
        out = np.empty((self.n, 80, 80), np.uint8)
        for k in range(self.n):
            out[k] = self.obs(k)
        return out

    def render_full(self, k: int) -> np.ndarray:
        f = np.empty((288, 512, 3), np.uint8)
        lib().fo_env_render(self._h[k], f.ctypes.data)
        return f

    def obs(self, k: int) -> np.ndarray:
        return preprocess(self.render_full(k))
