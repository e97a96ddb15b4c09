"""ctypes mirror of include/rlfc.h.

`AFCCylinderBatch` plays the role of the reference's `AFCCylinder` object (AFCCylinder.pde:1-61) for a
whole batch of environments: `update2()` is one solver step, `step(actions)` is one RL step as
clientCFD.draw() drives it (clientCFD.pde:35-55).  All numerics run in librlfc.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
# RLFC_LIBRARY selects another build of the same library (e.g. an instrumented one made by tools/); never a fallback
_LIB_PATH = Path(os.environ["RLFC_LIBRARY"]).resolve() if os.environ.get("RLFC_LIBRARY") else _PKG / "librlfc.so"
NUM_PROBES = 32


class RlfcError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("resolution", C.c_int), ("x_lengths", C.c_int), ("y_lengths", C.c_int), ("re", C.c_int),
        ("dR", C.c_float), ("gR", C.c_float), ("theta", C.c_float), ("t_step", C.c_float),
        ("action_scale", C.c_float), ("substeps", C.c_int), ("init_time", C.c_float),
        ("episode_time", C.c_float), ("n_envs", C.c_int), ("device", C.c_int), ("exact", C.c_int),
        ("mg_max_iters", C.c_int), ("init_bdim_path", C.c_char_p), ("stream", C.c_void_p),
        ("n_groups", C.c_int), ("n_devices", C.c_int),
    ]


_lib = None


def library_path() -> Path:
    return _LIB_PATH


def default_init_state() -> Path:
    """The developed-wake state every reference episode resumes from (saved/init/init.bdim,
    AFCCylinder.pde:35-37), shipped here in binary form."""
    return _PKG / "data" / "init_state.bdimb"


def load_library():
    """dlopen librlfc.so; fails loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RlfcError(f"{_LIB_PATH} is missing: build it with `python -m rlfluidcontrol_b200.build` "
                        "(there is no CPU fallback)")
    L = C.CDLL(str(_LIB_PATH))
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
    L.rlfc_default_config.argtypes = [C.POINTER(Config)]
    L.rlfc_default_config.restype = None
    L.rlfc_env_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.rlfc_env_destroy.argtypes = [vp]
    L.rlfc_env_destroy.restype = None
    L.rlfc_env_reset.argtypes = [vp, ip, C.c_int, C.c_int]
    L.rlfc_env_step.argtypes = [vp, fp, fp, fp, ip]
    L.rlfc_env_step_device.argtypes = [vp, vp, vp, vp, vp]
    L.rlfc_env_substep.argtypes = [vp, fp, fp, fp]
    L.rlfc_env_substep_device.argtypes = [vp, vp]
    L.rlfc_env_get_fields.argtypes = [vp, C.c_int, fp, fp, fp]
    L.rlfc_env_set_fields.argtypes = [vp, C.c_int, fp, fp, fp]
    L.rlfc_env_save_bdim.argtypes = [vp, C.c_int, C.c_char_p]
    L.rlfc_env_load_bdim.argtypes = [vp, C.c_int, C.c_char_p]
    L.rlfc_env_dims.argtypes = [vp, ip, ip, ip]
    L.rlfc_env_get_time.argtypes = [vp, fp]
    L.rlfc_env_get_mg_iters.argtypes = [vp, ip]
    L.rlfc_env_check_cfl.argtypes = [vp, fp]
    L.rlfc_env_running.argtypes = [vp, ip]
    L.rlfc_env_get_flags.argtypes = [vp, ip]
    L.rlfc_env_field_sum.argtypes = [vp, fp]
    L.rlfc_env_field_sum_stats.argtypes = [vp, ip]
    L.rlfc_env_get_static.argtypes = [vp, C.c_char_p, C.c_int, fp, ip, ip]
    L.rlfc_geometry_static.argtypes = [C.POINTER(Config), C.c_char_p, C.c_int, fp, ip, ip, ip]
    L.rlfc_env_num_levels.argtypes = [vp]
    L.rlfc_env_set_profiling.argtypes = [vp, C.c_int]
    L.rlfc_env_get_profile.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double),
                                       C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
    L.rlfc_env_stream.argtypes = [vp]
    L.rlfc_env_stream.restype = vp
    L.rlfc_env_slab_info.argtypes = [vp, ip, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.rlfc_env_launch_count.argtypes = [vp]
    L.rlfc_env_launch_count.restype = C.c_longlong
    L.rlfc_env_model_bytes_per_solver_step.argtypes = [vp]
    L.rlfc_env_model_bytes_per_solver_step.restype = C.c_double
    L.rlfc_last_error.restype = C.c_char_p
    L.rlfc_version.restype = C.c_char_p
    _lib = L
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


class AFCCylinderBatch:
    """A batch of AFCCylinder environments on one GPU."""

    def __init__(self, n_envs=1, init_state="default", device=-1, stream=None, **overrides):
        self._L = load_library()
        cfg = Config()
        self._L.rlfc_default_config(C.byref(cfg))
        cfg.n_envs = int(n_envs)
        cfg.device = int(device)
        if init_state == "default":
            init_state = default_init_state()
        self._init_path = None if init_state is None else str(init_state).encode()
        cfg.init_bdim_path = self._init_path
        if stream is not None:
            cfg.stream = C.c_void_p(int(stream))
        for k, v in overrides.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown config field {k}")
            setattr(cfg, k, v)
        self.cfg = cfg
        h = C.c_void_p()
        rc = self._L.rlfc_env_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise RlfcError(f"rlfc_env_create failed ({rc}): {self._L.rlfc_last_error().decode()}")
        self._h = h
        n, m, b = C.c_int(), C.c_int(), C.c_int()
        self._L.rlfc_env_dims(h, C.byref(n), C.byref(m), C.byref(b))
        self.n, self.m, self.n_envs = n.value, m.value, b.value

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.rlfc_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise RlfcError(f"{what} failed ({rc}): {self._L.rlfc_last_error().decode()}")

    # -- reference-facing interface ---------------------------------------------------------
    def reset(self, env_ids=None, reset_accumulators=False):
        if env_ids is None:
            rc = self._L.rlfc_env_reset(self._h, None, self.n_envs, int(reset_accumulators))
        else:
            ids = np.ascontiguousarray(env_ids, dtype=np.int32)
            rc = self._L.rlfc_env_reset(self._h, ids.ctypes.data_as(C.POINTER(C.c_int)), len(ids), int(reset_accumulators))
        self._check(rc, "rlfc_env_reset")

    def step(self, actions, want_reward=True, want_done=True):
        """One RL step: actions (n_envs, 2) in [-1, 1] -> obs (n_envs, 2) = (Cl, Cd), reward, done."""
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n_envs, 2)
        obs = np.empty((self.n_envs, 2), np.float32)
        rew = np.empty(self.n_envs, np.float32) if want_reward else None
        done = np.empty(self.n_envs, np.int32) if want_done else None
        rc = self._L.rlfc_env_step(self._h, _fp(a), _fp(obs), _fp(rew),
                                   done.ctypes.data_as(C.POINTER(C.c_int)) if done is not None else None)
        self._check(rc, "rlfc_env_step")
        return obs, rew, done

    def step_device(self, d_actions, d_obs, d_reward=0, d_done=0):
        """Asynchronous RL step on device pointers (ints, e.g. torch.Tensor.data_ptr())."""
        rc = self._L.rlfc_env_step_device(self._h, C.c_void_p(d_actions), C.c_void_p(d_obs),
                                          C.c_void_p(d_reward) if d_reward else None,
                                          C.c_void_p(d_done) if d_done else None)
        self._check(rc, "rlfc_env_step_device")

    def update2(self, actions=None, want_probes=False):
        """One solver step (AFCCylinder.update2).  Returns force (n_envs, 2) [, probes (n_envs, 32)]."""
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n_envs, 2)
        force = np.empty((self.n_envs, 2), np.float32)
        probes = np.empty((self.n_envs, NUM_PROBES), np.float32) if want_probes else None
        rc = self._L.rlfc_env_substep(self._h, _fp(a), _fp(force), _fp(probes))
        self._check(rc, "rlfc_env_substep")
        return (force, probes) if want_probes else force

    def update2_device(self, d_actions=0):
        """One solver step, asynchronous on the handle's stream; d_actions = device pointer (int) or 0 to keep xi."""
        self._check(self._L.rlfc_env_substep_device(self._h, C.c_void_p(d_actions) if d_actions else None), "rlfc_env_substep_device")

    # -- state / introspection --------------------------------------------------------------
    def get_fields(self, e=0):
        ux, uy, p = (np.empty((self.n, self.m), np.float32) for _ in range(3))
        self._check(self._L.rlfc_env_get_fields(self._h, e, _fp(ux), _fp(uy), _fp(p)), "rlfc_env_get_fields")
        return ux, uy, p

    def set_fields(self, e, ux=None, uy=None, p=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float32).reshape(self.n, self.m) for a in (ux, uy, p)]
        self._check(self._L.rlfc_env_set_fields(self._h, e, *[_fp(a) for a in arrs]), "rlfc_env_set_fields")

    def save_bdim(self, e, path):
        self._check(self._L.rlfc_env_save_bdim(self._h, e, str(path).encode()), "rlfc_env_save_bdim")

    def load_bdim(self, e, path):
        self._check(self._L.rlfc_env_load_bdim(self._h, e, str(path).encode()), "rlfc_env_load_bdim")

    @property
    def t(self):
        t = np.empty(self.n_envs, np.float32)
        self._check(self._L.rlfc_env_get_time(self._h, _fp(t)), "rlfc_env_get_time")
        return t

    def check_cfl(self):
        """BDIM.checkCFL per environment: min(1/(max(|ux|+|uy|) + 3 nu), 1)."""
        dt = np.empty(self.n_envs, np.float32)
        self._check(self._L.rlfc_env_check_cfl(self._h, _fp(dt)), "rlfc_env_check_cfl")
        return dt

    def running(self):
        """Environments that had not emitted their observation when the last RL-step round ended."""
        n = C.c_int()
        self._check(self._L.rlfc_env_running(self._h, C.byref(n)), "rlfc_env_running")
        return n.value

    def flags(self):
        """Per-environment health flags (bit 0: a non-finite force was produced since the last reset)."""
        f = np.empty(self.n_envs, np.int32)
        self._check(self._L.rlfc_env_get_flags(self._h, f.ctypes.data_as(C.POINTER(C.c_int))), "rlfc_env_get_flags")
        return f

    def mg_iters(self):
        it = np.empty((self.n_envs, 2), np.int32)
        self._check(self._L.rlfc_env_get_mg_iters(self._h, it.ctypes.data_as(C.POINTER(C.c_int))), "rlfc_env_get_mg_iters")
        return it

    def field_sum(self):
        """Field.sum() of every env's pressure field (Field.pde:311-318), evaluated on the device."""
        s = np.empty(self.n_envs, np.float32)
        self._check(self._L.rlfc_env_field_sum(self._h, _fp(s)), "rlfc_env_field_sum")
        return s

    def field_sum_stats(self):
        """Counters of the last Field.sum evaluation per env (see rlfc_env_field_sum_stats)."""
        st = np.zeros((self.n_envs, 8), np.int32)
        self._check(self._L.rlfc_env_field_sum_stats(self._h, st.ctypes.data_as(C.POINTER(C.c_int))), "rlfc_env_field_sum_stats")
        return st

    @property
    def num_levels(self):
        return self._L.rlfc_env_num_levels(self._h)

    def get_static(self, name, level=0):
        n, m = C.c_int(), C.c_int()
        self._check(self._L.rlfc_env_get_static(self._h, name.encode(), level, None, C.byref(n), C.byref(m)), "rlfc_env_get_static")
        out = np.empty((n.value, m.value), np.float32)
        self._check(self._L.rlfc_env_get_static(self._h, name.encode(), level, _fp(out), None, None), "rlfc_env_get_static")
        return out

    def set_profiling(self, on=True):
        self._check(self._L.rlfc_env_set_profiling(self._h, int(on)), "rlfc_env_set_profiling")

    def get_profile(self):
        """[{name, ms, launches, bytes_per_launch}] accumulated since profiling was switched on."""
        out, idx = [], 0
        while True:
            name = C.create_string_buffer(64)
            ms, cnt, by = C.c_double(), C.c_longlong(), C.c_double()
            rc = self._L.rlfc_env_get_profile(self._h, idx, name, 64, C.byref(ms), C.byref(cnt), C.byref(by))
            if rc == 1:
                break
            self._check(rc, "rlfc_env_get_profile")
            out.append(dict(name=name.value.decode(), ms=ms.value, launches=cnt.value, bytes_per_launch=by.value))
            idx += 1
        return out

    @property
    def stream(self):
        return self._L.rlfc_env_stream(self._h)

    @property
    def launch_count(self):
        return int(self._L.rlfc_env_launch_count(self._h))

    def slab_info(self):
        """(devices sharing the domain, device-wide barriers executed so far, bytes of the shared address range)."""
        n, b, by = C.c_int(), C.c_longlong(), C.c_longlong()
        self._check(self._L.rlfc_env_slab_info(self._h, C.byref(n), C.byref(b), C.byref(by)), "rlfc_env_slab_info")
        return n.value, b.value, by.value

    def model_bytes_per_solver_step(self):
        return float(self._L.rlfc_env_model_bytes_per_solver_step(self._h))


def geometry_static(name, level=0, **overrides):
    """Host-side static geometry/coefficient field for a configuration (no GPU needed)."""
    L = load_library()
    cfg = Config()
    L.rlfc_default_config(C.byref(cfg))
    for k, v in overrides.items():
        setattr(cfg, k, v)
    n, m, nl = C.c_int(), C.c_int(), C.c_int()
    rc = L.rlfc_geometry_static(C.byref(cfg), name.encode(), level, None, C.byref(n), C.byref(m), C.byref(nl))
    if rc:
        raise RlfcError(f"rlfc_geometry_static failed ({rc}): {L.rlfc_last_error().decode()}")
    out = np.empty((n.value, m.value), np.float32)
    rc = L.rlfc_geometry_static(C.byref(cfg), name.encode(), level, _fp(out), None, None, None)
    if rc:
        raise RlfcError(f"rlfc_geometry_static failed ({rc}): {L.rlfc_last_error().decode()}")
    return out, nl.value
