"""ctypes binding of libqecmc.so (include/qecmc.h).  There is no CPU fallback: if the
library is missing or no CUDA device is present, every entry point raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("QECMC_LIB") or os.path.join(HERE, "csrc", "libqecmc.so")

TORIC, PLANAR, ROTATED, XZZX = 0, 1, 2, 3
GEOM_NAMES = {"toric": TORIC, "planar": PLANAR, "rotated": ROTATED, "xzzx": XZZX}
POW_NUMBA, POW_LIBM = 0, 1


class QecmcError(RuntimeError):
    pass


class DevInfo(C.Structure):
    _fields_ = [("device", C.c_int32), ("sm_count", C.c_int32), ("sm_clock_khz", C.c_int32), ("cc_major", C.c_int32),
                ("cc_minor", C.c_int32), ("total_mem", C.c_int64), ("free_mem", C.c_int64), ("name", C.c_char * 64)]


class Stats(C.Structure):
    _fields_ = [("metropolis_steps", C.c_int64), ("accepted", C.c_int64), ("samples", C.c_int64),
                ("distinct", C.c_int64), ("table_slots", C.c_int64), ("waves", C.c_int64),
                ("kernel_launches", C.c_int64), ("chain_kernel_ms", C.c_double), ("total_ms", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class ChainCfg(C.Structure):
    _fields_ = [("geom_chain", C.c_int32), ("L", C.c_int32), ("pow_kind", C.c_int32), ("reserved", C.c_int32),
                ("p", C.c_double), ("seed", C.c_uint64), ("stream_offset", C.c_uint64)]


class StdcCfg(C.Structure):
    _fields_ = [("geom_code", C.c_int32), ("geom_chain", C.c_int32), ("L", C.c_int32), ("droplets", C.c_int32),
                ("iters", C.c_int32), ("per_class_inits", C.c_int32), ("randomize", C.c_int32), ("reserved", C.c_int32),
                ("steps", C.c_int64), ("p_error", C.c_double), ("p_sampling", C.c_double), ("conv_mult", C.c_double),
                ("seed", C.c_uint64), ("u_nb", C.c_void_p), ("u_np", C.c_void_p)]


class LadderCfg(C.Structure):
    _fields_ = [("geom", C.c_int32), ("L", C.c_int32), ("kind", C.c_int32), ("Nc", C.c_int32), ("iters", C.c_int32),
                ("reserved", C.c_int32), ("bottom", C.c_double), ("param_b", C.c_double), ("p_logical", C.c_double),
                ("seed", C.c_uint64), ("u_nb", C.c_void_p), ("u_py", C.c_void_p), ("n_nb", C.c_int64), ("n_py", C.c_int64)]


class LadderIO(C.Structure):
    _fields_ = [("qm0", C.c_void_p), ("resume", C.c_int32), ("reserved", C.c_int32), ("rung_states", C.c_void_p),
                ("flags", C.c_void_p), ("tops0", C.c_void_p), ("n_eff", C.c_void_p), ("n_eff_parts", C.c_void_p),
                ("snap_states", C.c_void_p), ("snap_flags", C.c_void_p), ("snap_tops0", C.c_void_p)]


class PteqCfg(C.Structure):
    _fields_ = [("ladder", LadderCfg), ("SEQ", C.c_int32), ("TOPS", C.c_int32), ("tops_burn", C.c_int32),
                ("use_conv", C.c_int32), ("eps", C.c_double), ("steps", C.c_int64)]


class PtdcCfg(C.Structure):
    _fields_ = [("ladder", LadderCfg), ("droplets", C.c_int32), ("per_class_inits", C.c_int32), ("steps", C.c_int64),
                ("p_error", C.c_double), ("conv_mult", C.c_double), ("steps_done", C.c_void_p)]


class AlphaCfg(C.Structure):
    _fields_ = [("geom", C.c_int32), ("L", C.c_int32), ("iters", C.c_int32), ("per_class_inits", C.c_int32),
                ("steps", C.c_int64), ("pz_tilde_sampling", C.c_double), ("alpha", C.c_double), ("pz_tilde", C.c_double),
                ("seed", C.c_uint64), ("u_nb", C.c_void_p), ("u_py", C.c_void_p), ("n_nb", C.c_int64), ("n_py", C.c_int64)]


class XyzCfg(C.Structure):
    _fields_ = [("geom_code", C.c_int32), ("geom_chain", C.c_int32), ("L", C.c_int32), ("droplets", C.c_int32),
                ("iters", C.c_int32), ("per_class_inits", C.c_int32), ("use_xyz_sampling", C.c_int32), ("reserved", C.c_int32),
                ("steps", C.c_int64), ("p_xyz", C.c_double * 3), ("p_sampling_xyz", C.c_double * 3), ("p_sampling", C.c_double),
                ("seed", C.c_uint64), ("u_nb", C.c_void_p)]


class NoiseCfg(C.Structure):
    _fields_ = [("geom", C.c_int32), ("L", C.c_int32), ("toric_form", C.c_int32), ("reserved", C.c_int32),
                ("p_error", C.c_double), ("p_x", C.c_double), ("p_y", C.c_double), ("p_z", C.c_double),
                ("seed", C.c_uint64), ("u", C.c_void_p), ("pauli", C.c_void_p)]


LADDER_DEPOLARIZING, LADDER_ALPHA, LADDER_BIASED = 0, 1, 2
DISTR_F64, DISTR_U8 = 0, 1

_lib = None


def load():
    """Load libqecmc.so; raises QecmcError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise QecmcError(f"{SO} not found: build it with `python __graft_entry__.py` "
                         "(nvcc, sm_100a); there is no CPU fallback")
    L = C.CDLL(SO)
    L.qecmc_last_error.restype = C.c_char_p
    L.qecmc_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.qecmc_destroy.argtypes = [C.c_void_p]
    L.qecmc_destroy.restype = None
    L.qecmc_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.qecmc_set_table_budget.argtypes = [C.c_void_p, C.c_int64]
    L.qecmc_device_info.argtypes = [C.c_void_p, C.POINTER(DevInfo)]
    L.qecmc_debug_set.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.qecmc_last_plan.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.qecmc_chain_update.argtypes = [C.c_void_p, C.POINTER(ChainCfg), C.c_void_p, C.c_int64, C.c_int64, C.POINTER(Stats)]
    L.qecmc_replay_chain.argtypes = [C.c_void_p, C.POINTER(ChainCfg), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.qecmc_stdc.argtypes = [C.c_void_p, C.POINTER(StdcCfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
    L.qecmc_stdc_dev.argtypes = L.qecmc_stdc.argtypes
    L.qecmc_strc.argtypes = [C.c_void_p, C.POINTER(StdcCfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.POINTER(Stats)]
    L.qecmc_single_temp.argtypes = [C.c_void_p, C.POINTER(StdcCfg), C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(Stats)]
    L.qecmc_ladder_run.argtypes = [C.c_void_p, C.POINTER(LadderCfg), C.POINTER(LadderIO), C.c_int64, C.c_int64,
                                   C.POINTER(Stats)]
    L.qecmc_pteq.argtypes = [C.c_void_p, C.POINTER(PteqCfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.POINTER(Stats)]
    L.qecmc_pteq_shortest.argtypes = [C.c_void_p, C.POINTER(PteqCfg), C.c_void_p, C.c_int64] + [C.c_void_p] * 5 + \
        [C.POINTER(Stats)]
    L.qecmc_pteq_dev.argtypes = [C.c_void_p, C.POINTER(PteqCfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                 C.POINTER(Stats)]
    L.qecmc_ptdc.argtypes = [C.c_void_p, C.POINTER(PtdcCfg), C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(Stats)]
    L.qecmc_ptrc.argtypes = [C.c_void_p, C.POINTER(PtdcCfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.POINTER(Stats)]
    L.qecmc_chain_update_xyz.argtypes = [C.c_void_p, C.POINTER(ChainCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_int64, C.POINTER(Stats)]
    L.qecmc_stdc_general_noise.argtypes = [C.c_void_p, C.POINTER(XyzCfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.POINTER(Stats)]
    L.qecmc_stdc_alpha.argtypes = [C.c_void_p, C.POINTER(AlphaCfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.POINTER(Stats)]
    L.qecmc_generate_errors.argtypes = [C.c_void_p, C.POINTER(NoiseCfg), C.c_int64, C.c_void_p, C.c_void_p]
    L.qecmc_generate_errors_dev.argtypes = L.qecmc_generate_errors.argtypes
    L.qecmc_define_equivalence_class.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
    L.qecmc_define_equivalence_class_dev.argtypes = L.qecmc_define_equivalence_class.argtypes
    L.qecmc_apply_random_logical.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p,
                                             C.c_void_p]
    L.qecmc_apply_random_logical_dev.argtypes = L.qecmc_apply_random_logical.argtypes
    L.qecmc_count_failures.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                       C.c_void_p, C.POINTER(C.c_int64)]
    L.qecmc_count_failures_dev.argtypes = L.qecmc_count_failures.argtypes
    L.qecmc_mwpm_planar.argtypes = [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_int32]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise QecmcError(f"libqecmc error {rc}: {load().qecmc_last_error().decode()}")


def nsites(geom, L):
    return 2 * L * L if geom in (TORIC, PLANAR) else L * L


def neq(geom):
    return 16 if geom == TORIC else 4


def ndraws(geom):
    return 3 if geom in (TORIC, PLANAR) else 5


class Context:
    """One per GPU; not thread-safe (include/qecmc.h)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(load().qecmc_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            load().qecmc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(load().qecmc_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_table_budget(self, nbytes):
        _check(load().qecmc_set_table_budget(self._h, int(nbytes)))

    def debug_set(self, key, value):
        """Test switch between code paths that must agree (include/qecmc.h); value < 0 restores the default."""
        _check(load().qecmc_debug_set(self._h, key.encode(), int(value)))

    def last_plan(self):
        """(wave_capacity, round_chains) of the most recent STDC-family call: syndromes one wave may hold within the table
        budget, chains one round of full-size CTAs over all SMs holds."""
        w, r = C.c_int64(0), C.c_int64(0)
        _check(load().qecmc_last_plan(self._h, C.byref(w), C.byref(r)))
        return int(w.value), int(r.value)

    def device_info(self):
        d = DevInfo()
        _check(load().qecmc_device_info(self._h, C.byref(d)))
        out = {f: getattr(d, f) for f, _ in d._fields_}
        out["name"] = d.name.decode()
        return out

    # ---- plain chains -------------------------------------------------
    def chain_update(self, geom_chain, L, qm, p, iters, seed=0, stream_offset=0, pow_kind=POW_NUMBA):
        """qm: uint8 [chains, n_sites] (C-contiguous), updated in place.  Returns stats dict."""
        _require_u8(qm)
        chains = qm.shape[0]
        cfg = ChainCfg(geom_chain, L, pow_kind, 0, p, seed, stream_offset)
        st = Stats()
        _check(load().qecmc_chain_update(self._h, C.byref(cfg), qm.ctypes.data, chains, iters, C.byref(st)))
        return st.as_dict()

    def replay_chain(self, geom_chain, L, qm0, u, p, pow_kind=POW_NUMBA, want_traj=False):
        """qm0 [chains, n_sites]; u [chains, iters, k+1].  Returns (qm_final, dE, accepted[, traj])."""
        _require_u8(qm0)
        u = np.ascontiguousarray(u, np.float64)
        chains, iters = u.shape[0], u.shape[1]
        assert u.shape[2] == ndraws(geom_chain) + 1 and qm0.shape[0] == chains
        cfg = ChainCfg(geom_chain, L, pow_kind, 0, p, 0, 0)
        out = np.empty_like(qm0)
        dE = np.zeros((chains, iters), np.int8)
        acc = np.zeros((chains, iters), np.uint8)
        traj = np.zeros((chains, iters, qm0.shape[1]), np.uint8) if want_traj else None
        _check(load().qecmc_replay_chain(self._h, C.byref(cfg), qm0.ctypes.data, u.ctypes.data, chains, iters,
                                         out.ctypes.data, dE.ctypes.data, acc.ctypes.data,
                                         traj.ctypes.data if want_traj else None))
        return (out, dE, acc, traj) if want_traj else (out, dE, acc)

    # ---- STDC -----------------------------------------------------------
    def _stdc_cfg(self, geom_code, geom_chain, L, droplets, steps, p_error, p_sampling, iters, per_class, randomize,
                  conv_mult, seed, u_nb, u_np):
        return StdcCfg(geom_code, geom_chain, L, droplets, iters, int(per_class), int(randomize), 0, steps, p_error,
                       p_sampling, conv_mult, seed, u_nb, u_np)

    def stdc(self, geom_code, geom_chain, L, qm, p_error, p_sampling, droplets, steps, iters=5, per_class=False,
             randomize=True, conv_mult=0.0, seed=0, u_nb=None, u_np=None, want_hist=False):
        """Host buffers in, host buffers out.  qm: [S, n_sites] or [S, n_eq, n_sites] uint8.
        Returns (eqdistr [S, n_eq] float64, stats dict[, N_hist [S, n_eq, n_sites+1] uint32])."""
        _require_u8(qm)
        S = qm.shape[0]
        n_eq, n = neq(geom_code), nsites(geom_code, L)
        assert qm.size == S * (n_eq if per_class else 1) * n, "qubit_matrix batch has the wrong shape"
        keep = []
        if u_nb is not None:
            u_nb = np.ascontiguousarray(u_nb, np.float64)
            keep.append(u_nb)
            assert u_nb.size == S * n_eq * droplets * steps * iters * (ndraws(geom_chain) + 1)
        if u_np is not None:
            u_np = np.ascontiguousarray(u_np, np.float64)
            keep.append(u_np)
            assert u_np.size == S * n_eq * droplets * 2 * L * L
        cfg = self._stdc_cfg(geom_code, geom_chain, L, droplets, steps, p_error, p_sampling, iters, per_class, randomize,
                             conv_mult, seed, u_nb.ctypes.data if u_nb is not None else None,
                             u_np.ctypes.data if u_np is not None else None)
        out = np.zeros((S, n_eq), np.float64)
        hist = np.zeros((S, n_eq, n + 1), np.uint32) if want_hist else None
        st = Stats()
        _check(load().qecmc_stdc(self._h, C.byref(cfg), qm.ctypes.data, S, out.ctypes.data,
                                 hist.ctypes.data if want_hist else None, C.byref(st)))
        return (out, st.as_dict(), hist) if want_hist else (out, st.as_dict())

    def _replay_args(self, S, n_eq, geom_chain, L, droplets, steps, iters, u_nb, u_np):
        keep = []
        if u_nb is not None:
            u_nb = np.ascontiguousarray(u_nb, np.float64)
            keep.append(u_nb)
            assert u_nb.size == S * n_eq * droplets * steps * iters * (ndraws(geom_chain) + 1)
        if u_np is not None:
            u_np = np.ascontiguousarray(u_np, np.float64)
            keep.append(u_np)
            assert u_np.size == S * n_eq * droplets * 2 * L * L
        return (u_nb.ctypes.data if u_nb is not None else None, u_np.ctypes.data if u_np is not None else None, keep)

    def strc(self, geom_code, geom_chain, L, qm, p_error, p_sampling, droplets, steps, iters=5, per_class=False,
             randomize=True, conv_mult=0.0, seed=0, u_nb=None, u_np=None, want_hist=False):
        """STRC over a batch.  Returns (eqdistr [S, n_eq], stats[, m_hist [S, n_eq, n_sites+1] uint64,
        short_info [S, n_eq, 4] int32])."""
        _require_u8(qm)
        S = qm.shape[0]
        n_eq, n = neq(geom_code), nsites(geom_code, L)
        assert qm.size == S * (n_eq if per_class else 1) * n, "qubit_matrix batch has the wrong shape"
        a, b, keep = self._replay_args(S, n_eq, geom_chain, L, droplets, steps, iters, u_nb, u_np)
        cfg = self._stdc_cfg(geom_code, geom_chain, L, droplets, steps, p_error, p_sampling, iters, per_class, randomize,
                             conv_mult, seed, a, b)
        out = np.zeros((S, n_eq), np.float64)
        mh = np.zeros((S, n_eq, n + 1), np.uint64) if want_hist else None
        info = np.zeros((S, n_eq, 4), np.int32) if want_hist else None
        st = Stats()
        _check(load().qecmc_strc(self._h, C.byref(cfg), qm.ctypes.data, S, out.ctypes.data,
                                 mh.ctypes.data if want_hist else None, info.ctypes.data if want_hist else None, C.byref(st)))
        return (out, st.as_dict(), mh, info) if want_hist else (out, st.as_dict())

    def single_temp(self, geom_code, geom_chain, L, qm, p, max_iters, iters=5, per_class=False, seed=0, u_nb=None):
        """single_temp over a batch: mean chain length [S, n_eq] of one chain per class."""
        _require_u8(qm)
        S = qm.shape[0]
        n_eq, n = neq(geom_code), nsites(geom_code, L)
        assert qm.size == S * (n_eq if per_class else 1) * n, "qubit_matrix batch has the wrong shape"
        a, _, keep = self._replay_args(S, n_eq, geom_chain, L, 1, max_iters, iters, u_nb, None)
        cfg = self._stdc_cfg(geom_code, geom_chain, L, 1, max_iters, p, p, iters, per_class, False, 0.0, seed, a, None)
        out = np.zeros((S, n_eq), np.float64)
        st = Stats()
        _check(load().qecmc_single_temp(self._h, C.byref(cfg), qm.ctypes.data, S, out.ctypes.data, C.byref(st)))
        return out, st.as_dict()

    # ---- tempering ladders ------------------------------------------------
    @staticmethod
    def _ladder_cfg(geom, L, kind, Nc, iters, bottom, param_b, p_logical, seed, u_nb, u_py, n_ladders):
        keep = []
        a = b = None
        n_nb = n_py = 0
        if u_nb is not None:
            u_nb = np.ascontiguousarray(u_nb, np.float64).reshape(n_ladders, -1)
            u_py = np.ascontiguousarray(u_py, np.float64).reshape(n_ladders, -1)
            keep += [u_nb, u_py]
            a, b, n_nb, n_py = u_nb.ctypes.data, u_py.ctypes.data, u_nb.shape[1], u_py.shape[1]
        return LadderCfg(geom, L, kind, Nc, iters, 0, bottom, param_b, p_logical, seed, a, b, n_nb, n_py), keep

    def ladder_run(self, geom, L, kind, qm0, bottom, Nc, steps, iters=10, param_b=0.0, p_logical=0.0, seed=0, u_nb=None,
                   u_py=None, snapshots=False, resume=None):
        """`steps` Ladder.step(iters) calls on S ladders.  Fresh start: qm0 [S, n_sites].  resume: a dict
        returned by an earlier call (rung_states / flags / tops0 / n_eff_parts are continued in place).
        Returns a dict with rung_states [S, Nc, n_sites], flags [S, Nc], tops0 [S], n_eff [S, Nc],
        n_eff_parts [S, Nc, 2], stats (+ snap_* after every step)."""
        n = nsites(geom, L)
        if resume is None:
            _require_u8(qm0)
            S = qm0.shape[0]
            assert qm0.size == S * n
            out = dict(rung_states=np.zeros((S, Nc, n), np.uint8), flags=np.zeros((S, Nc), np.int32),
                       tops0=np.zeros(S, np.int64), n_eff_parts=np.zeros((S, Nc, 2), np.int32))
        else:
            out = resume
            S = out["rung_states"].shape[0]
            _require_u8(out["rung_states"])
        out["n_eff"] = np.zeros((S, Nc), np.float64)
        cfg, keep = self._ladder_cfg(geom, L, kind, Nc, iters, bottom, param_b, p_logical, seed, u_nb, u_py, S)
        snap = [None, None, None]
        if snapshots:
            snap = [np.zeros((S, steps, Nc, n), np.uint8), np.zeros((S, steps, Nc), np.int32), np.zeros((S, steps), np.int64)]
            out.update(snap_states=snap[0], snap_flags=snap[1], snap_tops0=snap[2])
        io = LadderIO(qm0.ctypes.data if resume is None else None, 0 if resume is None else 1, 0,
                      out["rung_states"].ctypes.data, out["flags"].ctypes.data, out["tops0"].ctypes.data,
                      out["n_eff"].ctypes.data, out["n_eff_parts"].ctypes.data,
                      *[x.ctypes.data if x is not None else None for x in snap])
        st = Stats()
        _check(load().qecmc_ladder_run(self._h, C.byref(cfg), C.byref(io), S, steps, C.byref(st)))
        out["stats"] = st.as_dict()
        return out

    def pteq(self, geom, L, kind, qm, bottom, Nc=None, param_b=0.0, SEQ=2, TOPS=10, tops_burn=2, eps=0.1, steps=1000000,
             iters=10, conv=True, p_logical=0.5, seed=0, u_nb=None, u_py=None):
        """PTEQ / PTEQ_alpha / PTEQ_biased over a batch qm [S, n_sites].  Returns (percent uint8 [S, n_eq], info dict)."""
        _require_u8(qm)
        S, n, n_eq = qm.shape[0], nsites(geom, L), neq(geom)
        assert qm.size == S * n
        Nc = Nc or L
        lc, keep = self._ladder_cfg(geom, L, kind, Nc, iters, bottom, param_b, p_logical, seed, u_nb, u_py, S)
        cfg = PteqCfg(lc, SEQ, TOPS, tops_burn, int(bool(conv)), eps, int(steps))
        pct = np.zeros((S, n_eq), np.uint8)
        counts = np.zeros((S, n_eq), np.int64)
        info = np.zeros((S, 4), np.int64)
        st = Stats()
        _check(load().qecmc_pteq(self._h, C.byref(cfg), qm.ctypes.data, S, pct.ctypes.data, counts.ctypes.data,
                                 info.ctypes.data, C.byref(st)))
        return pct, dict(steps=info[:, 0], since_burn=info[:, 1], tops0=info[:, 2], converged=info[:, 3], counts=counts,
                         stats=st.as_dict())

    def pteq_shortest(self, geom, L, kind, qm, bottom, Nc=None, param_b=0.0, SEQ=2, TOPS=10, tops_burn=2, eps=0.1,
                      steps=1000000, iters=10, conv=True, p_logical=0.5, seed=0, u_nb=None, u_py=None):
        """PTEQ_alpha_with_shortest over a batch.  Returns (percent uint8 [S, n_eq], short_len, short_n, short_unique
        [S, n_eq], info dict)."""
        _require_u8(qm)
        S, n, n_eq = qm.shape[0], nsites(geom, L), neq(geom)
        assert qm.size == S * n
        Nc = Nc or L
        lc, keep = self._ladder_cfg(geom, L, kind, Nc, iters, bottom, param_b, p_logical, seed, u_nb, u_py, S)
        cfg = PteqCfg(lc, SEQ, TOPS, tops_burn, int(bool(conv)), eps, int(steps))
        pct = np.zeros((S, n_eq), np.uint8)
        slen, sn, su = np.zeros((S, n_eq)), np.zeros((S, n_eq), np.int64), np.zeros((S, n_eq), np.int64)
        info = np.zeros((S, 4), np.int64)
        st = Stats()
        _check(load().qecmc_pteq_shortest(self._h, C.byref(cfg), qm.ctypes.data, S, pct.ctypes.data, slen.ctypes.data,
                                          sn.ctypes.data, su.ctypes.data, info.ctypes.data, C.byref(st)))
        return pct, slen, sn, su, dict(steps=info[:, 0], since_burn=info[:, 1], tops0=info[:, 2], converged=info[:, 3],
                                       stats=st.as_dict())

    def pteq_dev(self, geom, L, kind, d_qm_ptr, S, d_out_ptr, bottom, Nc=None, param_b=0.0, SEQ=2, TOPS=10, tops_burn=2,
                 eps=0.1, steps=1000000, iters=10, conv=True, p_logical=0.5, seed=0):
        Nc = Nc or L
        lc, _ = self._ladder_cfg(geom, L, kind, Nc, iters, bottom, param_b, p_logical, seed, None, None, S)
        cfg = PteqCfg(lc, SEQ, TOPS, tops_burn, int(bool(conv)), eps, int(steps))
        info = np.zeros((S, 4), np.int64)
        st = Stats()
        _check(load().qecmc_pteq_dev(self._h, C.byref(cfg), C.c_void_p(d_qm_ptr), S, C.c_void_p(d_out_ptr), info.ctypes.data,
                                     C.byref(st)))
        return st.as_dict(), info

    def ptdc(self, geom, L, qm, p_error, p_sampling, droplets, Nc, steps, iters=10, per_class=False, seed=0, u_nb=None,
             u_py=None, conv_mult=0.0, want_steps=False):
        """PTDC over a batch; `steps` is the per-ladder step count (the reference's steps // Nc); conv_mult != 0 is the
        early stop of PTDC_droplet."""
        _require_u8(qm)
        S, n, n_eq = qm.shape[0], nsites(geom, L), neq(geom)
        assert qm.size == S * (n_eq if per_class else 1) * n
        lc, keep = self._ladder_cfg(geom, L, LADDER_DEPOLARIZING, Nc, iters, p_sampling, 0.0, 0.0, seed, u_nb, u_py,
                                    S * n_eq * droplets)
        done = np.zeros((S, n_eq, droplets), np.int64)
        cfg = PtdcCfg(lc, droplets, int(per_class), int(steps), p_error, float(conv_mult), done.ctypes.data if want_steps else None)
        out = np.zeros((S, n_eq), np.float64)
        st = Stats()
        _check(load().qecmc_ptdc(self._h, C.byref(cfg), qm.ctypes.data, S, out.ctypes.data, C.byref(st)))
        return (out, st.as_dict(), done) if want_steps else (out, st.as_dict())

    def ptrc(self, geom, L, qm, p_error, p_sampling, droplets, Nc, steps, iters=10, per_class=False, seed=0, u_nb=None,
             u_py=None, want_hist=False):
        """PTRC over a batch; `steps` is the per-ladder step count.  Returns (eqdistr float64 [S, n_eq], stats
        [, N_hist, m_hist int64 [S, n_eq, Nc, n_sites+1]])."""
        _require_u8(qm)
        S, n, n_eq = qm.shape[0], nsites(geom, L), neq(geom)
        assert qm.size == S * (n_eq if per_class else 1) * n
        lc, keep = self._ladder_cfg(geom, L, LADDER_DEPOLARIZING, Nc, iters, p_sampling, 0.0, 0.0, seed, u_nb, u_py,
                                    S * n_eq * droplets)
        cfg = PtdcCfg(lc, droplets, int(per_class), int(steps), p_error)
        out = np.zeros((S, n_eq), np.float64)
        Nh = np.zeros((S, n_eq, Nc, n + 1), np.int64) if want_hist else None
        mh = np.zeros((S, n_eq, Nc, n + 1), np.int64) if want_hist else None
        st = Stats()
        _check(load().qecmc_ptrc(self._h, C.byref(cfg), qm.ctypes.data, S, out.ctypes.data,
                                 Nh.ctypes.data if want_hist else None, mh.ctypes.data if want_hist else None, C.byref(st)))
        return (out, st.as_dict(), Nh, mh) if want_hist else (out, st.as_dict())

    def stdc_alpha(self, geom, L, qm, pz_tilde_sampling, alpha, pz_tilde, steps, iters=5, per_class=False, seed=0, u_nb=None,
                   u_py=None):
        """EWD-style STDC_Nall_n_alpha over a batch.  Returns (eqdistr [S, n_eq], distinct [S, n_eq], stats)."""
        _require_u8(qm)
        S, n, n_eq = qm.shape[0], nsites(geom, L), neq(geom)
        assert qm.size == S * (n_eq if per_class else 1) * n
        keep, a, b, n_nb, n_py = [], None, None, 0, 0
        if u_nb is not None:
            u_nb = np.ascontiguousarray(u_nb, np.float64).reshape(S * n_eq, -1)
            u_py = np.ascontiguousarray(u_py, np.float64).reshape(S * n_eq, -1)
            keep += [u_nb, u_py]
            a, b, n_nb, n_py = u_nb.ctypes.data, u_py.ctypes.data, u_nb.shape[1], u_py.shape[1]
        cfg = AlphaCfg(geom, L, iters, int(per_class), int(steps), pz_tilde_sampling, alpha, pz_tilde, seed, a, b, n_nb, n_py)
        out = np.zeros((S, n_eq), np.float64)
        distinct = np.zeros((S, n_eq), np.int64)
        st = Stats()
        _check(load().qecmc_stdc_alpha(self._h, C.byref(cfg), qm.ctypes.data, S, out.ctypes.data, distinct.ctypes.data,
                                       C.byref(st)))
        return out, distinct, st.as_dict()

    def chain_update_xyz(self, geom_chain, L, qm, p_xyz, iters, seed=0, stream_offset=0, u=None):
        """Chain_xyz.update_chain_fast on qm [chains, n_sites] in place; u: replay draws [chains, iters, k+1]."""
        _require_u8(qm)
        chains = qm.shape[0]
        cfg = ChainCfg(geom_chain, L, POW_NUMBA, 0, 0.5, seed, stream_offset)
        pp = (C.c_double * 3)(*[float(x) for x in p_xyz])
        if u is not None:
            u = np.ascontiguousarray(u, np.float64)
            assert u.size == chains * iters * (ndraws(geom_chain) + 1)
        st = Stats()
        _check(load().qecmc_chain_update_xyz(self._h, C.byref(cfg), pp, u.ctypes.data if u is not None else None,
                                             qm.ctypes.data, chains, iters, C.byref(st)))
        return st.as_dict()

    def stdc_general_noise(self, geom_code, geom_chain, L, qm, p_xyz, p_sampling, droplets, steps, iters=5, per_class=False,
                           seed=0, u_nb=None):
        """STDC_general_noise(_shortest) over a batch; p_sampling: float (Chain) or array of 3 (Chain_xyz).
        Returns (eqdistr, eqdistr_shortest [S, n_eq] float64, distinct [S, n_eq], stats)."""
        _require_u8(qm)
        S, n, n_eq = qm.shape[0], nsites(geom_code, L), neq(geom_code)
        assert qm.size == S * (n_eq if per_class else 1) * n
        use_xyz = isinstance(p_sampling, np.ndarray)
        keep, a = [], None
        if u_nb is not None:
            u_nb = np.ascontiguousarray(u_nb, np.float64)
            assert u_nb.size == S * n_eq * droplets * steps * iters * (ndraws(geom_chain) + 1)
            keep.append(u_nb)
            a = u_nb.ctypes.data
        cfg = XyzCfg(geom_code, geom_chain, L, droplets, iters, int(per_class), int(use_xyz), 0, int(steps),
                     (C.c_double * 3)(*[float(x) for x in p_xyz]),
                     (C.c_double * 3)(*([float(x) for x in p_sampling] if use_xyz else [0.0, 0.0, 0.0])),
                     0.0 if use_xyz else float(p_sampling), seed, a)
        out, out_s = np.zeros((S, n_eq)), np.zeros((S, n_eq))
        distinct = np.zeros((S, n_eq), np.int64)
        st = Stats()
        _check(load().qecmc_stdc_general_noise(self._h, C.byref(cfg), qm.ctypes.data, S, out.ctypes.data, out_s.ctypes.data,
                                               distinct.ctypes.data, C.byref(st)))
        return out, out_s, distinct, st.as_dict()

    def stdc_dev(self, geom_code, geom_chain, L, d_qm_ptr, S, d_out_ptr, p_error, p_sampling, droplets, steps, iters=5,
                 per_class=False, randomize=True, seed=0, want_stats=True):
        """Device pointers (e.g. torch tensors' data_ptr()); work is enqueued on the context's stream."""
        cfg = self._stdc_cfg(geom_code, geom_chain, L, droplets, steps, p_error, p_sampling, iters, per_class, randomize,
                             0.0, seed, None, None)
        st = Stats()
        _check(load().qecmc_stdc_dev(self._h, C.byref(cfg), C.c_void_p(d_qm_ptr), S, C.c_void_p(d_out_ptr), None,
                                     C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None


    # ---- the workload loop around the decoders (generate_data.py:53-261) ----
    @staticmethod
    def _noise_cfg(geom, L, p_error, p_xyz, seed, u, pauli):
        toric_form = p_xyz is None
        px, py, pz = (0.0, 0.0, 0.0) if toric_form else [float(x) for x in p_xyz]
        return NoiseCfg(geom, L, int(toric_form), 0, float(p_error or 0.0), px, py, pz, seed, u, pauli)

    def generate_errors(self, geom, L, S, p_error=None, p_xyz=None, seed=0, u=None, pauli=None, want_class=True):
        """generate_random_error on the device: p_error = the toric form (toric_model.py:15-24), p_xyz = (p_x, p_y, p_z)
        (planar_model.py:18-37 and friends).  u / pauli replay the reference's own draws.  -> (qm [S, n_sites], eq_true [S])"""
        n = nsites(geom, L)
        if u is not None:
            u = np.ascontiguousarray(u, dtype=np.float64)
            assert u.size == S * n, "u must hold S * n_sites uniforms"
        if pauli is not None:
            pauli = np.ascontiguousarray(pauli, dtype=np.uint8)
            assert pauli.size == S * n
        cfg = self._noise_cfg(geom, L, p_error, p_xyz, seed, u.ctypes.data if u is not None else None,
                              pauli.ctypes.data if pauli is not None else None)
        qm = np.zeros((S, n), np.uint8)
        cls = np.zeros(S, np.int32)
        _check(load().qecmc_generate_errors(self._h, C.byref(cfg), S, qm.ctypes.data, cls.ctypes.data if want_class else None))
        return qm, (cls if want_class else None)

    def generate_errors_dev(self, geom, L, S, d_qm_ptr, d_eq_true_ptr, p_error=None, p_xyz=None, seed=0):
        cfg = self._noise_cfg(geom, L, p_error, p_xyz, seed, None, None)
        _check(load().qecmc_generate_errors_dev(self._h, C.byref(cfg), S, C.c_void_p(d_qm_ptr),
                                                C.c_void_p(d_eq_true_ptr) if d_eq_true_ptr else None))

    def define_equivalence_class(self, geom, L, qm):
        qm = np.ascontiguousarray(qm)
        _require_u8(qm)
        S = qm.size // nsites(geom, L)
        cls = np.zeros(S, np.int32)
        _check(load().qecmc_define_equivalence_class(self._h, geom, L, qm.ctypes.data, S, cls.ctypes.data))
        return cls

    def define_equivalence_class_dev(self, geom, L, d_qm_ptr, S, d_cls_ptr):
        _check(load().qecmc_define_equivalence_class_dev(self._h, geom, L, C.c_void_p(d_qm_ptr), S, C.c_void_p(d_cls_ptr)))

    def apply_random_logical(self, geom, L, qm, seed=0, u=None):
        """apply_random_logical on every lattice of the batch -> (new qm, ops [S, 2])"""
        qm = np.array(qm, dtype=np.uint8, order="C", copy=True)
        S = qm.size // nsites(geom, L)
        if u is not None:
            u = np.ascontiguousarray(u, dtype=np.float64)
            assert u.shape == (S, 6), "u must be [S, 6] (unused trailing entries are ignored)"
        ops = np.zeros((S, 2), np.int32)
        _check(load().qecmc_apply_random_logical(self._h, geom, L, qm.ctypes.data, S, seed, u.ctypes.data if u is not None else None,
                                                 ops.ctypes.data))
        return qm, ops

    def apply_random_logical_dev(self, geom, L, d_qm_ptr, S, seed=0):
        _check(load().qecmc_apply_random_logical_dev(self._h, geom, L, C.c_void_p(d_qm_ptr), S, seed, None, None))

    def count_failures(self, distr, eq_true, use_argmin=False):
        """np.argmax(distr) != eq_true per syndrome (np.argmin for single_temp) -> (failures, choice [S])"""
        distr = np.ascontiguousarray(distr)
        if distr.dtype == np.uint8:
            dt = DISTR_U8
        else:
            distr = np.ascontiguousarray(distr, dtype=np.float64)
            dt = DISTR_F64
        S, n_eq = distr.shape
        eq_true = np.ascontiguousarray(eq_true, dtype=np.int32)
        choice = np.zeros(S, np.int32)
        fails = C.c_int64(0)
        _check(load().qecmc_count_failures(self._h, distr.ctypes.data, dt, int(use_argmin), n_eq, S, eq_true.ctypes.data,
                                           choice.ctypes.data, C.byref(fails)))
        return int(fails.value), choice

    def count_failures_dev(self, d_distr_ptr, dtype, n_eq, S, d_eq_true_ptr, d_choice_ptr=None, use_argmin=False):
        fails = C.c_int64(0)
        _check(load().qecmc_count_failures_dev(self._h, C.c_void_p(d_distr_ptr), dtype, int(use_argmin), n_eq, S,
                                               C.c_void_p(d_eq_true_ptr), C.c_void_p(d_choice_ptr) if d_choice_ptr else None,
                                               C.byref(fails)))
        return int(fails.value)


def mwpm_planar(L, qm=None, vertex_defects=None, plaquette_defects=None, class_sorted=True, threads=0, unreduced=False):
    """qecmc_mwpm_planar: host-side matching (no device, no context).  qm [S][2][L][L] error chains, or the two defect
    arrays [S][L-1][L] / [S][L][L-1].  Returns (chains, weights): class_sorted -> chains [S][4][2][L][L] in class order,
    weights [S][2][2] = weight of solve_layer(layer, parity); else chains [S][2][L][L] = MWPM.solve(), weights [S][2]."""
    if qm is not None:
        qm = np.ascontiguousarray(qm, dtype=np.uint8).reshape(-1, 2, L, L)
        S = qm.shape[0]
        v = p = None
    else:
        v = np.ascontiguousarray(np.asarray(vertex_defects) != 0, dtype=np.uint8).reshape(-1, L - 1, L)
        p = np.ascontiguousarray(np.asarray(plaquette_defects) != 0, dtype=np.uint8).reshape(-1, L, L - 1)
        S = v.shape[0]
        if p.shape[0] != S:
            raise ValueError("vertex_defects and plaquette_defects disagree on the batch size")
    out = np.zeros((S, 4, 2, L, L) if class_sorted else (S, 2, L, L), np.uint8)
    w = np.zeros((S, 2, 2) if class_sorted else (S, 2), np.int32)
    _check(load().qecmc_mwpm_planar(L, S, qm.ctypes.data if qm is not None else None, v.ctypes.data if v is not None else None,
                                    p.ctypes.data if p is not None else None, int(bool(class_sorted)) | (2 if unreduced else 0), out.ctypes.data,
                                    w.ctypes.data, int(threads)))
    return out, w


def host_planar_class(qm):
    """Planar_code.define_equivalence_class (planar_model.py:379-390) of host lattices [S][2][L][L] -> int [S]."""
    q = np.asarray(qm)
    x = ((q[:, 0, :, 0] == 1) | (q[:, 0, :, 0] == 2)).sum(axis=1) % 2
    z = ((q[:, 0, 0, :] == 3) | (q[:, 0, 0, :] == 2)).sum(axis=1) % 2
    return (x + 2 * z).astype(np.int32)


def _require_u8(a):
    # the reference's njit signatures reject anything but C-contiguous uint8
    if not isinstance(a, np.ndarray) or a.dtype != np.uint8 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("qubit_matrix must be a C-contiguous numpy uint8 array")


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
