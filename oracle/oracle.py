"""ctypes wrapper around oracle/qec_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libqec_oracle.so")

TORIC, PLANAR, ROTATED, XZZX = 0, 1, 2, 3
GEOM = {"toric": TORIC, "planar": PLANAR, "rotated": ROTATED, "xzzx": XZZX}


def build(force=False):
    src = os.path.join(_HERE, "qec_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None
_p = C.c_void_p
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.qo_stream_mt.restype = _p
        L.qo_stream_mt.argtypes = [C.c_uint32]
        L.qo_stream_mt_pyseed.restype = _p
        L.qo_stream_mt_pyseed.argtypes = [C.c_uint64]
        L.qo_stream_replay.restype = _p
        L.qo_stream_replay.argtypes = [_p, C.c_int64]
        L.qo_stream_record.argtypes = [_p, _p, C.c_int64]
        L.qo_stream_drawn.restype = C.c_int64
        L.qo_stream_drawn.argtypes = [_p]
        L.qo_stream_next.restype = C.c_double
        L.qo_stream_next.argtypes = [_p]
        L.qo_stream_free.argtypes = [_p]
        L.qo_stream_philox.restype = _p
        L.qo_stream_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.qo_stream_ladder_native.restype = _p
        L.qo_stream_ladder_native.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.qo_stream_align.argtypes = [_p]
        L.qo_stream_next_bit.argtypes = [_p]
        L.qo_philox4x32_10.argtypes = [_p, _p, _p]
        L.qo_nstab.argtypes = [C.c_int, C.c_int]
        L.qo_stabilizer_by_index.argtypes = [C.c_int, C.c_int, C.c_int, _p, _p, _p]
        L.qo_draw_stabilizer.argtypes = [C.c_int, C.c_int, _p, _p, _p, _p]
        L.qo_numba_pow.restype = C.c_double
        L.qo_numba_pow.argtypes = [C.c_double, C.c_int64]
        L.qo_apply_stabilizer.argtypes = [C.c_int, C.c_int, _u8p, C.c_int, C.c_int, C.c_int]
        L.qo_apply_logical.argtypes = [C.c_int, C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.qo_apply_random_logical.argtypes = [C.c_int, C.c_int, _u8p, _p]
        L.qo_class.argtypes = [C.c_int, C.c_int, _u8p]
        L.qo_to_class.argtypes = [C.c_int, C.c_int, _u8p, C.c_int]
        L.qo_rain.argtypes = [C.c_int, C.c_int, _u8p, _p, C.c_double]
        L.qo_update_chain_fast.argtypes = [C.c_int, C.c_int, _u8p, C.c_double, C.c_int64, _p, _p, _p]
        L.qo_update_chain.argtypes = [C.c_int, C.c_int, _u8p, C.c_double, C.c_double, C.c_int64, _p, _p]
        L.qo_update_chain_weighted.argtypes = [C.c_int, C.c_int, C.c_int, _u8p, C.c_double, C.c_double,
                                               C.c_double, C.c_int64, _p, _p, _p]
        L.qo_ladder_step.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _f64p, _f64p, C.c_double,
                                     C.c_double, _p, _p, _p, C.c_int64, _p, _p]
        L.qo_stdc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_double, C.c_double, C.c_int,
                              C.c_int64, C.c_int64, C.c_int, C.c_double, _p, _p, _f64p, _p, _p]
        L.qo_strc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_double, C.c_double, C.c_int,
                              C.c_int64, C.c_int64, C.c_int, C.c_double, _p, _p, _f64p, _p, _p]
        L.qo_single_temp.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_double, C.c_int64,
                                     C.c_int64, _p, _f64p]
        L.qo_stdc_alpha.argtypes = [C.c_int, C.c_int, C.c_int, _u8p, C.c_double, C.c_double, C.c_double,
                                    C.c_int64, C.c_int64, _p, _p, _f64p, _p]
        L.qo_pteq.restype = C.c_int64
        L.qo_pteq.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _f64p, _f64p, C.c_double, C.c_double,
                              C.c_int, C.c_int, C.c_int, C.c_double, C.c_int64, C.c_int64, C.c_int, _p, _p,
                              _i64p, _p, _p, _u8p]
        L.qo_set_fast_windows.argtypes = [C.c_int]
        L.qo_pteq_ex.restype = C.c_int64
        L.qo_pteq_ex.argtypes = L.qo_pteq.argtypes + [_f64p, _i64p, _i64p]
        L.qo_update_chain_fast_xyz.argtypes = [C.c_int, C.c_int, _u8p, _f64p, C.c_int64, _p]
        L.qo_stdc_general_noise.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _f64p, C.c_int, _f64p, C.c_double,
                                            C.c_int, C.c_int64, C.c_int64, _p, _f64p, _f64p, _i64p]
        L.qo_ptxc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_double, C.c_double, C.c_int,
                              C.c_int64, C.c_int64, _p, _p, _f64p, _p, _p]
        L.qo_stdc_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _u8p, C.c_double, C.c_double,
                                    C.c_int, C.c_int64, C.c_int64, C.c_uint32, C.c_int64, _f64p]
        L.qo_stdc_class.restype = C.c_double
        L.qo_stdc_class.argtypes = [C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int64,
                                    C.c_int64, C.c_uint32, C.c_int64]
        _lib = L
    return _lib


def nsites(geom, L):
    return 2 * L * L if geom in (TORIC, PLANAR) else L * L


def neq(geom):
    return 16 if geom == TORIC else 4


class Stream:
    """One of the reference's three MT19937 streams, or a replay of explicit uniforms."""

    def __init__(self, handle, keep=None):
        self.h = handle
        self._keep = keep
        self._rec = None

    @classmethod
    def mt(cls, seed):  # numba random.seed / np.random.seed
        return cls(lib().qo_stream_mt(seed & 0xFFFFFFFF))

    @classmethod
    def py(cls, seed):  # CPython random.seed(int)
        return cls(lib().qo_stream_mt_pyseed(seed))

    @classmethod
    def philox(cls, key, stream_id, tag=0, call0=0):
        """Native stream of the product's kernels: Philox4x32-10 words of calls call0, call0 + 1, ... with
        counter (call, tag, id lo, id hi) and the 64-bit key (oracle/qec_oracle.c, "Native draws")."""
        return cls(lib().qo_stream_philox(key & 0xFFFFFFFFFFFFFFFF, stream_id, tag & 0xFFFFFFFF, call0 & 0xFFFFFFFF))

    @classmethod
    def ladder_native(cls, key, ladder_id, step0=0):
        """Native words of one tempering ladder, addressed by (Ladder.step, purpose, rung, iteration): pass it as BOTH the
        nb and the py stream of Ladder.step / pteq / ptxc / stdc_alpha (oracle/qec_oracle.c, ladder_step_native)."""
        return cls(lib().qo_stream_ladder_native(key & 0xFFFFFFFFFFFFFFFF, ladder_id, step0 & 0xFFFFFFFF))

    @classmethod
    def replay(cls, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        return cls(lib().qo_stream_replay(u.ctypes.data, u.size), keep=u)

    def record(self, cap):
        self._rec = np.zeros(cap, dtype=np.float64)
        lib().qo_stream_record(self.h, self._rec.ctypes.data, cap)
        return self

    def recorded(self):
        return self._rec[: min(self.drawn, self._rec.size)].copy()

    @property
    def drawn(self):
        return lib().qo_stream_drawn(self.h)

    def next(self):
        return lib().qo_stream_next(self.h)

    def next_bit(self):
        return lib().qo_stream_next_bit(self.h)

    def align(self):
        lib().qo_stream_align(self.h)
        return self

    def __del__(self):
        try:
            lib().qo_stream_free(self.h)
        except Exception:
            pass


def philox4x32_10(ctr, key):
    """One Philox4x32-10 call of the C oracle: ctr (4 words), key (2 words) -> 4 words."""
    c = np.ascontiguousarray(ctr, np.uint32)
    k = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().qo_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def nstab(geom, L):
    return lib().qo_nstab(geom, L)


def stabilizer_by_index(geom, L, idx):
    """(row, col, op) of stabilizer idx in the canonical numbering the native proposals use."""
    r, c_, o = C.c_int(), C.c_int(), C.c_int()
    lib().qo_stabilizer_by_index(geom, L, idx, C.byref(r), C.byref(c_), C.byref(o))
    return r.value, c_.value, o.value


def draw_stabilizer(geom, L, nb):
    r, c_, o = C.c_int(), C.c_int(), C.c_int()
    lib().qo_draw_stabilizer(geom, L, nb.h, C.byref(r), C.byref(c_), C.byref(o))
    return r.value, c_.value, o.value


def _flat(qm):
    a = np.ascontiguousarray(qm, dtype=np.uint8)
    return a.reshape(-1)


def numba_pow(a, b):
    return lib().qo_numba_pow(float(a), int(b))


def apply_stabilizer(geom, L, qm, row, col, op):
    out = _flat(qm).copy()
    d = lib().qo_apply_stabilizer(geom, L, out, row, col, op)
    return out.reshape(np.shape(qm)), d


def apply_logical(geom, L, qm, op, layer=0, X_pos=0, Z_pos=0):
    out = _flat(qm).copy()
    d = lib().qo_apply_logical(geom, L, out, op, layer, X_pos, Z_pos)
    return out.reshape(np.shape(qm)), d


def apply_random_logical(geom, L, qm, nb):
    out = _flat(qm).copy()
    d = lib().qo_apply_random_logical(geom, L, out, nb.h)
    return out.reshape(np.shape(qm)), d


def generate_errors(geom, L, u, p_error=None, p_xyz=None, pauli=None):
    """generate_random_error from explicit uniforms (toric form when p_xyz is None)."""
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
    out = np.zeros(nsites(geom, L), np.uint8)
    toric_form = p_xyz is None
    px, py, pz = (0.0, 0.0, 0.0) if toric_form else [float(x) for x in p_xyz]
    pa = np.ascontiguousarray(pauli, dtype=np.uint8).reshape(-1) if pauli is not None else np.zeros(1, np.uint8)
    f = lib().qo_generate_errors
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    f.restype = None
    f(geom, L, int(toric_form), float(p_error or 0.0), px, py, pz, u.ctypes.data, pa.ctypes.data, out.ctypes.data)
    return out


def count_failures(distr, eq_true, use_argmin=False):
    distr = np.ascontiguousarray(distr, dtype=np.float64)
    eq_true = np.ascontiguousarray(eq_true, dtype=np.int32)
    S, n_eq = distr.shape
    choice = np.zeros(S, np.int32)
    f = lib().qo_count_failures
    f.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    f.restype = C.c_int64
    return int(f(distr.ctypes.data, n_eq, S, int(use_argmin), eq_true.ctypes.data, choice.ctypes.data)), choice


def eq_class(geom, L, qm):
    return lib().qo_class(geom, L, _flat(qm))


def to_class(geom, L, qm, eq):
    out = _flat(qm).copy()
    lib().qo_to_class(geom, L, out, eq)
    return out.reshape(np.shape(qm))


def all_classes(geom, L, qm):
    return np.stack([_flat(to_class(geom, L, qm, e)) for e in range(neq(geom))])


def rain(geom, L, qm, np_stream, p=0.5):
    out = _flat(qm).copy()
    lib().qo_rain(geom, L, out, np_stream.h, p)
    return out.reshape(np.shape(qm))


def update_chain_fast(geom, L, qm, factor, iters, nb, trace=False):
    out = _flat(qm).copy()
    if trace:
        dE = np.zeros(iters, np.int8)
        acc = np.zeros(iters, np.uint8)
        lib().qo_update_chain_fast(geom, L, out, factor, iters, nb.h, dE.ctypes.data, acc.ctypes.data)
        return out.reshape(np.shape(qm)), dE, acc
    lib().qo_update_chain_fast(geom, L, out, factor, iters, nb.h, None, None)
    return out.reshape(np.shape(qm))


def update_chain(geom, L, qm, p, p_logical, iters, nb, py):
    out = _flat(qm).copy()
    lib().qo_update_chain(geom, L, out, p, p_logical, iters, nb.h, py.h)
    return out.reshape(np.shape(qm))


def update_chain_weighted(kind, geom, L, qm, a, b, p_logical, iters, nb, py, n_eff=0.0):
    out = _flat(qm).copy()
    ne = C.c_double(n_eff)
    lib().qo_update_chain_weighted(kind, geom, L, out, a, b, p_logical, iters, nb.h, py.h, C.addressof(ne))
    return out.reshape(np.shape(qm)), ne.value


class Ladder:
    """kind 0 depolarizing, 1 alpha, 2 biased.  Mirrors Ladder/Ladder_alpha/Ladder_biased."""

    def __init__(self, kind, geom, L, qm0, bottom, Nc, p_logical=0.0, param_b=0.0):
        self.kind, self.geom, self.L, self.Nc = kind, geom, L, Nc
        self.p_logical, self.param_b = p_logical, param_b
        if kind == 0:
            top = 0.75
        elif kind == 1:
            top = 1.0
        else:
            top = (param_b + 1) / (2 * param_b + 1)
        lad = np.linspace(bottom, top, Nc)
        self.ladder = np.ascontiguousarray(lad)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.diff = np.ascontiguousarray((lad[:-1] * (1 - lad[1:])) / (lad[1:] * (1 - lad[:-1])))
        q = _flat(qm0)
        self.qm = np.ascontiguousarray(np.tile(q, (Nc, 1)))
        self.flags = np.zeros(Nc, np.int32)
        self.flags[-1] = 1
        nx, ny, nz = [(q == k).sum() for k in (1, 2, 3)]
        self.n_eff = np.full(Nc, nz + param_b * (nx + ny), np.float64)
        self.tops0 = C.c_int64(0)

    def step(self, iters, nb, py):
        lib().qo_ladder_step(self.kind, self.geom, self.L, self.Nc, self.qm.reshape(-1), self.ladder,
                             self.diff if self.diff.size else np.zeros(1), self.param_b, self.p_logical,
                             self.flags.ctypes.data, self.n_eff.ctypes.data, C.addressof(self.tops0), iters,
                             nb.h, py.h)


def _stream_array(streams):
    arr = (C.c_void_p * len(streams))(*[s.h for s in streams])
    return arr


def stdc(geom_code, geom_chain, L, qm_classes, p_error, p_sampling, droplets, steps, nb, np_, iters=5,
         randomize=True, conv_mult=0.0, want_hist=False):
    """qm_classes: [n_eq][n_sites]; nb/np_: lists of n_eq*droplets Streams (may alias)."""
    n_eq = len(qm_classes)
    n = nsites(geom_code, L)
    q = np.ascontiguousarray(np.asarray(qm_classes, np.uint8).reshape(n_eq, n)).reshape(-1)
    out = np.zeros(n_eq)
    distinct = np.zeros(n_eq, np.int64)
    hist = np.zeros((n_eq, n + 1), np.int64)
    a, b = _stream_array(nb), _stream_array(np_)
    lib().qo_stdc(geom_code, geom_chain, L, n_eq, q, p_error, p_sampling, droplets, steps, iters,
                  int(randomize), conv_mult, C.addressof(a), C.addressof(b), out, distinct.ctypes.data,
                  hist.ctypes.data)
    return (out, distinct, hist) if want_hist else out


def strc(geom_code, geom_chain, L, qm_classes, p_error, p_sampling, droplets, steps, nb, np_, iters=5,
         randomize=True, conv_mult=0.0, want_hist=False):
    n_eq = len(qm_classes)
    n = nsites(geom_code, L)
    q = np.ascontiguousarray(np.asarray(qm_classes, np.uint8).reshape(n_eq, n)).reshape(-1)
    out = np.zeros(n_eq)
    mh = np.zeros((n_eq, n + 1), np.int64)
    info = np.zeros((n_eq, 4), np.int64)
    a, b = _stream_array(nb), _stream_array(np_)
    lib().qo_strc(geom_code, geom_chain, L, n_eq, q, p_error, p_sampling, droplets, steps, iters,
                  int(randomize), conv_mult, C.addressof(a), C.addressof(b), out, mh.ctypes.data,
                  info.ctypes.data)
    return (out, mh, info) if want_hist else out


def single_temp(geom_code, geom_chain, L, qm_classes, p, max_iters, nb, iters=5):
    n_eq = len(qm_classes)
    n = nsites(geom_code, L)
    q = np.ascontiguousarray(np.asarray(qm_classes, np.uint8).reshape(n_eq, n)).reshape(-1)
    out = np.zeros(n_eq)
    a = _stream_array(nb)
    lib().qo_single_temp(geom_code, geom_chain, L, n_eq, q, p, max_iters, iters, C.addressof(a), out)
    return out


def stdc_alpha(geom, L, qm_classes, pz_tilde_sampling, alpha, pz_tilde, steps, nb, py, iters=5):
    n_eq = len(qm_classes)
    n = nsites(geom, L)
    q = np.ascontiguousarray(np.asarray(qm_classes, np.uint8).reshape(n_eq, n)).reshape(-1)
    out = np.zeros(n_eq)
    distinct = np.zeros(n_eq, np.int64)
    lib().qo_stdc_alpha(geom, L, n_eq, q, pz_tilde_sampling, alpha, pz_tilde, steps, iters, nb.h, py.h, out,
                        distinct.ctypes.data)
    return out, distinct


def pteq(kind, geom, L, qm0, bottom, nb, py, Nc=None, param_b=0.0, SEQ=2, TOPS=10, tops_burn=2, eps=0.1,
         steps=50000000, iters=10, conv=True, p_logical=0.5):
    Nc = Nc or L
    lad = Ladder(kind, geom, L, qm0, bottom, Nc, p_logical, param_b)
    counts = np.zeros(neq(geom), np.int64)
    sb, tops = C.c_int64(0), C.c_int64(0)
    pct = np.zeros(neq(geom), np.uint8)
    used = lib().qo_pteq(kind, geom, L, Nc, _flat(qm0), lad.ladder, lad.diff if lad.diff.size else np.zeros(1),
                         param_b, p_logical, SEQ, TOPS, tops_burn, eps, steps, iters, int(conv), nb.h, py.h,
                         counts, C.addressof(sb), C.addressof(tops), pct)
    return pct, dict(steps=used, since_burn=sb.value, tops0=tops.value, counts=counts)


def pteq_with_shortest(kind, geom, L, qm0, bottom, nb, py, Nc=None, param_b=0.0, SEQ=2, TOPS=10, tops_burn=2, eps=0.1,
                       steps=50000000, iters=10, conv=True, p_logical=0.5):
    """PTEQ_alpha_with_shortest (decoders_biasednoise.py:93-172): returns (percent uint8, eqdistr from the distinct
    shortest chains, shortest_n percent, info)."""
    Nc = Nc or L
    lad = Ladder(kind, geom, L, qm0, bottom, Nc, p_logical, param_b)
    n_eq = neq(geom)
    counts = np.zeros(n_eq, np.int64)
    sb, tops = C.c_int64(0), C.c_int64(0)
    pct = np.zeros(n_eq, np.uint8)
    slen, sn, su = np.zeros(n_eq), np.zeros(n_eq, np.int64), np.zeros(n_eq, np.int64)
    used = lib().qo_pteq_ex(kind, geom, L, Nc, _flat(qm0), lad.ladder, lad.diff if lad.diff.size else np.zeros(1),
                            param_b, p_logical, SEQ, TOPS, tops_burn, eps, steps, iters, int(conv), nb.h, py.h,
                            counts, C.addressof(sb), C.addressof(tops), pct, slen, sn, su)
    beta = -np.log(bottom)
    z = su * np.exp(-beta * slen)
    with np.errstate(invalid="ignore", divide="ignore"):
        return pct, z / z.sum() * 100, sn / sn.sum() * 100, dict(steps=used, since_burn=sb.value, tops0=tops.value,
                                                                 short_len=slen, short_n=sn, short_unique=su)


def update_chain_fast_xyz(geom, L, qm, factors, iters, nb):
    out = _flat(qm).copy()
    lib().qo_update_chain_fast_xyz(geom, L, out, np.ascontiguousarray(factors, np.float64), iters, nb.h)
    return out.reshape(np.shape(qm))


def stdc_general_noise(geom_code, geom_chain, L, qm_classes, p_xyz, p_sampling, droplets, steps, nb, iters=5):
    """STDC_general_noise_shortest (decoders.py:435-508); p_sampling: float (Chain) or array of 3 (Chain_xyz).
    Returns (eqdistr, eqdistr_shortest, distinct)."""
    n_eq = len(qm_classes)
    n = nsites(geom_code, L)
    q = np.ascontiguousarray(np.asarray(qm_classes, np.uint8).reshape(n_eq, n)).reshape(-1)
    use_xyz = isinstance(p_sampling, np.ndarray)
    ps_xyz = np.ascontiguousarray(p_sampling if use_xyz else np.zeros(3), np.float64)
    out, out_s, distinct = np.zeros(n_eq), np.zeros(n_eq), np.zeros(n_eq, np.int64)
    a = _stream_array(nb)
    lib().qo_stdc_general_noise(geom_code, geom_chain, L, n_eq, q, np.ascontiguousarray(p_xyz, np.float64), int(use_xyz),
                                ps_xyz, 0.0 if use_xyz else float(p_sampling), droplets, steps, iters, C.addressof(a),
                                out, out_s, distinct)
    return out, out_s, distinct


def ptxc(mode, geom, L, qm_classes, p_error, p_sampling, droplets, Nc, steps, nb, py, iters=10, want_hist=False):
    """mode 0: PTDC, mode 1: PTRC (decoders.py:138-233, 584-742); `steps` = per-ladder step count.
    Returns class distribution in percent (float64)[, N_hist, m_hist [n_eq, Nc, n_sites+1] for PTRC]."""
    n_eq = len(qm_classes)
    n = nsites(geom, L)
    q = np.ascontiguousarray(np.asarray(qm_classes, np.uint8).reshape(n_eq, n)).reshape(-1)
    out = np.zeros(n_eq)
    Nh = np.zeros((n_eq, Nc, n + 1), np.int64)
    mh = np.zeros((n_eq, Nc, n + 1), np.int64)
    a, b = _stream_array(nb), _stream_array(py)
    lib().qo_ptxc(mode, geom, L, n_eq, Nc, q, p_error, p_sampling, droplets, steps, iters, C.addressof(a), C.addressof(b),
                  out, Nh.ctypes.data, mh.ctypes.data)
    return (out, Nh, mh) if want_hist else out


def ptdc_conv(geom, L, qm_classes, p_error, p_sampling, droplets, Nc, steps, conv_mult, nb, py, iters=10):
    """PTDC with the conv_mult early stop (decoders.py:138-233); `steps` = per-ladder step count.
    -> (class distribution in percent, steps done [n_eq, droplets], N_hist [n_eq, n_sites + 1])"""
    n_eq = len(qm_classes)
    n = nsites(geom, L)
    q = np.ascontiguousarray(np.asarray(qm_classes, np.uint8).reshape(n_eq, n)).reshape(-1)
    out = np.zeros(n_eq)
    done = np.zeros((n_eq, droplets), np.int64)
    Nh = np.zeros((n_eq, n + 1), np.int64)
    a, b = _stream_array(nb), _stream_array(py)
    f = lib().qo_ptdc_conv
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_double, C.c_double, C.c_int, C.c_int64, C.c_int64, C.c_double,
                  _p, _p, _f64p, _p, _p]
    f.restype = None
    f(geom, L, n_eq, Nc, q, p_error, p_sampling, droplets, steps, iters, float(conv_mult), C.addressof(a), C.addressof(b), out,
      done.ctypes.data, Nh.ctypes.data)
    return out, done, Nh


def stdc_batch(geom_code, geom_chain, L, qm, p_error, p_sampling, droplets, steps, seed, iters=5, threads=1):
    """CPU baseline: STDC over many syndromes on `threads` host threads (ctypes drops the GIL)."""
    qm = np.ascontiguousarray(qm, np.uint8)
    S = qm.shape[0]
    n = nsites(geom_code, L)
    q = qm.reshape(S, n)
    out = np.zeros((S, neq(geom_code)))
    L_ = lib()

    def work(lo, hi):
        if hi > lo:
            L_.qo_stdc_batch(geom_code, geom_chain, L, hi - lo, q[lo:hi].reshape(-1), p_error, p_sampling,
                             droplets, steps, iters, seed, lo, out[lo:hi].reshape(-1))

    if threads > S:
        # fewer syndromes than threads: the unit of work is one (syndrome, class) -- same streams, same results
        Z = np.zeros((S, neq(geom_code)))
        jobs = [(s, e) for s in range(S) for e in range(neq(geom_code))]
        lock = threading.Lock()

        def class_runner():
            while True:
                with lock:
                    if not jobs:
                        return
                    s, e = jobs.pop()
                Z[s, e] = L_.qo_stdc_class(geom_code, geom_chain, L, q[s], e, p_error, p_sampling, droplets, steps, iters, seed, s)

        ts = [threading.Thread(target=class_runner) for _ in range(min(threads, len(jobs)))]
        [t.start() for t in ts]
        [t.join() for t in ts]
        return Z / Z.sum(1, keepdims=True) * 100
    threads = max(1, min(threads, S))
    if threads == 1:
        work(0, S)
    else:
        # finer chunks than threads: per-syndrome cost varies with the number of distinct chains
        bounds = np.linspace(0, S, min(S, threads * 4) + 1).astype(int)
        jobs = list(zip(bounds[:-1], bounds[1:]))
        lock = threading.Lock()

        def runner():
            while True:
                with lock:
                    if not jobs:
                        return
                    lo, hi = jobs.pop()
                work(lo, hi)

        ts = [threading.Thread(target=runner) for _ in range(threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    return out
