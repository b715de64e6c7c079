"""ctypes binding of ``libbq_b200.so`` (C-ABI declared in ``include/bq_capi.h``).

This is the only way the Python host code reaches the GPU.  There is no fallback: if the shared library is
missing, or no CUDA device is visible, the calls raise.
"""
import ctypes
import os

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BQ_LIB", os.path.join(_HERE, "libbq_b200.so"))  # BQ_LIB: tuning builds only

BQ_OK = 0
BQ_ERR_INVALID, BQ_ERR_CUDA, BQ_ERR_UNSUPPORTED, BQ_ERR_NO_DEVICE, BQ_ERR_STATE = -1, -2, -3, -4, -5
BQ_MODE_SLICE, BQ_MODE_MH = 0, 1
BQ_UPD_NONE, BQ_UPD_BAYES, BQ_UPD_STEM = 0, 1, 2
BQ_FLAG_TMAX_PER_SWEEP, BQ_FLAG_NO_FINAL, BQ_FLAG_TRACE, BQ_FLAG_TIGHT, BQ_FLAG_REVERSE = 1, 2, 4, 8, 16

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_f64p = ctypes.POINTER(ctypes.c_double)


class BqModel(ctypes.Structure):
    _fields_ = [("n_queues", ctypes.c_int32), ("n_params", ctypes.c_int32), ("q_type", c_i32p),
                ("q_k", c_i32p), ("q_dist", c_i32p), ("q_param_off", c_i32p),
                ("n_slots", ctypes.c_int32), ("q_slot_off", c_i32p), ("slot_dist", c_i32p)]


class BqArrivals(ctypes.Structure):
    _fields_ = [("n_events", ctypes.c_int32), ("qid", c_i32p), ("tid", c_i32p), ("obs_a", c_u8p),
                ("obs_d", c_u8p), ("prev_byt", c_i32p), ("next_byt", c_i32p), ("prev_byq", c_i32p),
                ("next_byq", c_i32p), ("d", c_f64p)]


class BqStats(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int64) for k in ("nevents", "nreject", "num_diff", "diff_len", "slice_calls",
                                              "slice_evals", "n_sweeps", "n_launches", "param_evals", "n_stuck",
                                              "n_redone")]


class BqError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libbq_b200: %s (status %d)" % (msg, code))
        self.code = code


# every symbol include/bq_capi.h declares: name -> (restype, argtypes)
_H = ctypes.c_void_p
SYMBOLS = {
    "bq_last_error": (ctypes.c_char_p, []),
    "bq_version": (ctypes.c_int, []),
    "bq_device_count": (ctypes.c_int, []),
    "bq_create": (ctypes.c_int, [ctypes.POINTER(BqModel), ctypes.POINTER(BqArrivals), c_f64p, ctypes.c_int32,
                                 ctypes.c_int64, ctypes.c_uint64, ctypes.c_int32, ctypes.POINTER(_H)]),
    "bq_destroy": (ctypes.c_int, [_H]),
    "bq_gibbs_kernel": (ctypes.c_int, [_H, c_f64p, ctypes.c_int32, ctypes.c_int32, c_f64p]),
    "bq_set_params": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_int32, c_f64p]),
    "bq_get_params": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_int32, c_f64p]),
    "bq_set_state": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_int32, c_f64p]),
    "bq_broadcast_state": (ctypes.c_int, [_H, c_f64p]),
    "bq_get_state": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_int32, c_f64p, c_f64p, c_f64p, c_i32p, c_i32p, c_i32p]),
    "bq_set_aux": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_int32, c_u8p]),
    "bq_get_aux": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_int32, c_u8p]),
    "bq_state_dev": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64)]),
    "bq_sweep": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32]),
    "bq_synchronize": (ctypes.c_int, [_H]),
    "bq_init_chains": (ctypes.c_int, [_H, ctypes.c_double]),
    "bq_update_params": (ctypes.c_int, [_H, ctypes.c_int32]),
    "bq_update_params_from": (ctypes.c_int, [_H, ctypes.c_int32, c_f64p, ctypes.c_int32]),
    "bq_logcond": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_int32, c_f64p, ctypes.c_int32, c_f64p]),
    "bq_mh_proposal": (ctypes.c_int, [_H, ctypes.c_int32, ctypes.c_int32, c_f64p, ctypes.c_int32, c_f64p]),
    "bq_log_prob": (ctypes.c_int, [_H, c_f64p]),
    "bq_suffstats": (ctypes.c_int, [_H, c_f64p]),
    "bq_stats": (ctypes.c_int, [_H, ctypes.POINTER(BqStats)]),
    "bq_reset_stats": (ctypes.c_int, [_H]),
    "bq_trace_len": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_int32)]),
    "bq_get_trace": (ctypes.c_int, [_H, c_f64p, c_f64p, c_f64p]),
    "bq_clear_trace": (ctypes.c_int, [_H]),
    "bq_trace_dev": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                    ctypes.POINTER(ctypes.c_void_p)]),
    "bq_get_schedule": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), c_i32p, c_i32p]),
    "bq_get_sequential_order": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_int32), c_i32p]),
    "bq_get_mh_order": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_int32), c_i32p]),
    "bq_rng_uniforms": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.c_uint32, ctypes.c_int32, c_f64p]),
    "bq_rng_slice_uniforms": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.c_uint32, ctypes.c_int32, c_f64p]),
    "bq_philox_block": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), ctypes.c_int32,
                                       ctypes.POINTER(ctypes.c_uint32)]),
    "bq_kernel_info": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                      ctypes.POINTER(ctypes.c_int32)]),
    "bq_cluster_size": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_int32)]),
    "bq_set_relaxed": (ctypes.c_int, [_H, ctypes.c_int32]),
    "bq_get_sweep_counter": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_uint32)]),
    "bq_set_sweep_counter": (ctypes.c_int, [_H, ctypes.c_uint32]),
    "bq_last_sweep_ms": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_float)]),
    "bq_stream": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_void_p)]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built -- there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `make` or `python -c 'import __graft_entry__ as g; "
                              "g.build()'`.  bayes_qnet_b200 has no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != BQ_OK:
        raise BqError(rc, lib().bq_last_error().decode("utf-8", "replace"))


def device_count():
    n = lib().bq_device_count()
    return max(n, 0)


def as_i32(x):
    return numpy.ascontiguousarray(x, dtype=numpy.int32)


def as_f64(x):
    return numpy.ascontiguousarray(x, dtype=numpy.float64)


def ptr(arr, typ):
    return arr.ctypes.data_as(typ) if arr is not None else None


def rng_uniforms(seed, chain, sweep, ident, tag, n):
    out = numpy.empty(n, dtype=numpy.float64)
    check(lib().bq_rng_uniforms(seed, chain, sweep, ident, tag, n, ptr(out, c_f64p)))
    return out


def rng_slice_uniforms(seed, chain, sweep, ident, tag, n):
    """Uniforms of one finite-bound slice update in consumption order (level, candidate 1, 32-bit candidates 2..)."""
    out = numpy.empty(n, dtype=numpy.float64)
    check(lib().bq_rng_slice_uniforms(seed, chain, sweep, ident, tag, n, ptr(out, c_f64p)))
    return out


def philox_block(ctr, key, rounds):
    """One raw Philox4x32 block (known-answer hook)."""
    c = (ctypes.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (ctypes.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (ctypes.c_uint32 * 4)()
    check(lib().bq_philox_block(c, k, int(rounds), o))
    return [int(x) for x in o]
