"""ctypes binding of the C ABI in include/fav_b200.h (libfav_b200.so).

There is no CPU fallback: if the shared library is missing or no sm_100 device is present,
every compute entry point raises.  (The reference binds nothing -- it constructs Python
objects directly, platform/backend/main.py:110-118; INTEGRATION.md shows where this binding
slots in.)"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libfav_b200.so")

_lib = None

c_void_p, c_int, c_float, c_size_t, c_u64, c_uint = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_uint64, C.c_uint
_f3 = C.c_float * 3

SIGNATURES = {
    "fav_abi_version": (c_int, []),
    "fav_last_error": (C.c_char_p, []),
    "fav_init": (c_int, [c_int, C.POINTER(c_void_p)]),
    "fav_reset": (c_int, [c_void_p]),
    "fav_destroy": (c_int, [c_void_p]),
    "fav_corrupt_normalize": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_u64, c_u64,
                                      C.POINTER(c_float), C.POINTER(c_float), c_uint, c_void_p]),
    "fav_corrupt_normalize_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                         C.POINTER(c_float), c_int, C.POINTER(C.c_int32), c_int, c_void_p, c_size_t,
                                         c_void_p, c_size_t, c_u64, c_u64, C.POINTER(c_float), C.POINTER(c_float),
                                         c_uint, c_void_p]),
    "fav_corrupt_params": (c_int, [c_int, c_int, c_int, c_int, c_uint, C.POINTER(c_float), C.POINTER(c_int),
                                   C.POINTER(C.c_int32), C.POINTER(c_int), c_void_p, C.POINTER(c_size_t)]),
    "fav_corruption_constants": (c_int, [c_int, c_int, c_int, C.POINTER(C.c_double), c_int]),
    "fav_corrupt_scratch_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "fav_load_weights": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int]),
    "fav_reserve": (c_int, [c_void_p, c_int, c_int]),
    "fav_forward_mc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_u64, c_u64, c_void_p]),
    "fav_conv2d": (c_int, [c_void_p] + [c_void_p] * 5 + [c_int] * 12 + [c_void_p]),
    "fav_epilogue": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float] + [c_void_p] * 5 + [c_void_p]),
    "fav_hist_words": (c_size_t, [c_int, c_int, c_int]),
    "fav_accumulate": (c_int, [c_void_p] + [c_void_p] * 5 + [c_int, c_int, c_float, c_int, c_int, c_void_p, c_void_p]),
    "fav_epilogue_accumulate": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_int,
                                        c_void_p] + [c_void_p] * 5 + [c_void_p]),
    "fav_synth_images": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_u64, c_u64, c_void_p]),
    "fav_synth_labels": (c_int, [c_void_p, c_void_p, c_int, c_int, c_u64, c_u64, c_void_p]),
    "fav_frame_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "fav_trust_replay": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_double, c_int, c_int, c_void_p, c_void_p, c_void_p,
                         c_void_p, c_void_p, c_void_p]),
    "fav_comm_unique_id": (c_int, [c_void_p]),
    "fav_comm_init": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "fav_allreduce": (c_int, [c_void_p, c_void_p, C.c_size_t, c_void_p]),
    "fav_set_option": (c_int, [c_void_p, C.c_char_p, c_int]),
    "fav_launch_count": (c_u64, [c_void_p]),
    "fav_conv_timing_enable": (c_int, [c_void_p, c_int]),
    "fav_conv_timing_read": (c_int, [c_void_p, C.POINTER(c_float), C.POINTER(c_int)]),
    "fav_conv_stats_read": (c_int, [c_void_p, C.POINTER(c_u64), c_int, C.POINTER(c_int)]),
    "fav_conv_timing_read_all": (c_int, [c_void_p, C.POINTER(c_float), C.POINTER(c_float), c_int, C.POINTER(c_int)]),
    "fav_conv_timing_read_bytes": (c_int, [c_void_p, C.POINTER(c_float), c_int, C.POINTER(c_int)]),
}


def load():
    """dlopen libfav_b200.so and attach the prototypes.  Raises if the extension is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(failure-aware-vision_b200/csrc/build.sh).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().fav_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")


def f3(v):
    return _f3(*[float(x) for x in v])


PROFILE_FLAGS = {None: 0, "auto": 0, "cifar": 0x10, "imagenet": 0x20}


def corrupt_params(corruption_id, severity, height, width, profile=None):
    """Host-only (no GPU): (fparams list, iparams list, table uint8 ndarray or None) of one cell, as the library builds
    them (csrc/tables.cu) -- what fav_corrupt_normalize uses internally and fav_corrupt_normalize_ex takes as arguments."""
    import numpy as np
    lib = load()
    nf, ni, nb = c_int(0), c_int(0), c_size_t(0)
    flags = PROFILE_FLAGS[profile]
    check(min(0, lib.fav_corrupt_params(corruption_id, severity, height, width, flags, None, C.byref(nf), None, C.byref(ni),
                                        None, C.byref(nb))), "fav_corrupt_params")
    fp, ip = (c_float * max(1, nf.value))(), (C.c_int32 * max(1, ni.value))()
    tab = np.zeros(max(1, nb.value), dtype=np.uint8)
    rc = lib.fav_corrupt_params(corruption_id, severity, height, width, flags, fp, C.byref(nf), ip, C.byref(ni),
                                C.c_void_p(tab.ctypes.data), C.byref(nb))
    if rc != 1:
        check(rc if rc < 0 else -1, "fav_corrupt_params")
    return list(fp)[:nf.value], list(ip)[:ni.value], (tab[:nb.value] if nb.value else None)


def corruption_constants(profile, corruption_id, severity):
    """Severity constants of one cell as the library holds them (profile 'cifar' / 'imagenet')."""
    out = (C.c_double * 8)()
    n = load().fav_corruption_constants({"cifar": 0, "imagenet": 1}[profile], corruption_id, severity, out, 8)
    if n < 0:
        check(n, "fav_corruption_constants")
    return [out[i] for i in range(n)]


class Handle:
    """Owns one fav_handle (one per process / per connection, like the per-connection objects
    at main.py:110-118)."""

    def __init__(self, device=0):
        lib = load()
        h = c_void_p()
        check(lib.fav_init(int(device), C.byref(h)), "fav_init")
        self.lib, self.h, self.device = lib, h, int(device)

    def reset(self):
        check(self.lib.fav_reset(self.h), "fav_reset")

    def close(self):
        if getattr(self, "h", None):
            self.lib.fav_destroy(self.h)
            self.h = None

    def launches(self):
        return int(self.lib.fav_launch_count(self.h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
