"""ctypes face of oracle/libkta_oracle.so — TEST INFRASTRUCTURE ONLY (see oracle/kta_oracle.h).
Builds the oracle with its own Makefile if the shared object is missing."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libkta_oracle.so")

_lib = None


def build_oracle(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("kta_oracle.c", "kta_oracle.h", "Makefile")]
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(map(os.path.getmtime, src)):
        subprocess.run(["make", "-C", ORACLE_DIR, "-B"], check=True, capture_output=True)
    return ORACLE_SO


def olib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        P = C.c_void_p
        u64 = C.c_uint64
        L.kto_fnv32.restype = C.c_uint32
        L.kto_fnv32.argtypes = [P, C.c_size_t]
        L.kto_new.restype = P
        L.kto_new.argtypes = [C.c_int, C.c_int64, C.c_int32]
        L.kto_free.argtypes = [P]
        L.kto_handle_message.argtypes = [P, C.c_int32, C.c_int64, P, C.c_int32, C.c_int32]
        L.kto_handle_batch.argtypes = [P, C.c_int64, P, P, P, P, P]
        for f in ("total", "tombstones", "alive", "key_null", "key_non_null", "key_size_sum", "value_size_sum"):
            fn = getattr(L, "kto_" + f)
            fn.restype = u64
            fn.argtypes = [P, C.c_int32]
        for f in ("key_size_avg", "value_size_avg", "message_size_avg"):
            fn = getattr(L, "kto_" + f)
            fn.restype = C.c_int
            fn.argtypes = [P, C.c_int32, C.POINTER(u64)]
        L.kto_dirty_ratio.restype = C.c_float
        L.kto_dirty_ratio.argtypes = [P, C.c_int32]
        L.kto_earliest_message.argtypes = [P, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.kto_latest_message_s.restype = C.c_int64
        L.kto_latest_message_s.argtypes = [P]
        for f in ("smallest_message", "largest_message", "overall_count", "overall_size", "sum_all_alive"):
            fn = getattr(L, "kto_" + f)
            fn.restype = u64
            fn.argtypes = [P]
        L.kto_alive_contains.restype = C.c_int
        L.kto_alive_contains.argtypes = [P, C.c_uint32]
        L.kto_hist.argtypes = [P, C.c_int, C.c_int32, P]
        L.kto_hll_mix.restype = C.c_uint32
        L.kto_hll_mix.argtypes = [C.c_uint32]
        L.kto_hll_insert.argtypes = [P, C.c_int, C.c_uint32]
        L.kto_hll_estimate.restype = C.c_double
        L.kto_hll_estimate.argtypes = [P, C.c_int]
        L.kto_hll_stream_regs.argtypes = [P, C.c_int, P]
        L.kto_hll_alive_regs.argtypes = [P, C.c_int, P]
        L.kto_test_set_counter.argtypes = [P, C.c_int, C.c_int32, u64]
        _lib = L
    return _lib


COUNTERS = ("total", "tombstones", "alive", "key_null", "key_non_null", "key_size_sum", "value_size_sum")


def fnv32(b: bytes) -> int:
    buf = (C.c_char * max(len(b), 1)).from_buffer_copy(b or b"\0")
    return olib().kto_fnv32(C.cast(buf, C.c_void_p), len(b))


class Oracle:
    """The reference's two handlers, restated (oracle/kta_oracle.c)."""

    def __init__(self, count_alive_keys=False, track_stream=False, now=(4102444800, 123456789), no_hist=False):
        flags = (1 if count_alive_keys else 0) | (2 if track_stream else 0) | (4 if no_hist else 0)
        self.L = olib()
        self.o = self.L.kto_new(flags, now[0], now[1])

    def __del__(self):
        if getattr(self, "o", None):
            self.L.kto_free(self.o)
            self.o = None

    def handle_message(self, partition, ts_ms, key, value_len):
        """ts_ms None → not available; key None → null; value_len None → tombstone."""
        if key is None:
            self.L.kto_handle_message(self.o, partition, -1 if ts_ms is None else ts_ms, None, -1,
                                      -1 if value_len is None else value_len)
        else:
            buf = (C.c_char * max(len(key), 1)).from_buffer_copy(key or b"\0")
            self.L.kto_handle_message(self.o, partition, -1 if ts_ms is None else ts_ms, C.cast(buf, C.c_void_p),
                                      len(key), -1 if value_len is None else value_len)

    def handle_batch(self, partition, ts_ms, key_len, value_len, key_bytes):
        partition = np.ascontiguousarray(partition, dtype=np.int32)
        ts_ms = np.ascontiguousarray(ts_ms, dtype=np.int64)
        key_len = np.ascontiguousarray(key_len, dtype=np.int32)
        value_len = np.ascontiguousarray(value_len, dtype=np.int32)
        key_bytes = np.ascontiguousarray(key_bytes, dtype=np.uint8)
        kb = key_bytes.ctypes.data if key_bytes.size else None
        self.L.kto_handle_batch(self.o, partition.shape[0], partition.ctypes.data, ts_ms.ctypes.data,
                                key_len.ctypes.data, value_len.ctypes.data, kb)

    def counter(self, name, p):
        return getattr(self.L, "kto_" + name)(self.o, p)

    def avg(self, name, p):
        out = C.c_uint64()
        if getattr(self.L, "kto_" + name)(self.o, p, C.byref(out)):
            raise ZeroDivisionError("attempt to divide by zero (src/metric.rs:132-157)")
        return out.value

    def dirty_ratio(self, p):
        return self.L.kto_dirty_ratio(self.o, p)

    def earliest(self):
        s, ns = C.c_int64(), C.c_int32()
        self.L.kto_earliest_message(self.o, C.byref(s), C.byref(ns))
        return s.value, ns.value

    def latest(self):
        return self.L.kto_latest_message_s(self.o)

    def scalar(self, name):
        return getattr(self.L, "kto_" + name)(self.o)

    def hist(self, which, p):
        out = np.zeros(32, dtype=np.uint64)
        self.L.kto_hist(self.o, which, p, out.ctypes.data)
        return out

    def hll_stream_regs(self, precision):
        r = np.zeros(1 << precision, dtype=np.uint8)
        self.L.kto_hll_stream_regs(self.o, precision, r.ctypes.data)
        return r

    def hll_alive_regs(self, precision):
        r = np.zeros(1 << precision, dtype=np.uint8)
        self.L.kto_hll_alive_regs(self.o, precision, r.ctypes.data)
        return r

    def set_counter(self, name, p, v):
        self.L.kto_test_set_counter(self.o, COUNTERS.index(name), p, v)


def hll_estimate(regs, precision):
    regs = np.ascontiguousarray(regs, dtype=np.uint8)
    return olib().kto_hll_estimate(regs.ctypes.data, precision)
