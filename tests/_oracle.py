"""ctypes face of oracle/liboracle.so — the CPU checker (tests / smoke / bench cpu legs only)."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK_NVML = os.path.join(ROOT, "tests", "mock_nvml", "libnvidia-ml-mock.so")


class Verdict(C.Structure):
    _fields_ = [("index", C.c_int), ("uuid", C.c_char * 96), ("name", C.c_char * 96), ("mem_total", C.c_uint64),
                ("cc_major", C.c_int), ("cc_minor", C.c_int), ("healthy", C.c_int)]


def load():
    o = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    o.oracle_pattern_word.restype = C.c_uint32
    o.oracle_pattern_word.argtypes = [C.c_uint64, C.c_uint32]
    o.oracle_pattern_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32]
    o.oracle_checksum.argtypes = [C.c_void_p, C.c_uint64, u64p, u32p]
    o.oracle_pattern_checksum.argtypes = [C.c_uint64, C.c_uint32, u64p, u32p]
    o.oracle_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    o.oracle_verify.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, u64p, u64p]
    o.oracle_host_sweep.restype = C.c_double
    o.oracle_host_sweep.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint32, u64p, u32p]
    o.oracle_host_sweep_pinned.restype = C.c_double
    o.oracle_host_sweep_pinned.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint32, u64p, u32p]
    o.oracle_allowed_cpus.restype = C.c_int
    o.oracle_ph_time_poll_threads.restype = C.c_double
    o.oracle_ph_time_poll_threads.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    o.oracle_pin_self.argtypes = [C.c_int]
    o.oracle_a2a_chunk_seed.restype = C.c_uint32
    o.oracle_a2a_chunk_seed.argtypes = [C.c_uint32, C.c_int, C.c_int]
    o.oracle_gemm_elem.restype = C.c_double
    o.oracle_gemm_elem.argtypes = [C.c_uint64, C.c_uint32, C.c_int]
    o.oracle_gemm_elem_bits.restype = C.c_uint16
    o.oracle_gemm_elem_bits.argtypes = [C.c_uint64, C.c_uint32, C.c_int]
    o.oracle_gemm_dot.restype = C.c_double
    o.oracle_gemm_dot.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_int]
    o.oracle_bf16_rne.restype = C.c_uint16
    o.oracle_bf16_rne.argtypes = [C.c_float]
    o.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    o.oracle_gemm_uniform_bits.restype = C.c_uint16
    o.oracle_gemm_uniform_bits.argtypes = [C.c_uint64, C.c_uint32, C.c_int]
    o.oracle_gemm_uniform_elem.restype = C.c_double
    o.oracle_gemm_uniform_elem.argtypes = [C.c_uint64, C.c_uint32, C.c_int]
    o.oracle_gemm_uniform_dot.restype = C.c_double
    o.oracle_gemm_uniform_dot.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_int]
    o.oracle_gemm_uniform_tol.restype = C.c_double
    o.oracle_gemm_uniform_tol.argtypes = [C.c_int, C.c_double]
    o.oracle_ph_open.argtypes = [C.c_char_p, C.c_char_p]
    o.oracle_ph_poll.argtypes = [C.c_int]
    o.oracle_ph_verdicts.argtypes = [C.POINTER(Verdict), C.c_int, C.POINTER(C.c_int)]
    o.oracle_ph_time_enumerate.restype = C.c_double
    o.oracle_ph_time_enumerate.argtypes = [C.c_int]
    o.oracle_ph_time_poll.restype = C.c_double
    o.oracle_ph_time_poll.argtypes = [C.c_int, C.c_int]
    return o


def philox(o, ctr, key):
    c, k, out = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
    o.oracle_philox4x32_10(c, k, out)
    return list(out)


def pattern_checksum(o, words, seed):
    s, x = C.c_uint64(), C.c_uint32()
    o.oracle_pattern_checksum(words, seed, C.byref(s), C.byref(x))
    return s.value, x.value


def checksum(o, arr):
    """arr: C-contiguous numpy array; checksum over its u32 words."""
    s, x = C.c_uint64(), C.c_uint32()
    o.oracle_checksum(arr.ctypes.data, arr.nbytes // 4, C.byref(s), C.byref(x))
    return s.value, x.value


def verify(o, arr, seed, first_word=0):
    """(bad_words, first_bad_word) of a C-contiguous numpy buffer against the pattern; first = 2**64-1 when clean."""
    bad, first = C.c_uint64(), C.c_uint64()
    o.oracle_verify(arr.ctypes.data, first_word, arr.nbytes // 4, seed, C.byref(bad), C.byref(first))
    return bad.value, first.value


def pattern(o, first_word, words, seed):
    import numpy as np

    a = np.empty(words, dtype=np.uint32)
    o.oracle_pattern_fill(a.ctypes.data, first_word, words, seed)
    return a


def verdicts(o):
    arr = (Verdict * 64)()
    n = C.c_int()
    rc = o.oracle_ph_verdicts(arr, 64, C.byref(n))
    assert rc == 0
    return [(v.index, v.uuid.decode(), v.name.decode(), v.mem_total, (v.cc_major, v.cc_minor), "Healthy" if v.healthy else "Unhealthy")
            for v in arr[: n.value]]
