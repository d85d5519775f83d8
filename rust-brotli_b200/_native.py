"""ctypes binding of libbrotli_b200.so (built in-tree by build.py / __graft_entry__.build()).

There is no fallback: if the shared library is missing or CUDA is unavailable, every call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "libbrotli_b200.so")  # B200_LIB: A/B builds (tools/)

NUM_STAGES = 7
STAGE_NAMES = ("sort", "match", "parse", "finalize", "split", "header", "emit")
OPT_UNIT, OPT_MB_UNITS, OPT_LCAP, OPT_RLE_OPT, OPT_SPLIT, OPT_CTX_MODEL, OPT_TIMING, OPT_LANES, OPT_DICT, OPT_SHALLOW_MATCH, OPT_PAIR_PARSE = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11
OPT_HQ_SPLIT, OPT_HQ_UNIT, OPT_HQ_THREAD_UNITS, OPT_ONDEMAND, OPT_HQ_LEVELS = 12, 13, 14, 15, 16

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libbrotli_b200.so is not built (run `python __graft_entry__.py build`); "
                               "there is no CPU fallback for the compression path")
        L = ctypes.CDLL(LIB_PATH)
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        L.b200_device_count.restype = ctypes.c_int
        L.b200_encoder_create.restype = vp
        L.b200_encoder_create.argtypes = [ctypes.c_int]
        L.b200_encoder_destroy.argtypes = [vp]
        L.b200_encoder_set_option.argtypes = [vp, ctypes.c_int, ctypes.c_uint32]
        L.b200_encoder_set_option.restype = ctypes.c_int
        L.b200_max_compressed_size.argtypes = [sz]
        L.b200_max_compressed_size.restype = sz
        L.b200_encoder_compress.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, sz, vp, sz, ctypes.POINTER(sz), ctypes.c_int]
        L.b200_encoder_compress.restype = ctypes.c_int
        L.b200_encoder_compress_range.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, vp, sz, sz, sz, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_int, vp, sz, ctypes.POINTER(sz), ctypes.c_int]
        L.b200_encoder_compress_range.restype = ctypes.c_int
        L.b200_encoder_last_timings.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint32)]
        L.b200_stage_match.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, sz, vp]
        L.b200_stage_match.restype = ctypes.c_int
        _lib = L
    return _lib


def _inptr(data):
    return ctypes.cast(ctypes.c_char_p(data), ctypes.c_void_p) if len(data) else ctypes.c_void_p(0)


class DeviceEncoder:
    """One GPU, one stream, one reusable workspace (wraps B200Encoder*)."""

    def __init__(self, device: int = 0):
        self._L = lib()
        self._h = self._L.b200_encoder_create(device)
        if not self._h:
            raise RuntimeError("b200_encoder_create(%d) failed: no usable CUDA device" % device)
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, opt, value):
        if not self._L.b200_encoder_set_option(self._h, opt, int(value)):
            raise ValueError("bad option")

    def compress(self, data: bytes, quality: int = 5, lgwin: int = 22) -> bytes:
        n = len(data)
        cap = self._L.b200_max_compressed_size(n)
        out = ctypes.create_string_buffer(cap)
        osz = ctypes.c_size_t(0)
        ok = self._L.b200_encoder_compress(self._h, quality, lgwin, _inptr(data), n, ctypes.cast(out, ctypes.c_void_p), cap,
                                           ctypes.byref(osz), 0)
        if not ok:
            raise RuntimeError("b200_encoder_compress failed")
        return out.raw[:osz.value]

    def compress_range(self, data: bytes, start: int, length: int, quality: int, lgwin: int, first: bool, last: bool,
                       byte_align: bool, size_hint: int = 0) -> bytes:
        cap = self._L.b200_max_compressed_size(length)
        out = ctypes.create_string_buffer(cap)
        osz = ctypes.c_size_t(0)
        ok = self._L.b200_encoder_compress_range(self._h, quality, lgwin, size_hint, _inptr(data), len(data), start, length,
                                                 int(first), int(last), int(byte_align), ctypes.cast(out, ctypes.c_void_p), cap,
                                                 ctypes.byref(osz), 0)
        if not ok:
            raise RuntimeError("b200_encoder_compress_range failed")
        return out.raw[:osz.value]

    def compress_device(self, d_in_ptr: int, n: int, d_out_ptr: int, out_cap: int, quality: int = 5, lgwin: int = 22) -> int:
        """Device-resident input/output (raw CUDA pointers, e.g. torch tensor .data_ptr()); returns compressed size."""
        osz = ctypes.c_size_t(0)
        ok = self._L.b200_encoder_compress(self._h, quality, lgwin, ctypes.c_void_p(d_in_ptr), n, ctypes.c_void_p(d_out_ptr),
                                           out_cap, ctypes.byref(osz), 1)
        if not ok:
            raise RuntimeError("b200_encoder_compress (device io) failed")
        return osz.value

    def timings(self):
        ms = (ctypes.c_float * NUM_STAGES)()
        launches = ctypes.c_uint32(0)
        self._L.b200_encoder_last_timings(self._h, ms, ctypes.byref(launches))
        return dict(zip(STAGE_NAMES, [float(x) for x in ms])), int(launches.value)

    def stage_match(self, data: bytes, quality: int, lgwin: int):
        import numpy as np
        out = np.zeros(len(data), dtype=np.uint32)
        ok = self._L.b200_stage_match(self._h, quality, lgwin, _inptr(data), len(data), out.ctypes.data)
        if not ok:
            raise RuntimeError("b200_stage_match failed")
        return out
