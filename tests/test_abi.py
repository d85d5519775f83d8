"""CPU: the C-ABI shared library loads and exports every symbol that include/brotli_b200.h declares; without a GPU
the compression entry points fail loudly instead of falling back to any CPU path."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rust-brotli_b200", "libbrotli_b200.so")
HDR = os.path.join(ROOT, "include", "brotli_b200.h")


def _declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b((?:BrotliEncoder|b200_)\w+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_func")))


def test_library_is_built():
    assert os.path.exists(LIB), "run `python __graft_entry__.py build`"


def test_exports_every_declared_symbol():
    L = ctypes.CDLL(LIB)
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export " + n


def test_parameter_enum_values_match_reference():
    """src/enc/parameters.rs:1-32"""
    src = open(HDR).read()
    for name, val in [("BROTLI_PARAM_MODE", 0), ("BROTLI_PARAM_QUALITY", 1), ("BROTLI_PARAM_LGWIN", 2), ("BROTLI_PARAM_LGBLOCK", 3),
                      ("BROTLI_PARAM_SIZE_HINT", 5), ("BROTLI_PARAM_Q9_5", 150), ("BROTLI_PARAM_CATABLE", 167),
                      ("BROTLI_PARAM_APPENDABLE", 168), ("BROTLI_PARAM_MAGIC_NUMBER", 169), ("BROTLI_PARAM_BARE_STREAM", 173)]:
        assert re.search(r"\b%s\s*=\s*%d\b" % (name, val), src)


def test_pure_host_functions():
    L = ctypes.CDLL(LIB)
    L.BrotliEncoderMaxCompressedSize.restype = ctypes.c_size_t
    L.BrotliEncoderMaxCompressedSize.argtypes = [ctypes.c_size_t]
    def ref(n):  # encode.rs:1277-1299 restated (64-bit wrapping arithmetic)
        M = (1 << 64) - 1
        blocks = n >> 14
        tail = (n - (blocks << 24)) & M
        res = (n + 2 + 4 * blocks + (4 if tail > (1 << 20) else 3) + 1) & M
        return 17 if n == 0 else (0 if res < n else res + 16)
    assert L.BrotliEncoderMaxCompressedSize(0) == 17
    for n in (1, 1000, 16384, 1 << 20, (1 << 24) + 5, 1 << 30, 5 << 30):
        assert L.BrotliEncoderMaxCompressedSize(n) == ref(n) > n
    L.BrotliEncoderMaxCompressedSizeMulti.restype = ctypes.c_size_t
    L.BrotliEncoderMaxCompressedSizeMulti.argtypes = [ctypes.c_size_t, ctypes.c_size_t]
    assert L.BrotliEncoderMaxCompressedSizeMulti(1000, 8) == ref(1000) + 64  # encode.rs:1273-1275
    L.b200_effective_quality.restype = ctypes.c_int
    assert [L.b200_effective_quality(q) for q in (0, 4, 5, 9)] == [5, 5, 5, 9]
    L.BrotliEncoderVersion.restype = ctypes.c_uint32
    assert L.BrotliEncoderVersion() >> 24 == 8


def _has_gpu():
    L = ctypes.CDLL(LIB)
    L.b200_device_count.restype = ctypes.c_int
    return L.b200_device_count() > 0


@pytest.mark.skipif(_has_gpu(), reason="this check is for machines without a CUDA device")
def test_no_cpu_fallback_without_gpu():
    import rust_brotli_b200 as rb
    with pytest.raises((IOError, RuntimeError)):
        rb.encoder_compress(b"hello hello hello hello", 5, 22)
    L = ctypes.CDLL(LIB)
    L.BrotliEncoderCreateInstance.restype = ctypes.c_void_p
    assert not L.BrotliEncoderCreateInstance(None, None, None)
