"""Test-infrastructure harness (NOT product code).

* ``sys_decompress`` / ``sys_compress``: Google's C brotli 1.1.0 shipped in the image
  (libbrotlidec / libbrotlienc) through ctypes.  The decoder is the independent RFC 7932 round-trip oracle;
  the encoder is the size stand-in for the reference (the reference is a line-by-line
  port of that C code and cannot be built here because there is no Rust toolchain).
* ``Oracle``: ctypes binding of oracle/liboracle_brotli.so, the C restatement of the reference hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this module.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))

_enc = None
_dec = None


def _libs():
    global _enc, _dec
    if _enc is None:
        _enc = ctypes.CDLL("libbrotlienc.so.1")
        _dec = ctypes.CDLL("libbrotlidec.so.1")
        _enc.BrotliEncoderCompress.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
                                               ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p]
        _enc.BrotliEncoderCompress.restype = ctypes.c_int
        _dec.BrotliDecoderDecompress.argtypes = [ctypes.c_size_t, ctypes.c_char_p,
                                                 ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p]
        _dec.BrotliDecoderDecompress.restype = ctypes.c_int
    return _enc, _dec


def sys_compress(data: bytes, quality: int, lgwin: int) -> bytes:
    enc, _ = _libs()
    cap = len(data) + (len(data) >> 2) + 1024
    out = ctypes.create_string_buffer(cap)
    sz = ctypes.c_size_t(cap)
    ok = enc.BrotliEncoderCompress(quality, lgwin, 0, len(data), data, ctypes.byref(sz), out)
    if ok != 1:
        raise RuntimeError("BrotliEncoderCompress failed")
    return out.raw[:sz.value]


def sys_decompress(comp: bytes, max_out: int) -> bytes:
    """Decode with libbrotlidec; raises if the stream is invalid/truncated or exceeds max_out."""
    _, dec = _libs()
    cap = max_out + 16
    out = ctypes.create_string_buffer(cap)
    sz = ctypes.c_size_t(cap)
    res = dec.BrotliDecoderDecompress(len(comp), comp, ctypes.byref(sz), out)
    if res != 1:  # BROTLI_DECODER_RESULT_SUCCESS
        raise ValueError("libbrotlidec rejected the stream (result=%d)" % res)
    return out.raw[:sz.value]


def sys_decompress_with_dictionary(comp: bytes, max_out: int, dictionary: bytes) -> bytes:
    """Streaming decode through libbrotlidec with a raw (LZ77 prefix) dictionary attached -- what a decoder of a stream
    made after BrotliEncoderSetCustomDictionary has to do."""
    _, dec = _libs()
    dec.BrotliDecoderCreateInstance.restype = ctypes.c_void_p
    dec.BrotliDecoderCreateInstance.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    dec.BrotliDecoderAttachDictionary.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_char_p]
    dec.BrotliDecoderAttachDictionary.restype = ctypes.c_int
    dec.BrotliDecoderDecompressStream.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_void_p),
                                                  ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]
    dec.BrotliDecoderDecompressStream.restype = ctypes.c_int
    dec.BrotliDecoderDestroyInstance.argtypes = [ctypes.c_void_p]
    st = dec.BrotliDecoderCreateInstance(None, None, None)
    try:
        if not dec.BrotliDecoderAttachDictionary(st, 0, len(dictionary), dictionary):  # BROTLI_SHARED_DICTIONARY_RAW
            raise RuntimeError("BrotliDecoderAttachDictionary failed")
        out = ctypes.create_string_buffer(max_out + 16)
        avail_in, avail_out = ctypes.c_size_t(len(comp)), ctypes.c_size_t(len(out))
        next_in = ctypes.c_void_p(ctypes.cast(ctypes.c_char_p(comp), ctypes.c_void_p).value)
        next_out = ctypes.c_void_p(ctypes.addressof(out))
        total = ctypes.c_size_t(0)
        res = dec.BrotliDecoderDecompressStream(st, ctypes.byref(avail_in), ctypes.byref(next_in), ctypes.byref(avail_out),
                                                ctypes.byref(next_out), ctypes.byref(total))
        if res != 1:
            raise ValueError("libbrotlidec rejected the stream (result=%d)" % res)
        return out.raw[:len(out) - avail_out.value]
    finally:
        dec.BrotliDecoderDestroyInstance(st)


class OracleStats(ctypes.Structure):
    _fields_ = [("num_metablocks", ctypes.c_size_t), ("num_commands_total", ctypes.c_size_t),
                ("num_literals_total", ctypes.c_size_t), ("hasher_type", ctypes.c_int),
                ("bucket_bits", ctypes.c_int), ("block_bits", ctypes.c_int), ("hash_len", ctypes.c_int),
                ("n_last", ctypes.c_int)]


def build_oracle(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_brotli.so")
    src = os.path.join(_HERE, "brotli_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


class Oracle:
    def __init__(self):
        self.lib = ctypes.CDLL(build_oracle())
        L = self.lib
        L.oracle_brotli_compress.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t,
                                             ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                             ctypes.POINTER(OracleStats)]
        L.oracle_brotli_compress.restype = ctypes.c_size_t
        L.oracle_bits_entropy.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.oracle_bits_entropy.restype = ctypes.c_float
        L.oracle_population_cost.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.oracle_population_cost.restype = ctypes.c_float
        L.oracle_hash_keys.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t,
                                       ctypes.c_void_p]
        L.oracle_create_huffman_tree.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        L.oracle_convert_bit_depths_to_symbols.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.oracle_command_init.argtypes = [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
        L.oracle_compute_distance_code.argtypes = [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
        L.oracle_compute_distance_code.restype = ctypes.c_size_t
        L.oracle_backward_references.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t,
                                                 ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                                 ctypes.POINTER(ctypes.c_size_t)]
        L.oracle_backward_references.restype = ctypes.c_size_t
        L.oracle_optimize_huffman_counts_for_rle.argtypes = [ctypes.c_size_t, ctypes.c_void_p]
        L.oracle_build_and_store_huffman_tree.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_build_and_store_huffman_tree.restype = ctypes.c_size_t

    def compress(self, data: bytes, quality: int, lgwin: int, size_hint: int = 0, flags: int = 0):
        cap = len(data) + (len(data) >> 3) + 4096
        out = ctypes.create_string_buffer(cap)
        st = OracleStats()
        n = self.lib.oracle_brotli_compress(quality, lgwin, data, len(data), out, cap, size_hint, flags,
                                            ctypes.byref(st))
        if n == 0:
            raise RuntimeError("oracle_brotli_compress failed")
        return out.raw[:n], st
