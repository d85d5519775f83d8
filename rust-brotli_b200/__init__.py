"""rust-brotli_b200 -- host-side mirror of the reference's compression API surface over the C ABI.

The reference (dropbox/rust-brotli) exposes, for the compression path:
  * ``BrotliEncoderParams``                      src/enc/backward_references/mod.rs:71, defaults src/enc/encode.rs:318
  * ``CompressorWriter`` / ``CompressorReader``  src/enc/writer.rs:74, src/enc/reader.rs:65
  * ``BrotliCompress(r, w, params)``             src/enc/mod.rs:142
  * ``compress_multi(params, input, ...)``       src/enc/mod.rs:95
  * ``BrotliEncoderMaxCompressedSize{,Multi}``   src/enc/encode.rs:1273-1299
The host language of the reference is Rust; this image has no Rust toolchain, so the mirror is Python (what the
reference's own c/py wrapper does) calling the same ``BrotliEncoder*`` C ABI a Rust shim would bind -- see
INTEGRATION.md.  Every byte of compressed output comes from the CUDA library; nothing here compresses on the CPU
and importing this package on a machine without the built library raises at first use.
"""
import ctypes
import io
from dataclasses import dataclass

from . import _native
from ._native import DeviceEncoder, lib  # noqa: F401

# BrotliEncoderParameter (src/enc/parameters.rs:1-32)
BROTLI_PARAM_MODE, BROTLI_PARAM_QUALITY, BROTLI_PARAM_LGWIN, BROTLI_PARAM_LGBLOCK = 0, 1, 2, 3
BROTLI_PARAM_DISABLE_LITERAL_CONTEXT_MODELING, BROTLI_PARAM_SIZE_HINT, BROTLI_PARAM_LARGE_WINDOW = 4, 5, 6
BROTLI_PARAM_CATABLE, BROTLI_PARAM_APPENDABLE, BROTLI_PARAM_MAGIC_NUMBER = 167, 168, 169
BROTLI_PARAM_NO_DICTIONARY, BROTLI_PARAM_BYTE_ALIGN, BROTLI_PARAM_BARE_STREAM = 170, 172, 173
BROTLI_OPERATION_PROCESS, BROTLI_OPERATION_FLUSH, BROTLI_OPERATION_FINISH = 0, 1, 2
MAX_THREADS = 16  # src/enc/fixed_queue.rs:1


class BrotliEncoderThreadError(Exception):
    """src/enc/threading/mod.rs:33-40"""


class InsufficientOutputSpace(BrotliEncoderThreadError):
    pass


class OtherThreadPanic(BrotliEncoderThreadError):
    pass


@dataclass
class BrotliEncoderParams:
    """Subset of the reference struct that parameterises this path (defaults: encode.rs:318-357)."""
    quality: int = 11
    lgwin: int = 22
    lgblock: int = 0
    size_hint: int = 0
    mode: int = 0
    disable_literal_context_modeling: int = 0
    catable: bool = False
    appendable: bool = False
    magic_number: bool = False
    byte_align: bool = False
    bare_stream: bool = False
    use_dictionary: bool = True

    def as_key_values(self):
        kv = [(BROTLI_PARAM_QUALITY, self.quality), (BROTLI_PARAM_LGWIN, self.lgwin), (BROTLI_PARAM_MODE, self.mode)]
        if self.size_hint:
            kv.append((BROTLI_PARAM_SIZE_HINT, min(self.size_hint, 0xFFFFFFFF)))
        if self.disable_literal_context_modeling:
            kv.append((BROTLI_PARAM_DISABLE_LITERAL_CONTEXT_MODELING, 1))
        if self.lgblock:
            kv.append((BROTLI_PARAM_LGBLOCK, self.lgblock))
        if not self.use_dictionary:
            kv.append((BROTLI_PARAM_NO_DICTIONARY, 1))
        # framing parameters are forwarded, never dropped: the C ABI refuses the ones this path cannot produce
        for key, on in ((BROTLI_PARAM_CATABLE, self.catable), (BROTLI_PARAM_APPENDABLE, self.appendable),
                        (BROTLI_PARAM_MAGIC_NUMBER, self.magic_number), (BROTLI_PARAM_BYTE_ALIGN, self.byte_align),
                        (BROTLI_PARAM_BARE_STREAM, self.bare_stream)):
            if on:
                kv.append((key, 1))
        return kv


def _capi():
    L = lib()
    if not getattr(L, "_capi_ready", False):
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        L.BrotliEncoderCreateInstance.restype = vp
        L.BrotliEncoderCreateInstance.argtypes = [vp, vp, vp]
        L.BrotliEncoderDestroyInstance.argtypes = [vp]
        L.BrotliEncoderSetParameter.argtypes = [vp, ctypes.c_int, ctypes.c_uint32]
        L.BrotliEncoderSetParameter.restype = ctypes.c_int
        L.BrotliEncoderCompressStream.argtypes = [vp, ctypes.c_int, ctypes.POINTER(sz), ctypes.POINTER(vp), ctypes.POINTER(sz),
                                                  ctypes.POINTER(vp), ctypes.POINTER(sz)]
        L.BrotliEncoderCompressStream.restype = ctypes.c_int
        L.BrotliEncoderIsFinished.argtypes = [vp]
        L.BrotliEncoderHasMoreOutput.argtypes = [vp]
        L.BrotliEncoderTakeOutput.argtypes = [vp, ctypes.POINTER(sz)]
        L.BrotliEncoderTakeOutput.restype = vp
        L.BrotliEncoderMaxCompressedSize.argtypes = [sz]
        L.BrotliEncoderMaxCompressedSize.restype = sz
        L.BrotliEncoderMaxCompressedSizeMulti.argtypes = [sz, sz]
        L.BrotliEncoderMaxCompressedSizeMulti.restype = sz
        L.BrotliEncoderCompress.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, sz, vp, ctypes.POINTER(sz), vp]
        L.BrotliEncoderCompress.restype = ctypes.c_int
        L.BrotliEncoderCompressMulti.argtypes = [sz, vp, vp, sz, vp, ctypes.POINTER(sz), vp, sz, vp, vp, vp]
        L.BrotliEncoderCompressMulti.restype = ctypes.c_int32
        L.BrotliEncoderVersion.restype = ctypes.c_uint32
        L.BrotliEncoderSetCustomDictionary.argtypes = [vp, sz, vp]
        L.BrotliEncoderSetCustomDictionary.restype = None
        L.BrotliEncoderCompressStreaming.argtypes = [vp, ctypes.c_int, ctypes.POINTER(sz), vp, ctypes.POINTER(sz), vp]
        L.BrotliEncoderCompressStreaming.restype = ctypes.c_int
        L.b200_effective_quality.argtypes = [ctypes.c_int]
        L.b200_effective_quality.restype = ctypes.c_int
        L._capi_ready = True
    return L


def BrotliEncoderMaxCompressedSize(input_size: int) -> int:
    return _capi().BrotliEncoderMaxCompressedSize(input_size)


def BrotliEncoderMaxCompressedSizeMulti(input_size: int, num_threads: int) -> int:
    return _capi().BrotliEncoderMaxCompressedSizeMulti(input_size, num_threads)


def encoder_compress(data: bytes, quality: int = 11, lgwin: int = 22) -> bytes:
    """One-shot ``BrotliEncoderCompress`` (src/ffi/compressor.rs:194)."""
    L = _capi()
    cap = L.BrotliEncoderMaxCompressedSize(len(data)) + 16
    out = ctypes.create_string_buffer(cap)
    osz = ctypes.c_size_t(cap)
    ok = L.BrotliEncoderCompress(quality, lgwin, 0, len(data), _native._inptr(data), ctypes.byref(osz), ctypes.cast(out, ctypes.c_void_p))
    if not ok:
        raise IOError("BrotliEncoderCompress failed (no CUDA device or output too small)")
    return out.raw[:osz.value]


class _Stream:
    """BrotliEncoderState driven through BrotliEncoderCompressStream, as writer.rs / reader.rs do."""

    def __init__(self, params: BrotliEncoderParams):
        self.L = _capi()
        self.h = self.L.BrotliEncoderCreateInstance(None, None, None)
        if not self.h:
            raise IOError("BrotliEncoderCreateInstance failed: no usable CUDA device")
        for k, v in params.as_key_values():
            if not self.L.BrotliEncoderSetParameter(self.h, k, int(v)):
                self.close()
                raise ValueError("BrotliEncoderSetParameter(%d, %d) refused: not produced by this path" % (k, int(v)))

    def step(self, data: bytes, op: int) -> bytes:
        out = bytearray()
        avail_in = ctypes.c_size_t(len(data))
        inbuf = ctypes.create_string_buffer(data, len(data)) if data else None
        next_in = ctypes.c_void_p(ctypes.addressof(inbuf) if data else 0)
        obuf = ctypes.create_string_buffer(1 << 16)
        while True:
            avail_out = ctypes.c_size_t(len(obuf))
            next_out = ctypes.c_void_p(ctypes.addressof(obuf))
            total = ctypes.c_size_t(0)
            ok = self.L.BrotliEncoderCompressStream(self.h, op, ctypes.byref(avail_in), ctypes.byref(next_in), ctypes.byref(avail_out),
                                                    ctypes.byref(next_out), ctypes.byref(total))
            if not ok:
                raise IOError("BrotliEncoderCompressStream failed")  # io::ErrorKind::InvalidData in writer.rs:43-44
            out += obuf.raw[: len(obuf) - avail_out.value]
            if avail_in.value == 0 and not self.L.BrotliEncoderHasMoreOutput(self.h):
                break
        return bytes(out)

    def close(self):
        if self.h:
            self.L.BrotliEncoderDestroyInstance(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CompressorWriter:
    """``CompressorWriter::new(w, buffer_size, q, lgwin)`` / ``with_params`` (src/enc/writer.rs:83-118)."""

    def __init__(self, w, buffer_size: int = 4096, q: int = 11, lgwin: int = 22, params: BrotliEncoderParams = None):
        self._w = w
        self._params = params or BrotliEncoderParams(quality=q, lgwin=lgwin)
        self._s = _Stream(self._params)
        self._closed = False
        self.buffer_size = buffer_size

    @classmethod
    def with_params(cls, w, buffer_size, params):
        return cls(w, buffer_size, params=params)

    def write(self, buf: bytes) -> int:
        out = self._s.step(bytes(buf), BROTLI_OPERATION_PROCESS)
        if out:
            self._w.write(out)
        return len(buf)

    def flush(self):
        out = self._s.step(b"", BROTLI_OPERATION_FLUSH)
        if out:
            self._w.write(out)
        if hasattr(self._w, "flush"):
            self._w.flush()

    def close(self):
        """Finishes the stream (the reference does this on Drop, writer.rs:253-265)."""
        if not self._closed:
            out = self._s.step(b"", BROTLI_OPERATION_FINISH)
            if out:
                self._w.write(out)
            self._s.close()
            self._closed = True

    def into_inner(self):
        self.close()
        return self._w

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class CompressorReader:
    """``CompressorReader::new(r, buffer_size, q, lgwin)`` (src/enc/reader.rs:74-103): reads raw, yields compressed."""

    def __init__(self, r, buffer_size: int = 4096, q: int = 11, lgwin: int = 22, params: BrotliEncoderParams = None):
        self._r = r
        self._params = params or BrotliEncoderParams(quality=q, lgwin=lgwin)
        self._s = _Stream(self._params)
        self._buf = bytearray()
        self._eof = False
        self.buffer_size = max(1, buffer_size)

    def read(self, n: int = -1) -> bytes:
        while not self._eof and (n < 0 or len(self._buf) < n):
            chunk = self._r.read(self.buffer_size)
            if chunk:
                self._buf += self._s.step(chunk, BROTLI_OPERATION_PROCESS)
            else:
                self._buf += self._s.step(b"", BROTLI_OPERATION_FINISH)
                self._s.close()
                self._eof = True
        if n < 0:
            out, self._buf = bytes(self._buf), bytearray()
        else:
            out, self._buf = bytes(self._buf[:n]), self._buf[n:]
        return out

    def into_inner(self):
        return self._r


def BrotliCompress(r, w, params: BrotliEncoderParams) -> int:
    """``BrotliCompress(r, w, &params) -> io::Result<usize>`` (src/enc/mod.rs:142): returns bytes written."""
    cw = CompressorWriter.with_params(_CountingWriter(w), 4096, params)
    while True:
        chunk = r.read(1 << 20)
        if not chunk:
            break
        cw.write(chunk)
    cw.close()
    return cw._w.count


class _CountingWriter:
    def __init__(self, w):
        self.w = w
        self.count = 0

    def write(self, b):
        self.count += len(b)
        return self.w.write(b)


def compress_multi(params: BrotliEncoderParams, input_bytes: bytes, num_threads: int = 1) -> bytes:
    """``compress_multi`` (src/enc/mod.rs:95-133; CompressMulti src/enc/threading/mod.rs:413).

    The input is split into ``num_threads`` (<= 16) equal ranges (threading/mod.rs:333); range i > 0 sees the previous
    2^lgwin bytes as its LZ77 window.  Ranges are placed round-robin on the visible GPUs and their byte-aligned
    outputs are concatenated.  Raises BrotliEncoderThreadError subclasses on failure.
    """
    L = _capi()
    if num_threads < 1 or num_threads > MAX_THREADS:
        raise BrotliEncoderThreadError("num_threads must be in 1..=%d" % MAX_THREADS)
    kv = params.as_key_values()
    keys = (ctypes.c_int * len(kv))(*[k for k, _ in kv])
    vals = (ctypes.c_uint32 * len(kv))(*[int(v) for _, v in kv])
    cap = L.BrotliEncoderMaxCompressedSizeMulti(len(input_bytes), num_threads) + 64
    out = ctypes.create_string_buffer(cap)
    osz = ctypes.c_size_t(cap)
    ok = L.BrotliEncoderCompressMulti(len(kv), ctypes.cast(keys, ctypes.c_void_p), ctypes.cast(vals, ctypes.c_void_p), len(input_bytes),
                                      _native._inptr(input_bytes), ctypes.byref(osz), ctypes.cast(out, ctypes.c_void_p), num_threads,
                                      None, None, None)
    if not ok:
        raise OtherThreadPanic("BrotliEncoderCompressMulti failed")
    return out.raw[:osz.value]
