"""CPU: the sequential model of the GPU pipeline (tools/gpu_model.cpp, built from the same host/device headers as the
kernels).  It must produce valid brotli, stay within +0.5 % of the reference restatement, and match its goldens --
the GPU tests then assert that the kernels reproduce these streams bit for bit."""
import hashlib
import os

import pytest

from conftest import assert_size_parity, golden_bytes
from oracle.harness import sys_decompress

FILES = ["alice29.txt", "asyoulik.txt", "random_then_unicode", "quickfox_repeated", "random_org_10k.bin", "backward65536",
         "64x", "ukkonooa", "monkey", "x", "xyzzy", "10x10y", "aaabaaaa", "empty", "quickfox", "compressed_file"]


@pytest.mark.parametrize("name", FILES)
@pytest.mark.parametrize("q,w", [(5, 20), (5, 22), (7, 22), (9, 22), (9, 16), (5, 18), (10, 22), (11, 22), (11, 24), (10, 16)])
def test_model_golden_roundtrip_and_size(model, golden_table, name, q, w):
    d = golden_bytes(name)
    c, _ = model.compress(d, q, w)
    assert sys_decompress(c, len(d)) == d
    g = golden_table["%s|q%d|w%d" % (name, q, w)]
    assert hashlib.sha256(c).hexdigest() == g["model_sha256"]
    assert_size_parity(len(c), g["oracle_size"], "%s q%d w%d" % (name, q, w))


def test_model_against_reference_kats_q10_q11(model):
    """The reference's own exact size vectors for the binary-tree / Zopfli qualities (src/bin/integration_tests.rs:408-449):
    alice29.txt, lgwin 22: quality 10 -> 47 488 B, quality 11 -> 46 493 B.  The pipeline must land within +-0.5 %."""
    d = golden_bytes("alice29.txt")
    for q, pin in ((10, 47488), (11, 46493)):
        c, _ = model.compress(d, q, 22)
        assert sys_decompress(c, len(d)) == d
        assert abs(len(c) - pin) <= pin * 0.005, (q, len(c), pin)


def test_model_hq_options(model):
    """quality 10 / 11 knobs: every variant is a valid stream; the histogram stage (BrotliSplitBlock + context maps) and the
    static dictionary each pay for themselves on English text."""
    d = golden_bytes("asyoulik.txt")
    base = len(model.compress(d, 10, 22)[0])
    for kw in ({"hq_split": 0}, {"use_dict": 0}, {"ctx_model": 0}, {"unit": 65536, "mb_units": 64}, {"depth": 1024}):
        c, _ = model.compress(d, 10, 22, **kw)
        assert sys_decompress(c, len(d)) == d
        if "unit" not in kw and "depth" not in kw:
            assert len(c) > base, kw


@pytest.mark.parametrize("shards", [1, 2, 3, 5])
def test_model_sharded_seams(model, shards):
    """compress_multi split rule (threading/mod.rs:333) with byte-aligned seams: concatenation must decode."""
    d = golden_bytes("random_then_unicode")
    n = len(d)
    parts = []
    for i in range(shards):
        a, b = i * n // shards, (i + 1) * n // shards
        c, _ = model.compress_range(d, a, b - a, 5, 22, i == 0, i + 1 == shards, i + 1 != shards)
        parts.append(c)
    out = b"".join(parts)
    assert sys_decompress(out, n) == d
    if shards == 3:
        assert len(out) <= 144325  # src/bin/test_threading.rs:101 bound for 3 threads q5


@pytest.mark.parametrize("n", [0, 1, 2, 3, 7, 8, 9, 63, 64, 65, 4095, 4096, 4097, 8191, 8192, 12289])
def test_model_edge_sizes(model, n):
    d = (golden_bytes("alice29.txt") * 2)[:n]
    c, _ = model.compress(d, 5, 22)
    assert sys_decompress(c, max(n, 1)) == d


def test_model_options(model):
    d = golden_bytes("asyoulik.txt")
    base, _ = model.compress(d, 5, 22)
    for kw in ({"split": 0}, {"ctx_model": 0}, {"use_rle_opt": 0}, {"unit": 2048}, {"unit": 16384}, {"lcap": 32}, {"mb_units": 8}):
        c, _ = model.compress(d, 5, 22, **kw)
        assert sys_decompress(c, len(d)) == d
        assert len(c) < len(base) * 1.03


def test_windowed_parse_formulation_equals_sequential_spec(model):
    """The parse kernels resolve a window of G positions with the distance cache of the window start and then walk it with
    straight-line predicated code (G = 8: one unit per warp, G = 4: two units per warp).  tools/window_emul.cpp is that
    formulation on the CPU; it must reproduce parse_range() command for command, for both window sizes."""
    import ctypes
    import subprocess
    import numpy as np
    from conftest import assert_size_parity, golden_bytes
    from tools.model_harness import EncParams
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "tools", "libwindow_emul.so")
    src = os.path.join(root, "tools", "window_emul.cpp")
    if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-fwrapv", "-std=c++17", "-shared", "-fPIC", "-w", "-I",
                               os.path.join(root, "rust-brotli_b200", "csrc"), src, "-o", so])
    L = ctypes.CDLL(so)
    L.window_emul_check.argtypes = [ctypes.POINTER(EncParams), ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint32,
                                    ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    for name in ("alice29.txt", "random_then_unicode", "compressed_file", "quickfox_repeated"):
        d = golden_bytes(name)
        for q in (5, 6, 7, 9):  # q7 / q9: 10 / 16 cache candidates, the H9 scores, the 512-byte literal-spree window
            p = model.params(q, 22, len(d), len(d))
            best = np.zeros(len(d) + 1, dtype=np.uint32)
            model.compress(d, q, 22, best_out=best.ctypes.data)
            w4, w8 = ctypes.c_uint64(0), ctypes.c_uint64(0)
            bad = L.window_emul_check(ctypes.byref(p), d + bytes(512), best.ctypes.data, len(d), ctypes.byref(w4), ctypes.byref(w8))
            assert bad == 0, (name, q)
            assert w4.value <= w8.value * 1.25  # G = 4 costs few extra windows


def test_model_two_chunks_roundtrip(model):
    """Streams longer than one 24 MiB chunk: the chunk loop (shared with the device encoder) appends independent chunks to one
    bit stream; matches may reach back across the seam (left halo), nothing may read across it to the right."""
    from tools import datagen
    d = datagen.enwik_like(26_000_000, seed=11)
    c, st = model.compress(d, 5, 22)
    assert st.num_metablocks == 7  # 6 in the first chunk, 1 in the second
    assert sys_decompress(c, len(d)) == d
    assert len(c) < 0.36 * len(d)


def test_model_structured_logs_within_half_percent_of_reference_restatement(model, oracle):
    """The 256-byte warm-up in front of every parse unit keeps record-structured input at reference size
    (+1.4 % without it)."""
    from tools import datagen
    d = datagen.json_logs(4_000_000)
    for q in (5, 9):
        c = model.compress(d, q, 22)[0]
        assert sys_decompress(c, len(d)) == d
        assert len(c) <= len(oracle.compress(d, q, 22)[0]) * 1.005


def test_model_static_dictionary_reaches_libbrotlienc_on_english(model):
    """Config 1 (alice29, q5, lgwin 20): with static-dictionary matches the size is within 0.5 % of Google's encoder, whose
    code the reference is a port of (52 809 B); without them it is 0.7 % larger."""
    from oracle.harness import sys_compress
    d = golden_bytes("alice29.txt")
    on = model.compress(d, 5, 20)[0]
    off = model.compress(d, 5, 20, use_dict=0)[0]
    ref = len(sys_compress(d, 5, 20))
    assert sys_decompress(on, len(d)) == d and sys_decompress(off, len(d)) == d
    assert len(on) <= ref * 1.005 < len(off)


def _catable_from_model(model, d, q, lgwin=22, byte_align=False):
    """The framing csrc/bro_capi.cu:compress_framed builds around a catable stream, assembled here from the CPU model's ranges:
    window bits, the first two bytes as an uncompressed metablock, the rest without static dictionary, empty last metablock."""
    bits = []
    def put(n, v):
        bits.extend((v >> i) & 1 for i in range(n))
    put(4, ((lgwin - 17) << 1) | 1)
    n2 = min(2, len(d))
    if n2:
        put(1, 0); put(2, 0); put(16, n2 - 1); put(1, 1)
        bits.extend([0] * (-len(bits) % 8))
        for byte in d[:n2]:
            put(8, byte)
    head = bytes(sum(bits[i + j] << j for j in range(8)) for i in range(0, len(bits) - len(bits) % 8, 8))
    if len(d) <= 2:
        tail_bits = bits[len(head) * 8:] + [1, 1]
        tail_bits += [0] * (-len(tail_bits) % 8)
        return head + bytes(sum(tail_bits[i + j] << j for j in range(8)) for i in range(0, len(tail_bits), 8))
    if byte_align:
        body, _ = model.compress_range(d, 2, len(d) - 2, q, lgwin, False, False, True, use_dict=0)
        return head + body + b"\x03"
    body, _ = model.compress_range(d, 2, len(d) - 2, q, lgwin, False, True, False, use_dict=0)
    return head + body


@pytest.mark.parametrize("q", [5, 10])
def test_catable_framing_stitches_like_brocatli(model, q):
    """Streams framed as BROTLI_PARAM_CATABLE asks (encode.rs:2285-2333) go through the restated BroCatli splice
    (tests/brocatli_ref.py, src/concat/mod.rs) and decode to the concatenated inputs; a plain stream is refused by it."""
    import brocatli_ref as bc
    a, b = golden_bytes("alice29.txt")[:40000], golden_bytes("asyoulik.txt")[:30000]
    sa, sb, se, s1 = (_catable_from_model(model, a, q), _catable_from_model(model, b, q, byte_align=True),
                      _catable_from_model(model, b"", q), _catable_from_model(model, b"z", q))
    for s, d in ((sa, a), (sb, b), (se, b""), (s1, b"z")):
        assert sys_decompress(s, max(1, len(d))) == d
    whole = bc.concat([sa, sb, se, s1, sa])
    assert sys_decompress(whole, 2 * len(a) + len(b) + 1) == a + b + b"z" + a
    with pytest.raises(bc.NotCraftedForConcatenation):
        bc.concat([sa, model.compress(b, q, 22)[0]])
    assert bc.window_bits(sa) == (22, 4)


@pytest.mark.parametrize("q", [10, 11])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 7, 8, 9, 39, 40, 41, 63, 64, 65, 511, 512, 513, 8191, 8192, 8193, 16383, 16385, 70001])
def test_model_hq_edge_sizes(model, q, n):
    """quality 10 / 11 around every size that switches something on: 8 bytes (first bucket match), 8 + 8 / 16 / 32 bytes (the
    long-prefix levels), the 512-byte warm-up, the 8 / 16 KiB parse units, the 64 KiB statistics window."""
    d = (golden_bytes("alice29.txt") * 2)[:n]
    c, _ = model.compress(d, q, 22)
    assert sys_decompress(c, max(n, 1)) == d
