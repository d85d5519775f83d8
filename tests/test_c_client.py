"""A C client (tests/c_client/client.c: the calls of the reference's c/multiexample.c:51-146 and of c/brotli/encode.h)
is compiled against include/brotli_b200.h and linked with -lbrotli_b200.  CPU: it must build, link and report "no device"
through every entry point; GPU: every stream it writes must decode to the input through libbrotlidec."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "rust-brotli_b200")
SRC = os.path.join(ROOT, "tests", "c_client", "client.c")


def _build(tmp_path):
    exe = str(tmp_path / "client")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                           "-L", LIBDIR, "-lbrotli_b200", "-Wl,-rpath," + LIBDIR])
    return exe


def _has_gpu():
    L = ctypes.CDLL(os.path.join(LIBDIR, "libbrotli_b200.so"))
    L.b200_device_count.restype = ctypes.c_int
    return L.b200_device_count() > 0


@pytest.mark.skipif(_has_gpu(), reason="CPU-side check")
def test_c_client_builds_links_and_fails_loudly_without_gpu(tmp_path):
    exe = _build(tmp_path)
    inp = tmp_path / "in.bin"
    inp.write_bytes(b"Mary had a little lamb. Its fleece was white as snow.\n" * 50)
    r = subprocess.run([exe, str(inp), str(tmp_path / "out")])
    assert r.returncode == 77  # no device: NULL pool, NULL instance, BROTLI_FALSE -- and no output file
    assert not any(p.name.startswith("out.") for p in tmp_path.iterdir())


@pytest.mark.gpu
def test_c_client_streams_round_trip(tmp_path):
    from oracle.harness import sys_decompress
    exe = _build(tmp_path)
    data = open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read() * 3
    inp = tmp_path / "in.bin"
    inp.write_bytes(data)
    r = subprocess.run([exe, str(inp), str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for ext in ("pool", "multi", "oneshot", "stream", "streaming"):
        comp = (tmp_path / ("out." + ext)).read_bytes()
        assert sys_decompress(comp, len(data)) == data, ext
        assert len(comp) < len(data) * 0.4
