"""Small inputs through every kernel family added in round 2 (run under compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_brotli_b200 as rb
from tools import datagen
from oracle.harness import sys_decompress
N = rb._native
enc = rb.DeviceEncoder(0)
g = lambda f: open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f), "rb").read()
d = g("alice29.txt") + datagen.json_logs(200_000) + g("random_then_unicode")[:50_000]
enc.set_option(N.OPT_ONDEMAND, 2)
for q, w in ((9, 22), (8, 22), (7, 22), (9, 16), (5, 16), (5, 22), (10, 22), (11, 22)):
    c = enc.compress(d, q, w)
    assert sys_decompress(c, len(d)) == d
    print("q%d w%d ok %d" % (q, w, len(c)), flush=True)
import io
for kw in (dict(catable=True, magic_number=True), dict(appendable=True, byte_align=True), dict(catable=True, bare_stream=True)):
    wbuf = io.BytesIO()
    rb.BrotliCompress(io.BytesIO(d[:100_000]), wbuf, rb.BrotliEncoderParams(quality=5, lgwin=22, **kw))
    print("framing", kw, len(wbuf.getvalue()), flush=True)
