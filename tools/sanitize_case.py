"""compute-sanitizer driver: a few small compressions that touch every kernel variant (shallow / deep match, fast parse for
4 / 10 / 16 cache candidates, dictionary, long inserts, raw metablocks, two chunks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_brotli_b200 as rb
from oracle.harness import sys_decompress
from tools import datagen
g = lambda n: open(os.path.join(ROOT, "tests", "golden", n), "rb").read()
enc = rb.DeviceEncoder(0)
text = g("alice29.txt")
mixed = text[:60000] + datagen.pcg_random(3000) + text[60000:90000] + datagen.pcg_random(70000) + g("random_then_unicode")[:50000]
cases = [(text, 5, 22), (text, 6, 18), (text, 7, 22), (text, 9, 22), (text, 5, 16), (mixed, 5, 22), (mixed, 9, 20),
         (datagen.pcg_random(300000), 5, 22), (bytes(200000), 5, 22)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    cases = [(datagen.enwik_like(26_000_000), 5, 22)]
for d, q, w in cases:
    c = enc.compress(d, q, w)
    print(len(d), q, w, len(c), sys_decompress(c, len(d)) == d, flush=True)
