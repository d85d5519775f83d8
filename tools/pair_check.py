"""GPU box: k_parse_pair<4> (four units per warp) must give the same stream as the default path; prints timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_brotli_b200 as rb
from tools import datagen
g = lambda n: open(os.path.join(ROOT, "tests", "golden", n), "rb").read()
enc = rb.DeviceEncoder(0)
cases = [(g("alice29.txt"), 5, 22), (g("alice29.txt"), 6, 18), (g("random_then_unicode"), 5, 22), (g("compressed_file"), 5, 20),
         (datagen.json_logs(6_000_000), 5, 22), (datagen.enwik_like(30_000_000), 5, 22),
         (g("alice29.txt")[:70000] + datagen.pcg_random(5000) + g("asyoulik.txt"), 6, 22), (g("quickfox_repeated"), 5, 22)]
ok = True
for d, q, w in cases:
    outs = []
    for upw in (2, 4, 0):
        enc.set_option(rb._native.OPT_PAIR_PARSE, upw)
        outs.append(enc.compress(d, q, w))
    same = outs[0] == outs[1] == outs[2]
    ok &= same
    print(len(d), q, w, len(outs[0]), "identical" if same else "DIFFERENT", flush=True)
print("ALL IDENTICAL" if ok else "MISMATCH")
