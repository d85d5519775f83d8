"""GPU development check of the quality 10 / 11 path against the CPU model (run on the GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_brotli_b200 as rb
from tools.model_harness import Model
from tools import datagen
from oracle.harness import sys_decompress, sys_compress

def golden(n):
    return open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", n), "rb").read()

def main():
    hq_split = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    big = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    enc = rb.DeviceEncoder(0)
    enc.set_option(rb._native.OPT_HQ_SPLIT, hq_split)
    m = Model()
    cases = [("alice29", golden("alice29.txt")), ("rtu", golden("random_then_unicode")), ("qfr", golden("quickfox_repeated")),
             ("x", golden("x")), ("empty", b""), ("monkey", golden("monkey")), ("json1m", datagen.json_logs(1_000_000)),
             ("enwik1m", datagen.enwik_like(1_000_000)), ("pcg", datagen.pcg_random(300_000)), ("bw", golden("backward65536"))]
    bad = 0
    for name, d in cases:
        for q in (10, 11):
            t = time.time(); c = enc.compress(d, q, 22); dt = time.time() - t
            ok = sys_decompress(c, max(1, len(d))) == d
            ref = m.compress(d, q, 22, hq_split=hq_split)[0]
            same = c == ref
            print("%-8s q%d n=%d gpu=%d model=%d roundtrip=%s same=%s %.3fs" % (name, q, len(d), len(c), len(ref), ok, same, dt), flush=True)
            bad += (not ok) + (not same)
    if big:
        d = datagen.enwik_like(big)
        for q in (10, 11):
            enc.compress(d[:1000000], q, 22)
            t = time.time(); c = enc.compress(d, q, 22); dt = time.time() - t
            ok = sys_decompress(c, len(d)) == d
            print("big q%d n=%d gpu=%d roundtrip=%s %.3fs  %.1f MB/s" % (q, len(d), len(c), ok, dt, len(d) / 1e6 / dt), flush=True)
            enc.set_option(rb._native.OPT_TIMING, 1)
            enc.compress(d, q, 22)
            print("   stage ms", enc.timings())
            enc.set_option(rb._native.OPT_TIMING, 0)
            bad += not ok
    print("BAD" if bad else "ALL OK", bad)

main()
