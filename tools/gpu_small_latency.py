"""GPU: one-shot latency of small inputs through the C ABI (host buffers)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_brotli_b200 as rb
g = lambda f: open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f), "rb").read()
enc = rb.DeviceEncoder(0)
for name in ("alice29.txt", "random_then_unicode", "quickfox_repeated", "x"):
    d = g(name)
    for q in (5, 9, 11):
        for _ in range(3):
            enc.compress(d, q, 22)
        t = time.perf_counter()
        for _ in range(20):
            c = enc.compress(d, q, 22)
        dt = (time.perf_counter() - t) / 20
        print("%-20s n=%-7d q%-2d %.3f ms  %.1f MB/s  out=%d" % (name, len(d), q, dt * 1e3, len(d) / 1e6 / dt, len(c)), flush=True)
