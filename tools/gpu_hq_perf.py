"""GPU: throughput and per-stage times of the quality 10 / 11 path (A/B of the parse mapping)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_brotli_b200 as rb
from tools import datagen
from oracle.harness import sys_decompress
N = rb._native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
d = datagen.enwik_like(n)
enc = rb.DeviceEncoder(0)
ref = {}
for q in (10, 11):
    for tu in (0, 1):
        enc.set_option(N.OPT_HQ_THREAD_UNITS, tu)
        enc.compress(d[:2_000_000], q, 22)
        t = time.time(); c = enc.compress(d, q, 22); dt = time.time() - t
        ok = sys_decompress(c, len(d)) == d
        same = ref.setdefault(q, c) == c
        enc.set_option(N.OPT_TIMING, 1)
        enc.compress(d, q, 22)
        st = enc.timings()[0]
        enc.set_option(N.OPT_TIMING, 0)
        print("q%d thread_units=%d n=%d out=%d rt=%s same_as_first=%s %.3fs %.1f MB/s stages(ms)=%s" % (
            q, tu, n, len(c), ok, same, dt, n / 1e6 / dt, {k: round(v, 1) for k, v in st.items()}), flush=True)
