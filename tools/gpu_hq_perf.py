"""GPU: throughput and per-stage times of the quality 10 / 11 path (second call: the workspaces are allocated by the first)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_brotli_b200 as rb
from tools import datagen
from oracle.harness import sys_decompress
N = rb._native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
d = datagen.enwik_like(n) if kind == "text" else datagen.json_logs(n)
enc = rb.DeviceEncoder(0)
for q in (10, 11):
    enc.set_option(N.OPT_TIMING, 0)
    c = enc.compress(d, q, 22)
    t = time.time(); c = enc.compress(d, q, 22); dt = time.time() - t
    ok = sys_decompress(c, len(d)) == d
    enc.set_option(N.OPT_TIMING, 1)
    enc.compress(d, q, 22)
    st = enc.timings()[0]
    print("q%d %s n=%d out=%d rt=%s %.3fs %.1f MB/s (host buffers) stages(ms)=%s" % (
        q, kind, n, len(c), ok, dt, n / 1e6 / dt, {k: round(v, 1) for k, v in st.items()}), flush=True)
