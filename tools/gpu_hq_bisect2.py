import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_brotli_b200 as rb
from tools.model_harness import Model
from tools import datagen
from oracle.harness import sys_decompress
N = rb._native
m = Model()
enc = rb.DeviceEncoder(0)
def run(tag, d, q, opts, kw):
    for k, v in opts.items():
        enc.set_option(k, v)
    c = enc.compress(d, q, 22)
    for k in opts:
        enc.set_option(k, 1)
    ref = m.compress(d, q, 22, **kw)[0]
    try:
        ok = sys_decompress(c, len(d)) == d
    except Exception:
        ok = "invalid"
    first = next((i for i in range(min(len(c), len(ref))) if c[i] != ref[i]), None)
    print("%-34s n=%d gpu=%d model=%d rt=%s same=%s firstdiff=%s" % (tag, len(d), len(c), len(ref), ok, c == ref, first), flush=True)
# TMA staging sanity on the q5 / q9 paths
a = open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "alice29.txt"), "rb").read()
run("q5 alice", a, 5, {}, {})
run("q9 alice", a, 9, {}, {})
e = datagen.enwik_like(6_000_000)
run("q5 enwik6m", e, 5, {}, {})
for n in (1_000_000, 3_000_000, 4_194_304, 4_300_000, 6_000_000):
    run("q10 nosplit", e[:n], 10, {N.OPT_HQ_SPLIT: 0}, {"hq_split": 0})
    run("q10 full", e[:n], 10, {}, {})
    run("q10 full noctx", e[:n], 10, {N.OPT_CTX_MODEL: 0}, {"ctx_model": 0})
