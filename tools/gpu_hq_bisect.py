import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_brotli_b200 as rb
from tools.model_harness import Model
from oracle.harness import sys_decompress
N = rb._native
d = open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "alice29.txt"), "rb").read()
m = Model()
def rt(c):
    try:
        return sys_decompress(c, len(d)) == d
    except Exception as e:
        return "invalid"
def run(tag, q, opts, kw):
    enc = rb.DeviceEncoder(0)
    for k, v in opts.items():
        enc.set_option(k, v)
    c = enc.compress(d, q, 22)
    ref = m.compress(d, q, 22, **kw)[0]
    same = c == ref
    first = next((i for i in range(min(len(c), len(ref))) if c[i] != ref[i]), None)
    print("%-40s gpu=%d model=%d roundtrip=%s same=%s firstdiff=%s" % (tag, len(c), len(ref), rt(c), same, first), flush=True)
    enc.close()
run("q5 unit16384", 5, {N.OPT_UNIT: 16384, N.OPT_MB_UNITS: 256}, {"unit": 16384, "mb_units": 256})
run("q9 unit16384", 9, {N.OPT_UNIT: 16384, N.OPT_MB_UNITS: 256}, {"unit": 16384, "mb_units": 256})
run("q10 nosplit nodict", 10, {N.OPT_HQ_SPLIT: 0, N.OPT_DICT: 0}, {"hq_split": 0, "use_dict": 0})
run("q10 nosplit nodict noctx", 10, {N.OPT_HQ_SPLIT: 0, N.OPT_DICT: 0, N.OPT_CTX_MODEL: 0, N.OPT_SPLIT: 0}, {"hq_split": 0, "use_dict": 0, "ctx_model": 0, "split": 0})
run("q10 nosplit", 10, {N.OPT_HQ_SPLIT: 0}, {"hq_split": 0})
run("q10 full", 10, {}, {})
run("q11 full", 11, {}, {})
run("q10 full noctx", 10, {N.OPT_CTX_MODEL: 0}, {"ctx_model": 0})
