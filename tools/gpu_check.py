"""First-light check on a GPU box: GPU stream == CPU model stream, decodes bit-exact with libbrotlidec, stage timings."""
import importlib.util, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle.harness import sys_compress, sys_decompress
from tools.model_harness import Model
from tools import datagen

spec = importlib.util.spec_from_file_location("bnative", os.path.join(ROOT, "rust-brotli_b200", "_native.py"))
bn = importlib.util.module_from_spec(spec); spec.loader.exec_module(bn)

def main():
    sizes = [int(x) for x in sys.argv[1:]] or [0, 1, 5, 100, 5000, 70000, 300000, 2_000_000, 20_000_000]
    enc = bn.DeviceEncoder(0)
    enc.set_option(bn.OPT_TIMING, 1)
    model = Model()
    text = datagen.enwik_like(max(max(sizes), 1000))
    rnd = datagen.pcg_random(300000)
    cases = [("text", text[:n]) for n in sizes]
    cases += [("random", rnd), ("tiled", datagen.tiled(rnd[:10000], 3_000_000)), ("zeros", bytes(1_000_000)),
              ("mixed", rnd[:100000] + text[:200000] + bytes(50000) + text[100000:400000])]
    bad = 0
    for name, d in cases:
        for q, w in ((5, 22), (9, 22), (7, 18)) if len(d) <= 2_000_000 else ((5, 22),):
            t = time.time()
            try:
                c = enc.compress(d, q, w)
            except Exception as e:
                print(name, len(d), q, w, "GPU FAIL", e); bad += 1; continue
            dt = time.time() - t
            try:
                ok = sys_decompress(c, len(d)) == d
            except Exception as e:
                ok = "REJECT"
            mc, _ = model.compress(d, q, w)
            same = (mc == c)
            ref = len(sys_compress(d, q, w))
            tm, launches = enc.timings()
            print(name, len(d), "q%d w%d" % (q, w), "gpu", len(c), "model", len(mc), "sys", ref, "RT", ok, "==model", same,
                  "wall %.1f ms" % (dt * 1e3), {k: round(v, 3) for k, v in tm.items()}, "launches", launches, flush=True)
            if ok is not True or not same:
                bad += 1
                if not same:
                    m = next((i for i in range(min(len(c), len(mc))) if c[i] != mc[i]), None)
                    print("   first differing byte", m, "of", len(c), len(mc))
    print("FAILURES", bad)
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
