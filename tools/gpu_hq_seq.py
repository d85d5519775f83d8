import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_brotli_b200 as rb
from tools.model_harness import Model
from tools import datagen
from oracle.harness import sys_decompress
m = Model()
enc = rb.DeviceEncoder(0)
g = lambda n: open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", n), "rb").read()
e = datagen.enwik_like(6_000_000)
ref = m.compress(e, 10, 22)[0]
def check(tag):
    c = enc.compress(e, 10, 22)
    print(tag, "same", c == ref, len(c), len(ref), flush=True)
check("fresh")
d = datagen.tiled(g("quickfox_repeated"), 512 << 20)
c = enc.compress(d, 11, 24)
print("config5", len(c), hashlib.sha256(sys_decompress(c, len(d))).digest() == hashlib.sha256(d).digest(), flush=True)
check("after config5")
j = datagen.json_logs(30_000_000)
c = enc.compress(j, 9, 22)
check("after json q9 30MB")
c = enc.compress(datagen.enwik_like(60_000_000, seed=3), 10, 22)
check("after enwik q10 60MB")
check("again")
