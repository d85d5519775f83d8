import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_brotli_b200 as rb
from oracle.harness import sys_decompress
N = rb._native
d = open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "alice29.txt"), "rb").read()
q = int(sys.argv[1]); split = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
enc = rb.DeviceEncoder(0)
enc.set_option(N.OPT_HQ_SPLIT, split)
import hashlib
for r in range(reps):
    c = enc.compress(d, q, 22)
    try:
        ok = sys_decompress(c, len(d)) == d
    except Exception:
        ok = "invalid"
    print("rep", r, "q", q, "split", split, len(c), hashlib.sha256(c).hexdigest()[:12], ok, flush=True)
