import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_brotli_b200 as rb
from oracle.harness import sys_decompress
name, q, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
d = open(os.path.join(ROOT, "tests", "golden", name), "rb").read()
enc = rb.DeviceEncoder(0)
c = enc.compress(d, q, w)
print(len(c), sys_decompress(c, max(len(d), 1)) == d)
