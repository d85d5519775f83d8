"""GPU development: compare the quality >= 10 stages (matches, per-unit parse) of the device with the CPU model."""
import sys, os, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_brotli_b200 as rb
from tools.model_harness import Model
from tools import datagen
HQ_MAXM = 16
def golden(n):
    return open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", n), "rb").read()
def main():
    q = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    d = golden("alice29.txt") if len(sys.argv) < 3 else datagen.enwik_like(int(sys.argv[2]))
    n = len(d)
    unit = 16384
    nu = (n + unit - 1) // unit
    cu = unit // 2 + 1
    def bufs():
        return (np.zeros(n, np.uint8), np.zeros(n * HQ_MAXM * 2, np.uint32), np.zeros(nu * 3, np.uint32), np.zeros(nu * cu * 3, np.uint32))
    m = Model()
    mh = bufs()
    m.lib.gpu_model_debug_hq(*[ctypes.c_void_p(a.ctypes.data) for a in mh])
    ref = m.compress(d, q, 22, hq_split=0)[0]
    m.lib.gpu_model_debug_hq(None, None, None, None)
    enc = rb.DeviceEncoder(0)
    enc.set_option(rb._native.OPT_HQ_SPLIT, 0)
    L = rb.lib()
    gh = bufs()
    L.b200_stage_hq.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t] + [ctypes.c_void_p] * 4
    ok = L.b200_stage_hq(enc._h, q, 22, d, n, *[ctypes.c_void_p(a.ctypes.data) for a in gh])
    print("stage_hq ok", ok)
    names = ["hqn", "hqm", "units", "raw"]
    for nm, a, b in zip(names, mh, gh):
        if nm == "hqm":  # only the valid entries
            a = a.reshape(n, HQ_MAXM, 2); b = b.reshape(n, HQ_MAXM, 2)
            mask = np.arange(HQ_MAXM)[None, :] < mh[0][:, None]
            bad = np.nonzero(((a != b).any(axis=2)) & mask)
            print(nm, "mismatching (pos, k):", len(bad[0]), list(zip(bad[0][:5], bad[1][:5])))
            for p, k in list(zip(bad[0][:5], bad[1][:5])):
                print("   pos", p, "k", k, "model", a[p, k], "gpu", b[p, k], "nm", mh[0][p], gh[0][p])
        elif nm == "raw":
            a = a.reshape(nu, cu, 3); b = b.reshape(nu, cu, 3)
            for u in range(nu):
                k = int(mh[2][u])
                if not np.array_equal(a[u, :k], b[u, :k]):
                    w = np.nonzero((a[u, :k] != b[u, :k]).any(axis=1))[0]
                    print("raw unit", u, "first diff cmd", w[:3], a[u, w[0]], b[u, w[0]])
                    break
            else:
                print("raw equal")
        else:
            diff = np.nonzero(a != b)[0]
            print(nm, "mismatches:", len(diff), diff[:10], a[diff[:5]], b[diff[:5]])
    c = enc.compress(d, q, 22)
    print("stream equal to model:", c == ref, len(c), len(ref))
main()
