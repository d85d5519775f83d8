"""Large-input checks on the GPU box: multi-chunk streams, BASELINE configs 3/4-like inputs, timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_brotli_b200 as rb
from oracle.harness import sys_decompress, sys_compress
from tools import datagen

def check(name, d, q, w, enc, ref=False):
    t = time.perf_counter(); c = enc.compress(d, q, w); dt = time.perf_counter() - t
    t = time.perf_counter(); ok = sys_decompress(c, len(d)) == d; dd = time.perf_counter() - t
    msg = "%s n=%d q%d w%d -> %d (ratio %.5f) RT %s  gpu e2e %.1f ms (%.2f GB/s), cpu decode %.1fs" % (
        name, len(d), q, w, len(c), len(c) / len(d), ok, dt * 1e3, len(d) / dt / 1e9, dd)
    if ref:
        r = len(sys_compress(d, q, w)); msg += "  libbrotlienc %d (%+.3f%%)" % (r, 100.0 * (len(c) - r) / r)
    print(msg, flush=True)
    return ok

def main():
    enc = rb.DeviceEncoder(0)
    ok = True
    text = datagen.enwik_like(100_000_000)
    big = text * 3 + text[:50_000_000]          # 350 MB: 3 chunks of 128 MiB
    ok &= check("text350MB", big, 5, 22, enc)
    blk = open(os.path.join(ROOT, "tests", "golden", "random_org_10k.bin"), "rb").read()
    ok &= check("cfg3 random10k x 1e5 (1 GB)", datagen.tiled(blk, 1_000_000_000), 5, 22, enc)
    ok &= check("cfg3b pcg random 256MB", datagen.pcg_random(256_000_000), 5, 22, enc)
    js = datagen.json_logs(64_000_000)
    ok &= check("json 64MB q9", js, 9, 22, enc, ref=True)
    ok &= check("json 64MB q5", js, 5, 22, enc, ref=True)
    ok &= check("text 100MB q9", text, 9, 22, enc, ref=False)
    qf = open(os.path.join(ROOT, "tests", "golden", "quickfox_repeated"), "rb").read()
    ok &= check("cfg5-shaped quickfox x 512MiB (run as q9 device path)", datagen.tiled(qf, 512 << 20), 11, 24, enc)
    c = rb.compress_multi(rb.BrotliEncoderParams(quality=9, lgwin=22), js, 8)
    print("compress_multi json 8 shards q9:", len(c), sys_decompress(c, len(js)) == js)
    print("ALL OK" if ok else "FAILURES")

if __name__ == "__main__":
    main()
