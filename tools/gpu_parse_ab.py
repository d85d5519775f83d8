"""GPU: A/B of the q5 parse mapping (units per warp: 4 = quarter-warps, 32 = one unit per thread): identical streams, stage times
with the chunks serialised on one lane, and device-resident throughput with the 4-lane pipeline."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_brotli_b200 as rb
from tools import datagen
N = rb._native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 32]
inputs = {"text": datagen.enwik_like(n), "json": datagen.json_logs(min(n, 50_000_000)), "pcg": datagen.pcg_random(20_000_000)}
enc = rb.DeviceEncoder(0)
L = rb.lib()
ref = {}
for name, d in inputs.items():
    d_in = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
    cap = L.b200_max_compressed_size(len(d)) + 4096
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    for v in variants:
        enc.set_option(N.OPT_PAIR_PARSE, v)
        enc.set_option(N.OPT_TIMING, 0)
        enc.set_option(N.OPT_LANES, 4)
        for _ in range(3):
            nout = enc.compress_device(d_in.data_ptr(), len(d), d_out.data_ptr(), cap, 5, 22)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(5):
            nout = enc.compress_device(d_in.data_ptr(), len(d), d_out.data_ptr(), cap, 5, 22)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 5
        c = bytes(d_out[:nout].cpu().numpy())
        same = ref.setdefault(name, c) == c
        enc.set_option(N.OPT_TIMING, 1)
        enc.set_option(N.OPT_LANES, 1)
        enc.compress_device(d_in.data_ptr(), len(d), d_out.data_ptr(), cap, 5, 22)
        st = enc.timings()[0]
        print("%s n=%d parse_variant=%d out=%d same_as_first=%s %.3f ms %.1f MB/s stages(ms)=%s" % (
            name, len(d), v, nout, same, ms, len(d) / 1e3 / ms, {k: round(x, 2) for k, x in st.items()}), flush=True)
