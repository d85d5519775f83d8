"""GPU: device-resident q5 throughput on the bench workload for several lane counts (chunk size is a build-time constant)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_brotli_b200 as rb
from tools import datagen
N = rb._native
d = datagen.enwik_like(100_000_000)
enc = rb.DeviceEncoder(0)
L = rb.lib()
d_in = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
cap = L.b200_max_compressed_size(len(d)) + 4096
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
for lanes in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4,6").split(",")]:
    enc.set_option(N.OPT_LANES, lanes)
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
    print("lanes=%d out=%d %.3f ms %.1f MB/s" % (lanes, nout, ms, len(d) / 1e3 / ms), flush=True)
