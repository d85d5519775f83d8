"""GPU: q5 throughput on the bench workload with 4 KiB (shipped) and 2 KiB parse units."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_brotli_b200 as rb
from tools import datagen
from oracle.harness import sys_decompress
N = rb._native
d = datagen.enwik_like(100_000_000)
enc = rb.DeviceEncoder(0)
L = rb.lib()
d_in = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
cap = L.b200_max_compressed_size(len(d)) + 4096
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
for unit in (4096, 2048, 4096, 2048):
    enc.set_option(N.OPT_UNIT, unit)
    enc.set_option(N.OPT_MB_UNITS, (4 << 20) // unit)
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
    ok = sys_decompress(bytes(d_out[:nout].cpu().numpy()), len(d)) == d
    print("unit=%d out=%d rt=%s %.3f ms %.1f MB/s" % (unit, nout, ok, ms, len(d) / 1e3 / ms), flush=True)
