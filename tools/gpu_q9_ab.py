"""GPU: q7..q9 with the deep buckets searched for every position (k_match_deep + k_parse) vs on demand where the parse stands
(k_rank_sig + k_parse_ondemand): identical streams, stage times with one lane, throughput with the 4-lane pipeline."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_brotli_b200 as rb
from tools import datagen
N = rb._native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
gold = lambda f: open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f), "rb").read()
inputs = {"text": datagen.enwik_like(n), "json": datagen.json_logs(n), "alice": gold("alice29.txt"), "rtu": gold("random_then_unicode"),
          "pcg": datagen.pcg_random(8_000_000), "tiled10k": datagen.tiled(gold("random_org_10k.bin"), 30_000_000)}
enc = rb.DeviceEncoder(0)
L = rb.lib()
for q, lgwin in ((9, 22), (7, 22), (8, 22), (9, 16), (5, 16)):
    for name, d in inputs.items():
        if (q, lgwin) != (9, 22) and name in ("pcg", "tiled10k"):
            continue
        d_in = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
        cap = L.b200_max_compressed_size(len(d)) + 4096
        d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        ref = None
        for od in (0, 2):  # 2 = on demand whatever the size (1, the default, keeps inputs below 4 MiB on the up-front path)
            enc.set_option(N.OPT_ONDEMAND, od)
            enc.set_option(N.OPT_TIMING, 0)
            enc.set_option(N.OPT_LANES, 4)
            nout = enc.compress_device(d_in.data_ptr(), len(d), d_out.data_ptr(), cap, q, lgwin)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(2):
                nout = enc.compress_device(d_in.data_ptr(), len(d), d_out.data_ptr(), cap, q, lgwin)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / 2
            c = bytes(d_out[:nout].cpu().numpy())
            if ref is None:
                ref = c
            enc.set_option(N.OPT_TIMING, 1)
            enc.set_option(N.OPT_LANES, 1)
            enc.compress_device(d_in.data_ptr(), len(d), d_out.data_ptr(), cap, q, lgwin)
            st = enc.timings()[0]
            print("q%d w%d %-8s n=%d ondemand=%d out=%d same=%s %.2f ms %.0f MB/s stages=%s" % (
                q, lgwin, name, len(d), od, nout, ref == c, ms, len(d) / 1e3 / ms, {k: round(x, 2) for k, x in st.items()}), flush=True)
