"""Does running k handles side by side (each on its own stream, each with 1/k of the frames) beat one handle?
Upper bound for what an internal split of a batch over streams could give."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rspt_b200 import packer as R
sh = dict(bps=3, ch=12, ns=8192)
fb = 3 * 12 * 8192
F = 4096
x = [R.synth_ecg(i * F, F, **sh) for i in range(4)]
for k in (1, 2, 3, 4):
    Fk = F // k
    streams = [torch.cuda.Stream() for _ in range(k)]
    hs, outs, decs = [], [], []
    for s in streams:
        with torch.cuda.stream(s):
            hs.append(R.SignalPacker.new_xdelta_hzr(3, 12, 8192, 3, max_batch_frames=Fk))
            outs.append([hs[-1].alloc_output(Fk, sidecar=True) for _ in range(2)])
            decs.append(torch.empty(Fk * fb, dtype=torch.uint8, device="cuda"))
    def comp(i):
        r = []
        for j, s in enumerate(streams):
            with torch.cuda.stream(s):
                r.append(hs[j].compress_batch(x[i % 4][j * Fk * fb:(j + 1) * Fk * fb], out=outs[j][i & 1]))
        return r
    def dec(bs):
        for j, s in enumerate(streams):
            with torch.cuda.stream(s):
                hs[j].decompress_batch(bs[j], out=decs[j])
    for i in range(3):
        bs = comp(i)
    dec(bs)
    torch.cuda.synchronize()
    n = 24
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for i in range(n):
        bs = comp(i)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    tc = e0.elapsed_time(e1) / n
    e0.record()
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for i in range(n):
        dec(bs)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    td = e0.elapsed_time(e1) / n
    print(f"{k} handle(s) x {Fk} frames: compress {tc:.3f} ms = {k * Fk * fb / tc / 1e6:.0f} GB/s, decompress {td:.3f} ms = {k * Fk * fb / td / 1e6:.0f} GB/s", flush=True)
    for h in hs: h.close()
    del hs, outs, decs
    torch.cuda.empty_cache()
