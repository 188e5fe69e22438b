"""Small driver for ncu: a few compress (and optionally decompress) batches of one packer.

    python tools/prof_compress.py [kind] [frames] [--dec]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rspt_b200 import packer as R

kind = sys.argv[1] if len(sys.argv) > 1 else "xdelta_hzr"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 592
dec = "--dec" in sys.argv
shape = dict(bps=3, ch=12, ns=8192) if kind in ("xdelta_hzr", "hzr") else dict(bps=4, ch=12, ns=4096)
p = R.SignalPacker(kind, shape["bps"], shape["ch"], shape["ns"], 3, max_batch_frames=F)
x = R.synth_ecg(0, F, **shape)
out = p.alloc_output(F)
y = torch.empty_like(x)
for _ in range(3):
    b = p.compress_batch(x, out=out)
    if dec:
        p.decompress_batch(b, out=y)
torch.cuda.synchronize()
print("ok", kind, F, "CR", x.numel() / int(out.offsets[F].item()), "roundtrip", bool(torch.equal(x, y)) if dec else None)
