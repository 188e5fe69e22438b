"""Small, fixed workload for ncu captures: config-2 shape, 592 frames (4 CTAs per SM of the block kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rspt_b200 import packer as R
kind = sys.argv[1] if len(sys.argv) > 1 else "xdelta_hzr"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 592
sh = dict(bps=3, ch=12, ns=8192) if kind in ("xdelta_hzr", "hzr") else dict(bps=4, ch=12, ns=4096)
p = R.SignalPacker(kind, sh["bps"], sh["ch"], sh["ns"], 3, max_batch_frames=F)
x = R.synth_ecg(0, F, **sh)
out = p.alloc_output(F)
dec = torch.empty_like(x)
for _ in range(2):
    b = p.compress_batch(x, out=out)
    p.decompress_batch(b, out=dec)
torch.cuda.synchronize()
print("ok", kind, F, "CR", x.numel() / int(out.offsets[F].item()), "roundtrip", bool(torch.equal(x, dec)) if kind in ("xdelta_hzr", "hzr") else "lossy")
