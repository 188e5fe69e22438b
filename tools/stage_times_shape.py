"""Per-stage device times of one packer on one shape:  python tools/stage_times_shape.py kind bps ch ns [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rspt_b200 import packer as R
kind, bps, ch, ns = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
F = int(sys.argv[5]) if len(sys.argv) > 5 else 4096
p = R.SignalPacker(kind, bps, ch, ns, 3, max_batch_frames=F)
x = R.synth_ecg(0, F, bps=bps, ch=ch, ns=ns)
out = p.alloc_output(F)
y = torch.empty_like(x)
for _ in range(2):
    b = p.compress_batch(x, out=out)
    p.decompress_batch(b, out=y)
p.set_stage_timing(True)
p.stage_times(reset=True)
for _ in range(4):
    b = p.compress_batch(x, out=out)
    p.decompress_batch(b, out=y)
st = p.stage_times(reset=True)
raw = x.numel()
ms = {k: v[0] / max(v[1], 1) for k, v in st.items()}
comp = sum(ms[k] for k in ("transform", "hist", "tree", "layout", "encode"))
dec = sum(ms[k] for k in ("parse", "decode", "inverse"))
print(f"{kind} {bps}B x {ch} x {ns}, {F} frames, raw {raw/1e6:.1f} MB: compress {raw/comp/1e6:.1f} GB/s decompress {raw/dec/1e6:.1f} GB/s CR {raw/int(out.offsets[F].item()):.2f}  " +
      " ".join(f"{k}={v:.3f}" for k, v in ms.items()))
