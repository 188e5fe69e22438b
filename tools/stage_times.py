"""Per-stage device times (CUDA events on the handle's stream) of every packer kind.
    python tools/stage_times.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rspt_b200 import packer as R

F = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for kind, shape in (("xdelta_hzr", dict(bps=3, ch=12, ns=8192)), ("hzr", dict(bps=3, ch=12, ns=8192)),
                    ("hadamard", dict(bps=4, ch=12, ns=4096)), ("dct", dict(bps=4, ch=12, ns=4096))):
    p = R.SignalPacker(kind, shape["bps"], shape["ch"], shape["ns"], 3, max_batch_frames=F)
    x = R.synth_ecg(0, F, **shape)
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
    print(f"{kind:10s} raw {raw/1e6:7.1f} MB  compress {raw/comp/1e6:7.1f} GB/s  decompress {raw/dec/1e6:7.1f} GB/s  CR {raw/int(out.offsets[F].item()):.2f}  " +
          " ".join(f"{k}={v:.3f}" for k, v in ms.items()))
    p.close()
