import os, sys
sys.path.insert(0, "/root/repo")
import torch
from rspt_b200 import packer as R
F = 4096
shape = dict(bps=3, ch=12, ns=8192)
x = R.synth_ecg(0, F, **shape)
for skip in (0, 1, 2):
    os.environ["RSPT_DBG_SKIP"] = str(skip)
    p = R.SignalPacker("xdelta_hzr", 3, 12, 8192, 3, max_batch_frames=F)
    out = p.alloc_output(F, sidecar=True)
    for _ in range(2):
        b = p.compress_batch(x, out=out)
    p.set_stage_timing(True)
    p.stage_times(reset=True)
    for _ in range(4):
        b = p.compress_batch(x, out=out)
    st = p.stage_times(reset=True)
    ms = {k: v[0] / max(v[1], 1) for k, v in st.items()}
    print(skip, " ".join(f"{k}={v:.3f}" for k, v in ms.items() if k in ("transform","hist","tree","layout","encode")))
    p.close()
