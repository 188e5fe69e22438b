import os, sys
sys.path.insert(0, "/root/repo")
import torch
from rspt_b200 import packer as R
F = 4096
x = R.synth_ecg(0, F, bps=3, ch=12, ns=8192)
y = torch.empty_like(x)
for kb in (0, 60, 75, 110, 200):
    if kb: os.environ["RSPT_INV_SMEM_KB"] = str(kb)
    p = R.SignalPacker("xdelta_hzr", 3, 12, 8192, 3, max_batch_frames=F)
    out = p.alloc_output(F, sidecar=True)
    b = p.compress_batch(x, out=out)
    for _ in range(2):
        p.decompress_batch(b, out=y)
    p.set_stage_timing(True); p.stage_times(reset=True)
    for _ in range(4):
        p.decompress_batch(b, out=y)
    st = p.stage_times(reset=True)
    print(kb, {k: round(v[0] / max(v[1], 1), 3) for k, v in st.items() if k in ("parse", "decode", "inverse")}, bool(torch.equal(x, y)))
    p.close()
