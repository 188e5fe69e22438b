import os, sys
sys.path.insert(0, "/root/repo")
import torch
from rspt_b200 import packer as R
F = 2048
x = R.synth_ecg(0, F, bps=3, ch=12, ns=8192)
p = R.SignalPacker("xdelta_hzr", 3, 12, 8192, 3, max_batch_frames=F)
out = p.alloc_output(F, sidecar=True)
for _ in range(3):
    b = p.compress_batch(x, out=out)
torch.cuda.synchronize()
os.environ["RSPT_DBG_DUMP"] = "1"
b = p.compress_batch(x, out=out)
torch.cuda.synchronize()
