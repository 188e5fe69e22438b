import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch, json
from oracle import oracle as O
from rspt_b200 import packer as R
def to_dev(a): return torch.from_numpy(np.ascontiguousarray(a, np.uint8).reshape(-1).copy()).cuda()
# 1. dct decode
for direct in (0, 1):
    os.environ["RSPT_DCT_DIRECT"] = str(direct)
    bps, ch, ns = 4, 2, 512
    raws = O.synth_ecg(1, 2, bps, ch, ns)
    p = R.SignalPacker.new_dct(bps, ch, ns, max_batch_frames=2)
    d = to_dev(raws)
    batch = p.compress_batch(d)
    out = torch.full((2 * bps * ch * ns,), 0xAB, dtype=torch.uint8, device="cuda")
    p.decompress_batch(batch, out=out)
    torch.cuda.synchronize()
    dec = out.cpu().numpy().reshape(2, -1)
    o = O.OraclePacker("dct", bps, ch, ns)
    for i in range(2):
        want = o.compress(raws[i]); wd = np.frombuffer(o.decompress(want)[0], np.uint8)
        a = wd.view('<i4').reshape(ns, ch); b = dec[i].view('<i4').reshape(ns, ch)
        diff = (a.astype(np.int64) - b)
        print("direct", direct, "frame", i, "nonzero diffs", (diff != 0).sum(), "max", np.abs(diff).max(), "first rows", a[:3].tolist(), b[:3].tolist(), "raw", raws[i].view('<i4').reshape(ns,ch)[:3].tolist())
os.environ.pop("RSPT_DCT_DIRECT")
# 2. escalation
rng = np.random.default_rng(3)
def make_raw(rng, bps, ch, ns, amp):
    x = rng.integers(-amp, amp, (ns, ch)); x = np.clip(x, -(1 << (8*bps-1)), (1 << (8*bps-1))-1).astype(np.int32)
    return x.astype("<i4").view(np.uint8).reshape(ns, ch, 4)[:, :, :bps].copy().reshape(-1)
bps, ch, ns = 4, 3, 500
amps = [20, 20, 3000, 20, 20, 900000, 20, 20]
raws = np.stack([make_raw(rng, bps, ch, ns, a) for a in amps])
p = R.SignalPacker.new_xdelta_hzr(bps, ch, ns, 1, max_batch_frames=8)
o = O.OraclePacker("xdelta_hzr", bps, ch, ns, 1)
want_nb = []
for r in raws:
    o.compress(r); want_nb.append(o.nb)
batch = p.compress_batch(to_dev(raws)); torch.cuda.synchronize()
print("want_nb", want_nb, "got", batch.frame_nb.cpu().numpy().tolist(), "state", p.nb)
# 3. golden sine
g = json.load(open("/root/repo/tests/golden/golden.json"))
z = np.load("/root/repo/tests/golden/fixtures.npz")
for case in g["cases"][:8]:
    kind, bps, ch, ns = case["kind"], case["bps"], case["ch"], case["ns"]
    fb = bps*ch*ns
    if kind == "dct": os.environ["RSPT_DCT_DIRECT"] = "1"
    p = R.SignalPacker(kind, bps, ch, ns, case["nb"] or 3, max_batch_frames=1)
    os.environ.pop("RSPT_DCT_DIRECT", None)
    batch = p.compress_batch(to_dev(z[case["input"]][:fb])); torch.cuda.synchronize()
    offs = batch.offsets.cpu().numpy()
    print(case["name"], "got", int(offs[1]), "want", case["frames"][0]["len"], "nb", p.nb, "want nb", case["final_nb"], batch.frame_nb.cpu().numpy().tolist())
