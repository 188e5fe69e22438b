#!/usr/bin/env python
"""Randomised differential test: GPU stream == oracle stream, byte for byte, over random shapes and
data regimes (dense / sparse / bursty / constant / random), for the lossless packers and hadamard; decode
with the encoder's index, with a rebuilt index and through verify.

    python tools/stress.py [cases] [seed]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402  (test infrastructure)
from rspt_b200 import packer as R  # noqa: E402


def make_frames(rng, bps, ch, ns, nfr):
    bits = 8 * bps
    regime = rng.integers(0, 7)
    t = np.arange(ns)[None, :, None]
    if regime == 0:      # smooth + small noise (dense LSB plane, sparse upper planes)
        x = (rng.integers(1, 1 << min(bits - 2, 20)) * np.sin(t / rng.uniform(5, 400) + rng.uniform(0, 6, (nfr, 1, ch)))
             + rng.normal(0, rng.uniform(0.5, 40), (nfr, ns, ch)))
    elif regime == 1:    # constant per channel
        x = np.broadcast_to(rng.integers(-(1 << (bits - 2)), 1 << (bits - 2), (nfr, 1, ch)), (nfr, ns, ch)).astype(np.float64)
    elif regime == 2:    # white noise, full range
        x = rng.integers(-(1 << (bits - 1)), 1 << (bits - 1), (nfr, ns, ch)).astype(np.float64)
    elif regime == 3:    # bursts on a flat line
        x = np.zeros((nfr, ns, ch))
        for f in range(nfr):
            for _ in range(int(rng.integers(1, 12))):
                a = int(rng.integers(0, ns))
                w = int(rng.integers(1, 80))
                x[f, a:a + w, rng.integers(0, ch)] += rng.integers(-(1 << (bits - 3)), 1 << (bits - 3))
    elif regime == 4:    # ramp (constant delta: long zero runs after xdelta)
        x = (np.arange(ns)[None, :, None] * rng.integers(-3, 4, (nfr, 1, ch))).astype(np.float64)
    elif regime == 5:    # steps
        x = np.repeat(rng.integers(-500, 500, (nfr, (ns + 63) // 64, ch)), 64, axis=1)[:, :ns].astype(np.float64)
    else:                # mostly zero with isolated spikes
        x = np.where(rng.random((nfr, ns, ch)) < rng.uniform(0.001, 0.2), rng.integers(-200, 200, (nfr, ns, ch)), 0).astype(np.float64)
    lim = (1 << (bits - 1)) - 1
    xi = np.clip(np.rint(x), -lim - 1, lim).astype(np.int64)
    raw = np.ascontiguousarray((xi & ((1 << bits) - 1)).astype("<u4")).view(np.uint8).reshape(nfr, ns, ch, 4)[..., :bps]
    return np.ascontiguousarray(raw).reshape(nfr, -1), regime


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    done = 0
    for case in range(cases):
        kind = ["xdelta_hzr", "hzr", "xdelta_hzr", "hadamard"][int(rng.integers(0, 4))]
        bps = int(rng.integers(1, 5))
        ch = int(rng.integers(1, 6))
        if kind == "hadamard":
            ns = 1 << int(rng.integers(3, 14))
        else:
            ns = int(rng.choice([int(rng.integers(1, 300)), int(rng.integers(300, 9000)), int(rng.integers(9000, 40000)),
                                 65536 // ch, 65536 // ch + 1, 4096, 8192]))
        nb = int(rng.integers(1, bps + 1)) if kind == "xdelta_hzr" else 3
        nfr = int(rng.integers(1, 5))
        raws, regime = make_frames(rng, bps, ch, ns, nfr)
        o = O.OraclePacker(kind, bps, ch, ns, nb)
        p = R.SignalPacker(kind, bps, ch, ns, nb, max_batch_frames=nfr)
        dev = torch.from_numpy(raws.reshape(-1).copy()).cuda()
        b = p.compress_batch(dev)
        torch.cuda.synchronize()
        offs = b.offsets.cpu().numpy()
        stream = b.stream.cpu().numpy()
        tag = (case, kind, bps, ch, ns, nb, nfr, int(regime))
        want_dec = []
        for i in range(nfr):
            want = o.compress(raws[i])
            got = stream[offs[i]:offs[i + 1]].tobytes()
            assert got == want, ("stream", tag, i, len(got), len(want))
            want_dec.append(np.frombuffer(o.decompress(want)[0], np.uint8))
        want_dec = np.stack(want_dec)
        fnb = b.frame_nb.cpu().numpy()  # plane count used per frame: sticky, i.e. a running maximum
        assert fnb[-1] == o.nb and np.all(np.diff(fnb.astype(int)) >= 0), ("nb", tag, fnb, o.nb)
        for use_sc in (True, False):
            st = torch.zeros(nfr, dtype=torch.int32, device="cuda")
            dec = p.decompress_batch(b, status=st, use_sidecar=use_sc)
            torch.cuda.synchronize()
            assert not st.cpu().numpy().any(), ("status", tag, use_sc)
            assert np.array_equal(dec.cpu().numpy().reshape(nfr, -1), want_dec), ("decode", tag, use_sc)
        st = p.verify_batch(b)
        torch.cuda.synchronize()
        assert not st.cpu().numpy().any(), ("verify", tag)
        p.close()
        done += 1
    print(f"stress ok: {done} cases, seed {seed}")


if __name__ == "__main__":
    main()
