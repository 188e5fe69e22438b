"""Generate tests/golden/{fixtures.npz,golden.json} by running the UNMODIFIED reference
(oracle/_ref/libref.so, built from /root/reference by oracle/Makefile) in the build container.

    python tests/golden/make_golden.py

Inputs: the README sine (rspt_test.cpp:231-241), the reference's two ECG fixtures (unpacked from
lib_rspt_test/*.7z; the 12-channel one is cut to its first 16384 samples to keep the repo small),
and frames from this repo's synthetic generator.  For every case the reference's compressed
frame is recorded as (length, crc32, sha256), its decoded output as sha256, plus the final plane
count and, for lossy packers, PRDN per rspt_test.cpp:98-111.  Nothing under /root/reference is
read at test time.
"""
import hashlib
import json
import math
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
from extract_fixtures import unpack_7z_single_lzma2  # noqa: E402
from oracle import oracle as O  # noqa: E402


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    O.build()
    ds = np.frombuffer(unpack_7z_single_lzma2("/root/reference/lib_rspt_test/data_stream.7z"), np.uint8)
    e12 = np.frombuffer(unpack_7z_single_lzma2(
        "/root/reference/lib_rspt_test/12_chan_32bit_34199_samples_r00000135fghd8.raw.7z"), np.uint8)
    assert zlib.crc32(ds) == 0x4E9DAAF1 and zlib.crc32(e12) == 0x211C9CA6
    e12_16k = e12[: 4 * 12 * 16384].copy()
    sine = np.array([int(math.sin(i / 100.0) * 1000.0) for i in range(16384)], np.int32).view(np.uint8)
    inputs = {"data_stream": ds, "ecg12_16384": e12_16k, "sine": sine}
    synth_cfgs = {
        "synth_A": dict(first=0, n=2, bps=3, ch=12, ns=8192),
        "synth_B": dict(first=5, n=2, bps=4, ch=12, ns=4096),
        "synth_C16": dict(first=1000000, n=3, bps=2, ch=5, ns=1000),
    }
    synth = {}
    for k, c in synth_cfgs.items():
        synth[k] = O.synth_ecg(c["first"], c["n"], c["bps"], c["ch"], c["ns"]).reshape(-1)

    cases = [
        # (name, input, kind, bps, ch, ns, nb)
        ("sine_xdelta_nb3", "sine", "xdelta_hzr", 4, 1, 8192, 3),
        ("sine_xdelta_nb1", "sine", "xdelta_hzr", 4, 1, 8192, 1),
        ("sine_xdelta_nb2", "sine", "xdelta_hzr", 4, 1, 8192, 2),
        ("sine_xdelta_nb4", "sine", "xdelta_hzr", 4, 1, 8192, 4),
        ("sine_hzr", "sine", "hzr", 4, 1, 8192, 0),
        ("sine_hadamard", "sine", "hadamard", 4, 1, 8192, 0),
        ("sine_dct", "sine", "dct", 4, 1, 4096, 0),
        ("sine16k_xdelta", "sine", "xdelta_hzr", 4, 1, 16384, 3),
        ("ds_xdelta", "data_stream", "xdelta_hzr", 3, 3, 20000, 3),
        ("ds_hzr", "data_stream", "hzr", 3, 3, 20000, 0),
        ("ds_hadamard", "data_stream", "hadamard", 3, 3, 16384, 0),
        ("ds_dct", "data_stream", "dct", 3, 3, 4096, 0),
        ("ecg12_xdelta_8192", "ecg12_16384", "xdelta_hzr", 4, 12, 8192, 3),
        ("ecg12_xdelta_16384", "ecg12_16384", "xdelta_hzr", 4, 12, 16384, 3),
        ("ecg12_xdelta_16384_nb1", "ecg12_16384", "xdelta_hzr", 4, 12, 16384, 1),
        ("ecg12_hzr_16384", "ecg12_16384", "hzr", 4, 12, 16384, 0),
        ("ecg12_hadamard_16384", "ecg12_16384", "hadamard", 4, 12, 16384, 0),
        ("ecg12_hadamard_4096", "ecg12_16384", "hadamard", 4, 12, 4096, 0),
        ("ecg12_dct_4096", "ecg12_16384", "dct", 4, 12, 4096, 0),
        ("ecg12_dct_512", "ecg12_16384", "dct", 4, 12, 512, 0),
    ]
    for k, c in synth_cfgs.items():
        for kind in ("xdelta_hzr", "hzr", "hadamard"):
            if kind == "hadamard" and c["ns"] & (c["ns"] - 1):
                continue
            cases.append((f"{k}_{kind}", k, kind, c["bps"], c["ch"], c["ns"], 3))
    cases.append(("synth_B_dct", "synth_B", "dct", 4, 12, 4096, 0))

    all_inputs = dict(inputs)
    all_inputs.update(synth)
    golden = {"inputs": {k: {"bytes": int(v.size), "sha256": sha(v)} for k, v in all_inputs.items()},
              "synth": synth_cfgs, "cases": []}
    for name, key, kind, bps, ch, ns, nb in cases:
        data = all_inputs[key]
        fb = bps * ch * ns
        nframes = data.size // fb if key.startswith("synth") else 1
        r = O.RefPacker(kind, bps, ch, ns, nb)
        frames = []
        for f in range(nframes):
            src = data[f * fb:(f + 1) * fb]
            comp = r.compress(src)
            dec, used = r.decompress(comp)
            assert used == len(comp)
            rec = {"len": len(comp), "crc32": "%08x" % zlib.crc32(comp), "sha256": sha(comp),
                   "dec_sha256": sha(dec)}
            if kind in ("hadamard", "dct"):
                rec["prdn"] = O.prdn(src, dec, bps, ch, ns)
            else:
                assert dec == src.tobytes()
            frames.append(rec)
        # the plane count the instance ended with: probe via the oracle port (pinned to the
        # reference by the byte comparison below)
        o = O.OraclePacker(kind, bps, ch, ns, nb)
        for f in range(nframes):
            assert o.compress(data[f * fb:(f + 1) * fb]) is not None
        golden["cases"].append({"name": name, "input": key, "kind": kind, "bps": bps, "ch": ch,
                                "ns": ns, "nb": nb, "final_nb": o.nb, "frames": frames})
        print(name, [fr["len"] for fr in frames], "final_nb", o.nb)
    np.savez_compressed(os.path.join(HERE, "fixtures.npz"), **inputs)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1)
    # the 34199-sample answers quoted in BASELINE.md section 3 (input too large to commit)
    full = []
    for kind, nb in (("xdelta_hzr", 3), ("xdelta_hzr", 1), ("hzr", 0)):
        r = O.RefPacker(kind, 4, 12, 34199, nb)
        comp = r.compress(e12)
        full.append({"kind": kind, "nb": nb, "len": len(comp), "crc32": "%08x" % zlib.crc32(comp)})
    with open(os.path.join(HERE, "golden_full_fixture.json"), "w") as f:
        json.dump({"input_crc32": "211c9ca6", "cases": full}, f, indent=1)
    print("wrote", os.path.getsize(os.path.join(HERE, "fixtures.npz")), "byte fixtures.npz")


if __name__ == "__main__":
    main()
