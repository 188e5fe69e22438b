#!/usr/bin/env python
"""BASELINE.json configs[4]: all four packers on a 64 GB synthetic 12-channel 24-bit stream, sharded
contiguously over the GPUs of one box, with the NCCL all-gather of per-rank compressed byte totals
that places every shard in the concatenated output.

    python tools/config5.py [--gb 64] [--batch 8192] [--packers xdelta_hzr,hzr,hadamard,dct]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/config5.py

Frames are generated on the owning GPU batch by batch (the stream does not fit HBM next to the
scratch), compressed, decompressed and checked: lossless packers must give back the input bit for
bit, lossy ones report PRDN.  Times are CUDA events on the device, summed over the batches, max over
ranks.  One JSON line per packer on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from rspt_b200 import packer as R
    from rspt_b200 import dist as RD

    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=64.0, help="stream size in GB (1e9 bytes)")
    ap.add_argument("--batch", type=int, default=8192, help="frames per batch per GPU")
    ap.add_argument("--packers", default="xdelta_hzr,hzr,hadamard,dct")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        RD.init_process_group_quiet(dev)
    bps, ch, ns = 3, 12, 8192
    fb = bps * ch * ns
    total_frames = int(args.gb * 1e9) // fb
    lo, hi = RD.shard_range(total_frames, rank, world)
    B = args.batch

    for kind in args.packers.split(","):
        p = R.SignalPacker(kind, bps, ch, ns, 3, max_batch_frames=B)
        out = p.alloc_output(B, sidecar=True)
        raw = torch.empty(B * fb, dtype=torch.uint8, device=dev)
        dec = torch.empty_like(raw)
        total1 = torch.zeros(1, dtype=torch.int64, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_c = t_d = 0.0
        comp = 0
        exact = True
        prdn_num = prdn_den = 0.0
        for f0 in range(lo, hi, B):
            n = min(B, hi - f0)
            R.synth_ecg(f0, n, bps, ch, ns, out=raw)
            x = raw[: n * fb]
            ev[0].record()
            b = p.compress_batch(x, out=out)
            ev[1].record()
            p.decompress_batch(b, out=dec)
            ev[2].record()
            torch.cuda.synchronize()
            t_c += ev[0].elapsed_time(ev[1])
            t_d += ev[1].elapsed_time(ev[2])
            comp += int(b.offsets[n].item())  # frame offsets inside the shard = comp so far + b.offsets
            if kind in ("xdelta_hzr", "hzr"):
                exact = exact and bool(torch.equal(x, dec[: n * fb]))
            else:
                num, den = R.prdn_terms(x, dec[: n * fb], n, bps, ch, ns)
                prdn_num += num
                prdn_den += den
        # the path's only collective: per-rank compressed byte totals -> base offset of every shard in
        # the concatenated stream (rank-major = frame order); timed as part of compress
        total1.fill_(comp)
        ev[0].record()
        allt = RD.allgather_totals(total1)
        last = torch.tensor([0, comp], dtype=torch.int64, device=dev)
        RD.place_offsets(last, allt, rank)  # [shard base, shard end] in the global stream
        ev[1].record()
        torch.cuda.synchronize()
        t_c += ev[0].elapsed_time(ev[1])
        shard_lo, shard_hi = (int(v) for v in last.tolist())
        assert shard_hi - shard_lo == comp and shard_lo == int(allt[:rank].sum().item())
        stats = torch.tensor([t_c, t_d, float(comp), float((hi - lo) * fb), prdn_num, prdn_den, 1.0 if exact else 0.0],
                             dtype=torch.float64, device=dev)
        if world > 1:
            mx = stats.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stats.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            mn = stats.clone()
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        else:
            mx = sm = mn = stats
        if rank == 0:
            raw_total = float(sm[3].item())
            line = {
                "config": "BASELINE configs[4]: %.0f GB of 12 ch x 3 B x 8192 frames, contiguous shards" % args.gb,
                "packer": kind, "n_gpus": world, "frames": total_frames, "batch_frames_per_gpu": B,
                "compress_raw_GBps": raw_total / (float(mx[0].item()) * 1e-3) / 1e9,
                "decompress_raw_GBps": raw_total / (float(mx[1].item()) * 1e-3) / 1e9,
                "cr": raw_total / float(sm[2].item()),
                "collective": "one all_gather_into_tensor of 8 B per rank (NCCL)" if world > 1 else "none (1 GPU)",
                "shard_bytes": [int(v) for v in allt.tolist()],
            }
            if kind in ("xdelta_hzr", "hzr"):
                line["roundtrip_bit_exact_all_frames"] = bool(mn[6].item() == 1.0)
            else:
                line["prdn_percent"] = 100.0 * (float(sm[4].item()) / max(float(sm[5].item()), 1e-30)) ** 0.5
            print(json.dumps(line), flush=True)
        p.close()
        del out, raw, dec
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
