#!/usr/bin/env python
"""Benchmark of the signal-packer hot path (BASELINE.json: compress/decompress raw GB/s, CR, % HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one compress_batch call over `--frames` independent frames per GPU of BASELINE config 2
(xdelta_hzr, 12 ch x 3 B x 8192 samples, synthetic ECG-like data generated on the device).  The
headline `value` is raw-input GB/s with inputs resident in HBM; `e2e` is the same metric through
the host-buffer C-ABI call (H2D + D2H inside the timed region).  One JSON line on stdout (rank 0).

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libref.so when it
was built in the container, else the oracle port) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPES = {  # A = config 2, B = configs 3/4, C = config 1
    "A": dict(bps=3, ch=12, ns=8192),
    "B": dict(bps=4, ch=12, ns=4096),
    "C": dict(bps=4, ch=1, ns=8192),
}
WORKLOAD = "xdelta_hzr batched: 12 ch x 3 B/sample x 8192 samples, synthetic ECG-like frames (BASELINE configs[1])"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU packer on the host cores: one packer instance per thread, each looping
    compress over its own frames (ctypes releases the GIL)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    impl = "reference" if O.ref_available() else "port"
    if impl == "port":
        O.build(ref=False)
    sh = SHAPES["A"]
    fb = sh["bps"] * sh["ch"] * sh["ns"]
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = max(1, min(cores, args.cpu_threads or cores))
    per = args.cpu_frames_per_thread
    frames = O.synth_ecg(0, threads * per, sh["bps"], sh["ch"], sh["ns"]).reshape(threads, per, fb)
    packers = [O.make_packer("xdelta_hzr", sh["bps"], sh["ch"], sh["ns"], 3, impl) for _ in range(threads)]
    sizes = [None] * threads

    def work(i):
        _, sz = packers[i].compress_many(frames[i])
        sizes[i] = sz

    def step():
        ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    el = [step() for _ in range(args.steps)]
    total = sum(el)
    raw = args.steps * threads * per * fb
    gbs = raw / total / 1e9
    comp = sum(int(s.sum()) for s in sizes)
    line = {
        "impl": "reference", "metric": "compress_raw_GBps", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "packer": "xdelta_hzr", "bps": sh["bps"], "ch": sh["ch"], "ns": sh["ns"],
                   "frames_per_step": threads * per, "cr": threads * per * fb / comp},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": impl,
                         "sample": f"{threads} threads x {per} frames x {args.steps} steps, compress only "
                                   f"(includes the reference's built-in verify-decode)"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
def cpu_baseline(raw_frames: np.ndarray, sh: dict, budget_s: float = 12.0) -> dict:
    """Reference CPU packer, one thread, on a bounded sample of the same frames (rank 0, N = 1)."""
    from oracle import oracle as O
    impl = "reference" if O.ref_available() else "port"
    if impl == "port":
        O.build(ref=False)
    fb = sh["bps"] * sh["ch"] * sh["ns"]
    p = O.make_packer("xdelta_hzr", sh["bps"], sh["ch"], sh["ns"], 3, impl)
    frames = raw_frames.reshape(-1, fb)
    t0 = time.perf_counter()
    p.compress_many(frames[:8])
    per = (time.perf_counter() - t0) / 8
    n = int(max(8, min(frames.shape[0], budget_s * 0.65 / per)))
    t0 = time.perf_counter()
    dst, sizes = p.compress_many(frames[:n])
    tc = time.perf_counter() - t0
    t0 = time.perf_counter()
    p.decompress_many(dst, n)
    td = time.perf_counter() - t0
    return {"value": n * fb / tc / 1e9, "unit": "GB/s", "cores": 1, "kind": impl,
            "decompress_value": n * fb / td / 1e9,
            "sample": f"{n} of the benchmark's frames, compress then decompress, 1 thread ({tc + td:.1f} s)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from rspt_b200 import packer as R
    from rspt_b200 import dist as RD

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL's own banner / debug lines go to stderr
        RD.init_process_group_quiet(dev)
    sh = SHAPES["A"]
    fb = sh["bps"] * sh["ch"] * sh["ns"]
    F = args.frames
    hbm_peak, peak_src = peaks()

    p = R.SignalPacker.new_xdelta_hzr(sh["bps"], sh["ch"], sh["ns"], 3, max_batch_frames=F)
    nbuf = 2  # two distinct input batches, each larger than the 126 MB L2
    first = rank * nbuf * F  # contiguous shard of the global frame index space per rank
    inputs = [R.synth_ecg(first + i * F, F, **sh) for i in range(nbuf)]
    out = p.alloc_output(F, sidecar=True)
    total1 = torch.zeros(1, dtype=torch.int64, device=dev)

    def step(i):
        b = p.compress_batch(inputs[i % nbuf], out=out)
        if world > 1:
            # the path's only collective: 8 bytes per rank, places this shard in the global stream
            total1.copy_(b.offsets[F:F + 1])
            allt = RD.allgather_totals(total1)
            RD.place_offsets(b.offsets, allt, rank)
        return b

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    c0 = p.counters()["kernel_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = p.counters()["kernel_launches"] - c0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    raw_step = F * fb
    value = world * args.steps * raw_step / (ms * 1e-3) / 1e9

    # ---- everything below is outside the headline timed region -------------------------------
    comp_bytes = int(out.offsets[F].item()) - (int(out.offsets[0].item()))
    cr = raw_step / comp_bytes
    # decompress throughput (device resident)
    dec = torch.empty(F * fb, dtype=torch.uint8, device=dev)
    b = p.compress_batch(inputs[0], out=out)
    for _ in range(2):
        p.decompress_batch(b, out=dec)
    torch.cuda.synchronize()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nd = max(3, args.steps // 2)
    d0.record()
    for _ in range(nd):
        p.decompress_batch(b, out=dec)
    d1.record()
    torch.cuda.synchronize()
    dec_gbs = nd * raw_step / (d0.elapsed_time(d1) * 1e-3) / 1e9
    roundtrip_ok = bool(torch.equal(dec, inputs[0]))
    # decode of the same stream WITHOUT its index (what a stream from the CPU reference looks like):
    # the index is rebuilt on the device first (k_hzr_build_index), inside the timed region
    p.decompress_batch(b, out=dec, use_sidecar=False)
    torch.cuda.synchronize()
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0.record()
    for _ in range(3):
        p.decompress_batch(b, out=dec, use_sidecar=False)
    n1.record()
    torch.cuda.synchronize()
    dec_noindex_gbs = 3 * raw_step / (n0.elapsed_time(n1) * 1e-3) / 1e9
    roundtrip_ok = roundtrip_ok and bool(torch.equal(dec, inputs[0]))
    # integrity check (hzr_verify on the GPU: header walk + CRC-32C of every block), device resident
    vst = torch.zeros(F, dtype=torch.int32, device=dev)
    p.verify_batch(b, status=vst)
    torch.cuda.synchronize()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0.record()
    for _ in range(nd):
        p.verify_batch(b, status=vst)
    v1.record()
    torch.cuda.synchronize()
    verify_gbs = nd * comp_bytes / (v0.elapsed_time(v1) * 1e-3) / 1e9
    verify_ok = not bool(vst.any().item())

    # per-stage device times -> roofline of the dominant kernel
    p.set_stage_timing(True)
    p.stage_times(reset=True)
    ns_t = max(3, min(args.steps, 8))
    for i in range(ns_t):
        p.compress_batch(inputs[i % nbuf], out=out)
    for _ in range(ns_t):
        p.decompress_batch(b, out=dec)
    st = p.stage_times(reset=True)
    p.set_stage_timing(False)
    stage_ms = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in st.items()}
    comp_stages = {k: stage_ms[k] for k in ("transform", "hist", "tree", "layout", "encode")}
    dom = max(comp_stages, key=comp_stages.get)
    planes_bytes = F * 3 * sh["ch"] * sh["ns"]
    # algorithmic bytes of each stage per launch
    alg = {"transform": raw_step + planes_bytes, "hist": planes_bytes, "tree": 0,
           "layout": 0, "encode": planes_bytes + comp_bytes}
    achieved = alg[dom] / (comp_stages[dom] * 1e-3) / 1e9 if comp_stages[dom] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("frames_per_launch"):
                traffic = tj["kernels"].get(dom, {}).get("dram_bytes_per_launch")
                if traffic is not None:
                    traffic = traffic * F / tj["frames_per_launch"]
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": {"transform": "k_xdelta_planes", "hist": "k_hzr_hist", "tree": "k_hzr_tree",
                                           "layout": "k_scan_offsets", "encode": "k_hzr_encode_sparse + k_hzr_encode"}[dom],
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "peak_source": peak_src, "ms_per_launch": comp_stages[dom],
                "pipeline": {"algorithmic_bytes_per_step": raw_step + comp_bytes,
                             "achieved": (raw_step + comp_bytes) * world * args.steps / (ms * 1e-3) / 1e9 / world,
                             "frac": (raw_step + comp_bytes) * args.steps / (ms * 1e-3) / 1e9 / hbm_peak},
                "stage_ms": stage_ms}

    # end to end through the host-buffer C-ABI call, pinned host memory
    Fe = min(F, args.e2e_frames)
    h_src = torch.empty(Fe * fb, dtype=torch.uint8, pin_memory=True)
    h_src.copy_(inputs[0][: Fe * fb])
    h_dst = torch.empty(Fe * p.max_compressed_size, dtype=torch.uint8, pin_memory=True)
    h_off = torch.empty(Fe + 1, dtype=torch.int64, pin_memory=True)
    pe = R.SignalPacker.new_xdelta_hzr(sh["bps"], sh["ch"], sh["ns"], 3, max_batch_frames=Fe)
    np_src, np_dst, np_off = h_src.numpy(), h_dst.numpy(), h_off.numpy().view(np.uint64)
    for _ in range(2):
        tot = pe.compress_batch_host(np_src, np_dst, np_off)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ne = max(3, min(args.steps, 6))
    t0 = time.perf_counter()
    for _ in range(ne):
        tot = pe.compress_batch_host(np_src, np_dst, np_off)
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([te], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t.item())
    e2e = {"value": world * ne * Fe * fb / te / 1e9, "unit": "GB/s", "h2d_bytes_per_step": Fe * fb,
           "d2h_bytes_per_step": int(tot) + 8 * (Fe + 1), "frames_per_step": Fe,
           "api": "rspt_gpu_compress_batch_host (pinned host buffers)"}

    extras = {}
    if rank == 0 and world == 1 and not args.quick:
        extras["packers"] = other_packers(R, torch, args)
        # the pre-filter step in front of the packers (rspt_test.cpp:116-136), device resident, in place
        n5 = [1.00000000000, -3.14332095199, 3.70064088865, -1.97083923944, 0.41351972908]
        d5 = [0.06722876941, 0.00000000000, -0.13445753881, 0.00000000000, 0.06722876941]
        fir = (np.hanning(33) / np.hanning(33).sum()).tolist()
        work = inputs[0].clone()
        pf = {}
        for name, fn in (("iir_bandpass_5", lambda: p.prefilter_iir(work, n5, d5, 2000)),
                         ("fir_33", lambda: p.prefilter_fir(work, fir))):
            fn()
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(3):
                fn()
            f1.record()
            torch.cuda.synchronize()
            pf[name + "_raw_GBps"] = 3 * raw_step / (f0.elapsed_time(f1) * 1e-3) / 1e9
        del work
        # the IIR chain is serial per frame (latency-bound, one warp per 32 frames): a larger batch takes
        # about the same time, so its throughput scales with the batch
        try:
            Fb = 8 * F
            pb = R.SignalPacker.new_xdelta_hzr(sh["bps"], sh["ch"], sh["ns"], 3, max_batch_frames=Fb)
            big = R.synth_ecg(first, Fb, **sh)
            pb.prefilter_iir(big, n5, d5, 2000)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            pb.prefilter_iir(big, n5, d5, 2000)
            f1.record()
            torch.cuda.synchronize()
            pf["iir_bandpass_5_raw_GBps_at_%d_frames" % Fb] = Fb * fb / (f0.elapsed_time(f1) * 1e-3) / 1e9
            pb.close()
            del big
            torch.cuda.empty_cache()
        except Exception as ex:  # out of memory on a shared box: the figure is optional
            pf["iir_large_batch_error"] = str(ex)[:120]
        pf["note"] = ("bit-identical to i_filter; the IIR walks a frame's channels in sequence like the reference "
                      "(one thread per frame, latency-bound: its time barely depends on the batch size)")
        extras["prefilter"] = pf
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(inputs[0][: min(F, 4096) * fb].cpu().numpy(), sh, args.cpu_budget)

    if rank == 0:
        line = {
            "metric": "compress_raw_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "packer": "xdelta_hzr", "bps": sh["bps"], "ch": sh["ch"], "ns": sh["ns"],
                       "nb": 3, "frames_per_step_per_gpu": F, "raw_bytes_per_step_per_gpu": raw_step,
                       "l2": "inputs (%.2f GB per step, 2 alternating batches) exceed the 126 MB L2" % (raw_step / 1e9),
                       "sharding": "contiguous frame ranges per rank; one NCCL all-gather of 8 B/rank per step" if world > 1 else "single GPU",
                       "cr": cr},
            "decompress_raw_GBps": dec_gbs, "roundtrip_bit_exact": roundtrip_ok, "cr": cr,
            "decompress_noindex_raw_GBps": dec_noindex_gbs, "verify_compressed_GBps": verify_gbs, "verify_ok": verify_ok,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def other_packers(R, torch, args):
    """Short device-resident runs of the other packers (reported, not the headline)."""
    res = {}
    for kind, shape, F in (("hzr", "A", 2048), ("hadamard", "B", 2048), ("dct", "B", 2048)):
        sh = SHAPES[shape]
        fb = sh["bps"] * sh["ch"] * sh["ns"]
        p = R.SignalPacker(kind, sh["bps"], sh["ch"], sh["ns"], 3, max_batch_frames=F)
        x = R.synth_ecg(0, F, **sh)
        out = p.alloc_output(F)
        dec = torch.empty_like(x)
        for _ in range(2):
            b = p.compress_batch(x, out=out)
            p.decompress_batch(b, out=dec)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        n = 5
        ev[0].record()
        for _ in range(n):
            b = p.compress_batch(x, out=out)
        ev[1].record()
        for _ in range(n):
            p.decompress_batch(b, out=dec)
        ev[2].record()
        torch.cuda.synchronize()
        comp = int(out.offsets[F].item())
        r = {"shape": sh, "frames": F, "compress_raw_GBps": n * F * fb / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9,
             "decompress_raw_GBps": n * F * fb / (ev[1].elapsed_time(ev[2]) * 1e-3) / 1e9, "cr": F * fb / comp}
        if kind in ("hadamard", "dct"):
            r["prdn_percent"] = R.prdn(x, dec, F, sh["bps"], sh["ch"], sh["ns"])
        else:
            r["roundtrip_bit_exact"] = bool(torch.equal(x, dec))
        res[kind] = r
        p.close()
        del x, out, dec
        torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames per step per GPU")
    ap.add_argument("--e2e-frames", type=int, default=4096)
    ap.add_argument("--quick", action="store_true", help="skip the other packers")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--cpu-frames-per-thread", type=int, default=64)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
