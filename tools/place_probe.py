"""2 ranks (torchrun): where the time of per-batch placement goes in the headline loop (host enqueue vs GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rspt_b200 import packer as R
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
D = bench.Dist(torch, dev)
rank, world = D.rank, D.world
sh = bench.SHAPES["A"]
F, NB = 4096, int(os.environ.get("NB", "8"))
p = R.SignalPacker.new_xdelta_hzr(sh["bps"], sh["ch"], sh["ns"], 3, max_batch_frames=F)
inputs = [R.synth_ecg(rank * NB * F + i * F, F, **sh) for i in range(NB)]
outs = [p.alloc_output(F, sidecar=True) for _ in range(2)]
for mode in ("none", "place", "place_sync_step", "place"):
    def step(rec):
        th = 0.0
        for i in range(NB):
            b = p.compress_batch(inputs[i], out=outs[i & 1])
            if mode == "place" or mode == "place_sync_each" or (mode == "place_every_4" and i % 4 == 3):
                t0 = time.perf_counter()
                p.place_offsets_async(D.comm, b, rank, world)
                th += time.perf_counter() - t0
            if mode == "place_sync_each":
                torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode != "none":
            p.place_join()
        th += time.perf_counter() - t0
        if mode == "place_sync_step":
            torch.cuda.synchronize()
        rec.append(th)
    for _ in range(3):
        step([])
    torch.cuda.synchronize(); D.barrier()
    rec = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10):
        step(rec)
    e1.record(); tq = time.perf_counter() - t0
    torch.cuda.synchronize(); tw = time.perf_counter() - t0
    print(f"rank {rank} {mode}: gpu {e0.elapsed_time(e1) / 10:.2f} ms/step, host enqueue {tq * 100:.2f} ms/step, "
          f"wall {tw * 100:.2f}, in place calls {sum(rec) * 100:.3f} ms/step", flush=True)
    D.barrier()
p.close()
