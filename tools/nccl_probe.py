"""2 ranks (torchrun): latency of the C-ABI all-gather of one uint64 per rank, and of torch's, with CUDA events."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from rspt_b200 import _lib
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
L = _lib.lib()
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    buf = (C.c_uint8 * 128)()
    L.rspt_gpu_comm_unique_id(buf)
    uid = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
dist.broadcast(uid, 0)
host = (C.c_uint8 * 128)(*uid.cpu().tolist())
comm = C.c_void_p()
assert L.rspt_gpu_comm_init(world, host, rank, local, C.byref(comm)) == 0
tot = torch.full((1,), rank + 1, dtype=torch.int64, device=dev)
allt = torch.zeros(world, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for name, fn in (("c_abi", lambda: L.rspt_gpu_allgather_totals(comm, tot.data_ptr(), allt.data_ptr(), st)),
                 ("torch", lambda: dist.all_gather_into_tensor(allt, tot))):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(200):
        fn()
    e1.record(); th = time.perf_counter() - t0
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{name}: gpu {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call, host enqueue {th / 200 * 1e6:.1f} us per call, result {allt.tolist()}", flush=True)
L.rspt_gpu_comm_destroy(comm)
dist.destroy_process_group()
