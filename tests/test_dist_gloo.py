"""Host-side logic of the multi-GPU path on CPU: two ranks over gloo.  Frames shard contiguously,
the only exchange is one all-gather of a uint64 byte total per rank, and every rank places its
frame offsets in the concatenated stream (rspt_b200/dist.py)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from rspt_b200 import dist as RD
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    F = 37                                   # frames of the whole job: ragged split over 2 ranks
    lo, hi = RD.shard_range(F, rank, world)
    all_ranges = [RD.shard_range(F, r, world) for r in range(world)]
    assert all_ranges[0][0] == 0 and all_ranges[-1][1] == F
    assert all(all_ranges[r][1] == all_ranges[r + 1][0] for r in range(world - 1))
    rng = np.random.default_rng(5)           # same stream on both ranks: "sizes" of all F frames
    sizes = rng.integers(67, 295000, F).astype(np.int64)
    mine = torch.from_numpy(sizes[lo:hi])
    offsets = torch.zeros(hi - lo + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(mine, 0)      # what compress_batch returns for this shard
    total = offsets[-1:].clone()
    allt = RD.allgather_totals(total)        # the path's only collective: 8 bytes per rank
    assert allt.tolist() == [int(sizes[a:b].sum()) for a, b in all_ranges]
    placed = RD.place_offsets(offsets, allt, rank)
    want = np.concatenate([[0], np.cumsum(sizes)])[lo:hi + 1]
    assert placed.tolist() == want.tolist(), (rank, placed.tolist()[:4], want.tolist()[:4])
    assert RD.base_offset(allt, rank) == int(want[0])
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_two_ranks_place_their_shards(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
