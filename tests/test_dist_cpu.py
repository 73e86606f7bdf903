"""CPU, world_size 2 over gloo: the N>1 bookkeeping of bench.py (replicas, MAX-over-ranks timing, rank-0 reporting)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    from sigma_b200 import dist_util
    w, r, l = dist_util.init("gloo")
    assert (w, r, l) == (world, rank, rank)
    slow = dist_util.max_over_ranks(10.0 + 5.0 * rank)          # rank 1 is the slow one
    g = torch.Generator().manual_seed(dist_util.shard_seed(1234, rank))
    shard = torch.randn(4, generator=g)
    gathered = [torch.zeros(4) for _ in range(world)]
    dist.all_gather(gathered, shard)
    q.put((rank, slow, dist_util.aggregate_images_per_s(8, w, 10, slow), [t.tolist() for t in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_replica_bookkeeping_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, slow, ips, shards in res:
        assert slow == 15.0                                     # MAX over ranks, identical on every rank
        assert abs(ips - 8 * 2 * 10 / 15e-3) < 1e-6             # whole-job images/s over the slowest rank's time
        assert shards[0] != shards[1]                           # disjoint synthetic shards (seed + rank)
    assert res[0][3] == res[1][3]


def test_reference_arm_under_torchrun_prints_once():
    """`bench.py --impl reference` launched with 2 ranks: rank 0 times the CPU oracle, rank 1 exits 0 without output."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + os.getpid() % 400), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--height", "64", "--width", "96"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["n_gpus"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
