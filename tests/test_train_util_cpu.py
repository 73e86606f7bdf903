"""CPU, world_size 2 over gloo: the DDP training-step plumbing of bench.py --mode train (sigma_b200/train_util.py) with a
stand-in model that has the reference's forward(rgb, modal_x, label) -> loss signature (the Sigma modules themselves have no
CPU path).  Checks: gradients are averaged over ranks (DDP all-reduce), `sync=False` skips the exchange, parameters stay
identical across ranks after steps, group_weight reproduces the reference's decay / no-decay split."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(6, 8, 3, padding=1)
        self.norm = nn.LayerNorm(8)
        self.head = nn.Linear(8, 5)
        self.bare = nn.Parameter(torch.ones(8))           # like A_logs / Ds: in no optimizer group (init_func.py quirk)
        self.criterion = nn.CrossEntropyLoss(ignore_index=255)

    def forward(self, rgb, modal_x, label=None):
        y = self.conv(torch.cat([rgb, modal_x], 1)).permute(0, 2, 3, 1) * self.bare
        out = self.head(self.norm(y)).permute(0, 3, 1, 2)
        return out if label is None else self.criterion(out, label.long())


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    from sigma_b200 import dist_util, train_util
    dist_util.init("gloo")
    torch.manual_seed(0)
    model = Toy()
    opt = train_util.make_optimizer(model)
    ddp = train_util.wrap_ddp(model)
    step = train_util.TrainStep(ddp, opt, amp_dtype=None, device_type="cpu")
    g = torch.Generator().manual_seed(100 + rank)                 # every rank its own shard
    rgb, mx = torch.randn(2, 3, 8, 8, generator=g), torch.randn(2, 3, 8, 8, generator=g)
    gt = torch.randint(0, 5, (2, 8, 8), generator=g)
    # reference gradient: mean over ranks of the local gradients
    loc = Toy()
    loc.load_state_dict(model.state_dict())
    loc(rgb, mx, gt).backward()
    gl = loc.head.weight.grad.clone()
    gsum = gl.clone()
    dist.all_reduce(gsum)
    with ddp.no_sync():
        ddp(rgb, mx, gt).backward()
    unsynced = model.head.weight.grad.clone()
    model.zero_grad()
    ddp(rgb, mx, gt).backward()
    synced = model.head.weight.grad.clone()
    for _ in range(3):
        loss = step(rgb, mx, gt)
    w = model.head.weight.detach().clone()
    ws = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    t, bw = train_util.allreduce_bus_bandwidth(train_util.grad_bytes(model), torch.device("cpu"), reps=2)
    q.put((rank, float((unsynced - gl).abs().max()), float((synced - gsum / world).abs().max()), float((ws[0] - ws[1]).abs().max()),
           float(loss), t > 0 and bw > 0, float((model.bare.detach() - 1).abs().max())))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_train_step_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e_unsynced, e_synced, e_w, loss, bw_ok, bare_moved in res:
        assert e_unsynced < 1e-6, "no_sync must leave the local gradient untouched"
        assert e_synced < 1e-6, "DDP gradient must be the mean over ranks"
        assert e_w == 0.0, "parameters diverged across ranks"
        assert loss == loss and bw_ok
        assert bare_moved == 0.0, "bare nn.Parameters are in no optimizer group (utils/init_func.py:33-56)"


def test_group_weight_split():
    sys.path.insert(0, ROOT)
    from sigma_b200 import train_util
    m = Toy()
    g = train_util.group_weight(m, 1e-3)
    assert {id(p) for p in g[0]["params"]} == {id(m.conv.weight), id(m.head.weight)}
    assert {id(p) for p in g[1]["params"]} == {id(m.conv.bias), id(m.head.bias), id(m.norm.weight), id(m.norm.bias)}
    assert g[1]["weight_decay"] == 0.0
