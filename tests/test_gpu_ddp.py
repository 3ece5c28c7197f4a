"""GPU x2 (skipped on a 1-GPU box): the propagation module inside stock DistributedDataParallel
over NCCL -- SURVEY 8(f1)/8(e): the path itself has no collective; the only cross-GPU quantity is
the gradient of the scalar gamma (and of whatever network feeds the module), which rides in DDP's
ordinary all-reduce.  A small convolutional head stands in for the reference's encoder-decoder."""
import os
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import torch.nn as nn
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import make_inputs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    K, T, B, H, W = 3, 6, 2, 32, 48

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)                       # same weights on every rank
            self.head = nn.Conv2d(4, 3 * (K * K - 1) + 2, 3, padding=1)
            self.prop = NLSPN(prop_kernel=K, prop_time=T)

        def forward(self, rgb, dep):
            f = self.head(torch.cat([rgb, dep], 1))
            guidance, init, conf = f[:, :3 * (K * K - 1)], torch.relu(f[:, -2:-1]), torch.sigmoid(f[:, -1:])
            return self.prop(init, guidance.contiguous(), conf, dep)[0]

    inp = make_inputs(B, H, W, K, seed=100 + rank, device=dev)       # different data per rank
    rgb = torch.randn(B, 3, H, W, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
    net = Net().to(dev)
    # reference: per-rank gradients without DDP, averaged by hand
    loss = (net(rgb, inp["feat_fix"]).clamp(min=0) - inp["gt"]).abs().mean()
    loss.backward()
    manual = [p.grad.clone() for p in net.parameters() if p.requires_grad]
    for g in manual:
        dist.all_reduce(g)
        g /= world
    net.zero_grad(set_to_none=True)
    ddp = nn.parallel.DistributedDataParallel(net, device_ids=[rank])
    loss = (ddp(rgb, inp["feat_fix"]).clamp(min=0) - inp["gt"]).abs().mean()
    loss.backward()
    got = [p.grad for p in net.parameters() if p.requires_grad]
    ok = all(torch.allclose(a, b, rtol=1e-4, atol=1e-7) for a, b in zip(got, manual))
    gam = net.prop.aff_scale_const.grad
    flag = torch.tensor([1.0 if ok and gam is not None and torch.isfinite(gam).all() else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        with open(os.path.join(out_dir, "ok"), "w") as f:
            f.write("%d %r" % (int(flag.item()), float(gam)))
    dist.destroy_process_group()


def test_nlspn_under_ddp_two_gpus(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    txt = open(tmp_path / "ok").read()
    assert txt.startswith("1"), txt
