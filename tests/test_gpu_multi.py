"""Two-rank NCCL run of the bag-parallel path on a box with >= 2 GPUs: sharded encode + the single
gather must equal the serial per-bag forward bit for bit."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, results):
    import torch.distributed as dist
    from oracle import rrt_oracle as O
    from rrt_mil_b200 import RRTEncoder, parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = O.EncoderConfig()
        m = RRTEncoder(**cfg.to_dict()).cuda().eval()
        m.load_state_dict({k: v.float() for k, v in O.make_weights(cfg, 3).items()})
        g = torch.Generator().manual_seed(0)
        bags = [torch.randn(n, 512, generator=g) for n in (3000, 777, 2048, 1500, 64)]
        outs = parallel.encode_bags_parallel(m, bags)
        with torch.no_grad():
            ok = all(torch.equal(o, m(b.cuda())) for o, b in zip(outs, bags))
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_gather_equals_serial():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}


def _train_worker(rank, world, port, results):
    """Data-parallel RRTMIL train step (training): every rank runs forward + CE + backward on
    ITS bag, one flat all-reduce averages the gradients, the fused Adam step updates the replicas; rank 0
    also runs the serial two-bag average and compares."""
    import torch.distributed as dist
    import torch.nn.functional as F
    from rrt_mil_b200 import RRTMIL, parallel
    from rrt_mil_b200.optim import Adam
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        def model():
            torch.manual_seed(5)
            m = RRTMIL(input_dim=512, n_classes=2).cuda().train()
            m._dropout_seed, m.online_encoder._dropout_seed = 100, 101
            return m
        g = torch.Generator().manual_seed(1)
        bags = [torch.randn(1, 700 + 150 * i, 512, generator=g) for i in range(world)]
        labels = [i % 2 for i in range(world)]
        m = model()
        opt = Adam(m.parameters(), lr=2e-4, weight_decay=1e-5)
        F.cross_entropy(m(bags[rank].cuda()), torch.tensor([labels[rank]], device="cuda")).backward()
        n_coll = parallel.allreduce_gradients(list(m.parameters()))
        opt.step()
        ok = n_coll == 1          # 10.8 MB of gradients: one bucket, one collective
        ms = model()              # serial reference on this rank's GPU: average of the per-bag gradients
        acc = [torch.zeros_like(p) for p in ms.parameters()]
        for b, y in zip(bags, labels):
            ms.zero_grad(set_to_none=True)
            F.cross_entropy(ms(b.cuda()), torch.tensor([y], device="cuda")).backward()
            for a, p in zip(acc, ms.parameters()):
                a += p.grad / world
        for a, p in zip(acc, ms.parameters()):
            p.grad = a
        Adam(ms.parameters(), lr=2e-4, weight_decay=1e-5).step()
        for (n, p), q in zip(m.named_parameters(), ms.parameters()):
            # same kernels, same masks; only the order of fp32 atomic sums and of the all-reduce differs,
            # and Adam turns a sign flip of a ~0 gradient into a 2*lr step
            d = (p - q).abs()
            ok = ok and float(d.mean()) < 2e-6 and float(d.max()) <= 2 * 2e-4 + 1e-6
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_data_parallel_train_step_equals_serial_average():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_train_worker, args=(2, _free_port(), results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}
