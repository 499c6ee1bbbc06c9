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
