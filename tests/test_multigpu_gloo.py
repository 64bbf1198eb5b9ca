"""CPU, world_size 2, gloo: the batch-sharded sampling path (shard -> sample -> all-gather) gives the
same global result as a single process.  The "sampler" here is the oracle's DDIM on a toy eps-model —
the collective plumbing is what is under test."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restate as R
from oracle.make_golden import toy_model_fn
from sdb200.distributed import gather_images, per_sample_randn, shard_range

B, SHAPE = 5, (4, 8, 8)


def _sample(indices):
    x_T = per_sample_randn(indices, SHAPE, 1000)
    c = per_sample_randn(indices, (5, 6), 2000)
    if x_T.shape[0] == 0:
        return x_T
    z, _ = R.DDIMOracle(R.ModelShim(toy_model_fn, R.sd_alphas_cumprod())).sample(10, x_T.shape[0], SHAPE, conditioning=c, x_T=x_T)
    return z


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(B, rank, world)
    full = gather_images(_sample(range(lo, hi)), B)
    if rank == 0:
        torch.save(full, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sampling_equals_single_process(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "full.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    assert got.shape == (B,) + SHAPE
    assert torch.equal(got, _sample(range(B)))
