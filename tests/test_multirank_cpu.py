"""world_size-2 gloo test (CPU) of the only multi-GPU logic of the path: batch sharding and the final
all-reduce (SUM) of the metric accumulators, which must reproduce the single-process global / per-frame means
(reference src/lib/metrics.py:207-212)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, psnr, mse, ssim, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from textocvp_b200 import rollout
    lo, hi = rollout.shard_range(rank, world, psnr.shape[0])
    m = rollout.MetricSums(psnr.shape[1], torch.device("cpu"))
    for s in range(lo, hi, 3):                       # several "batches" per rank
        m.accumulate(psnr[s:min(hi, s + 3)], mse[s:min(hi, s + 3)], ssim[s:min(hi, s + 3)])
    res = m.all_reduce().results()
    if rank == 0:
        out.put(res)
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    from textocvp_b200 import rollout
    for total in (1, 7, 256, 2048):
        for world in (1, 2, 3, 8):
            spans = [rollout.shard_range(r, world, total) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_metric_allreduce_two_ranks():
    g = torch.Generator().manual_seed(0)
    psnr = torch.rand(11, 19, generator=g) * 40
    mse = torch.rand(11, 19, generator=g)
    ssim = torch.rand(11, 19, generator=g)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, psnr, mse, ssim, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["count"] == 11
    assert abs(res["psnr_mean"] - psnr.double().mean().item()) < 1e-9
    assert torch.allclose(torch.tensor(res["psnr_per_frame"], dtype=torch.float64), psnr.double().mean(0), atol=1e-9)
    assert torch.allclose(torch.tensor(res["mse_per_frame"], dtype=torch.float64), mse.double().mean(0), atol=1e-9)
    assert torch.allclose(torch.tensor(res["ssim_per_frame"], dtype=torch.float64), ssim.double().mean(0), atol=1e-9)
    assert res["lpips_mean"] == 0.0          # slot reserved, never filled offline
