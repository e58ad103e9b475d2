"""Data-parallel host logic on CPU: 2 processes, gloo.  (The fused kernels need a GPU; here the
per-rank gradients come from the oracle so that the DP contract itself is what is tested:
row-sharded batches + loss divided by the GLOBAL batch + all-reduce(sum) == single-process gradient.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import nif_b200
    from nif_b200.distributed import DataParallel
    from oracle import nif_oracle as O
    torch.set_num_threads(1)
    dp = DataParallel("gloo")
    assert dp.world == world and dp.rank == rank
    cfg_s = {"use_resblock": False, "connectivity": "full", "input_dim": 2, "output_dim": 1, "units": 8, "nlayers": 1,
             "weight_init_factor": 0.1, "omega_0": 30.0}
    cfg_p = {"use_resblock": False, "input_dim": 1, "latent_dim": 2, "units": 6, "nlayers": 1, "activation": "swish"}
    spec = O.spec_from_cfg("NIFMultiScale", cfg_s, cfg_p)
    # replicas start from rank 0's parameters
    net = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=10 + rank, device="cpu")
    dp.broadcast_(net.theta)
    ref = nif_b200.NIFMultiScale(cfg_s, cfg_p, seed=10, device="cpu")
    assert torch.equal(net.theta, ref.theta)
    prm = {k: v.detach().double() for k, v in net.variables.items()}
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (12, 3)).astype(np.float32)
    Y = rng.uniform(-1, 1, (12, 1)).astype(np.float32)
    ds = nif_b200.Dataset.from_tensor_slices((X, Y)).shuffle(12, seed=3).batch(8).shard(world, rank)
    full = nif_b200.Dataset.from_tensor_slices((X, Y)).shuffle(12, seed=3).batch(8)
    for (gb, mine), (_, allrows) in zip(ds.batches(0), full.batches(0)):
        # local gradient of  sum_local mse_rows / global_batch
        xi, yi = mine[0].double(), mine[1].double()
        loss, g, _, _ = O.loss_and_grads(spec, prm, xi, yi)
        scale = xi.shape[0] / gb
        flat = torch.cat([g[k].reshape(-1) * scale for k in prm])
        dp.allreduce_(flat)
        l1, g1, _, _ = O.loss_and_grads(spec, prm, allrows[0].double(), allrows[1].double())
        flat1 = torch.cat([g1[k].reshape(-1) for k in prm])
        assert torch.allclose(flat, flat1, atol=1e-12), float((flat - flat1).abs().max())
        lt = torch.tensor([float(loss) * scale], dtype=torch.float64)
        dp.allreduce_(lt)
        assert abs(float(lt) - float(l1)) < 1e-12
    # the split, overlapped reduction the training step uses (head gradient first, trunk gradient second)
    buf = torch.arange(6, dtype=torch.float32) + rank
    h_head = dp.allreduce_start(buf[2:])
    h_trunk = dp.allreduce_start(buf[:2])
    dp.allreduce_finish(h_head)
    dp.allreduce_finish(h_trunk)
    assert torch.equal(buf, world * torch.arange(6, dtype=torch.float32) + sum(range(world)))
    m = torch.tensor([float(rank)])
    dp.max_(m)
    assert float(m) == world - 1
    dp.barrier()
    dp.shutdown()
    out.put(rank)


def test_two_rank_gloo_data_parallel_contract():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]
