"""world_size-2 (gloo, CPU) test of the frame-sharded data-parallel step: each rank runs the oracle's
per-rank phases on its shard, the two exchanges of SURVEY.md 8e go through torch.distributed, and the
result must equal the unsharded minibatch (alpha exactly as defined over the GLOBAL minibatch)."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ml, beta, ret):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ls, Ms = [24, 20, 11], 16
    Mg = Ms * world
    rng = np.random.RandomState(6)
    W, b = O.init_weights(ls, seed=7)
    x = rng.randn(2 * Mg, ls[0]).astype(np.float32); t = rng.randn(2 * Mg, ls[-1]).astype(np.float32)
    net = O.OracleNet(ls, Ms, 0.1, 0.9, 1e-5, beta, ml, W, b)
    for step in range(2):
        xs = x[step * Mg + rank * Ms: step * Mg + (rank + 1) * Ms]
        ts = t[step * Mg + rank * Ms: step * Mg + (rank + 1) * Ms]
        local = net.dp_colsum(Mg, xs, ts)
        glob = torch.from_numpy(local.copy())
        dist.all_reduce(glob)                                  # exchange 1: 257-float sum |e|^beta
        net.dp_backward(Mg, xs, ts, glob.numpy(), local)
        gw, gb = net.grad_views()
        for g in gw + gb:
            tg = torch.from_numpy(g)                           # exchange 2: weight + bias gradients, in place
            dist.all_reduce(tg)
        net.dp_update(Mg)
    if rank == 0:
        Wn, bn = net.weights()
        ret["W"] = [w.copy() for w in Wn]; ret["b"] = [v.copy() for v in bn]; ret["alpha"] = net.alpha()
    dist.destroy_process_group()


@pytest.mark.parametrize("ml,beta", [(1, 1.5), (0, 2.0)])
def test_two_rank_dp_equals_single(ml, beta):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    O.lib()
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ml, beta, ret), nprocs=world, join=True)
    ls, Mg = [24, 20, 11], 32
    rng = np.random.RandomState(6)
    W, b = O.init_weights(ls, seed=7)
    x = rng.randn(2 * Mg, ls[0]).astype(np.float32); t = rng.randn(2 * Mg, ls[-1]).astype(np.float32)
    ref = O.OracleNet(ls, Mg, 0.1, 0.9, 1e-5, beta, ml, W, b)
    ref.train(x, t)
    Wr, br = ref.weights()
    for a, c in zip(ret["W"] + ret["b"], Wr + br):
        assert np.linalg.norm(a - c) <= 1e-6 * np.linalg.norm(c)
    if ml:
        assert np.allclose(ret["alpha"], ref.alpha(), rtol=1e-6)
