"""Frame-sharded data parallelism on real GPUs (needs >= 2 devices; `gpurun --gpus 2`): two ranks, each training on
its shard with NCCL allreduce of sum|e|^beta and of the gradients, must reproduce the unsharded global minibatch."""
import os
import queue as _queue
import sys
import time
import numpy as np
import pytest


def _collect(q, ps, timeout):
    """results of all ranks; fails FAST when a rank process died (a 2-GPU box is charged twice per second waited)"""
    res, t0 = [], time.time()
    while len(res) < len(ps):
        try:
            res.append(q.get(timeout=2))
        except _queue.Empty:
            dead = [p.exitcode for p in ps if p.exitcode not in (None, 0)]
            if dead:
                for p in ps:
                    if p.is_alive():
                        p.terminate()
                raise RuntimeError("a rank process died (exit codes %s)" % dead)
            if time.time() - t0 > timeout:
                for p in ps:
                    if p.is_alive():
                        p.terminate()
                raise RuntimeError("ranks did not finish within %d s" % timeout)
    return sorted(res, key=lambda r: r[0])

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LS, MS, NB = [70, 96, 80, 33], 128, 6
NAMED = [1799, 2048, 2048, 2048, 257]


def _data(world, LS=LS, NB=NB):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    rng = np.random.RandomState(13)
    W, b = O.init_weights(LS, seed=3)
    x = rng.randn(world, NB * MS, LS[0]).astype(np.float32)     # [rank][frames][dim]
    t = rng.randn(world, NB * MS, LS[-1]).astype(np.float32)
    return W, b, x, t


def _rank(rank, world, uid, ml, beta, precision, q, LS=LS, NB=NB):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_pkg
    pkg = load_pkg()
    W, b, x, t = _data(world, LS, NB)
    net = pkg.BP_GPU(0, rank, len(LS), LS, MS, 0.1, 0.9, 1e-5, W, b, beta, ml, precision=precision, world_size=world, rank=rank,
                     nccl_unique_id=uid)
    net.train(NB * MS, x[rank], t[rank])
    Wn, bn = net.returnWeights()
    q.put((rank, [w.copy() for w in Wn], [v.copy() for v in bn], net.alpha(), net.losses()))
    net.close()


@pytest.mark.parametrize("ml,beta,precision", [(1, 1.5, 0), (1, 1.5, 1), (0, 2.0, 0)])
def test_two_gpu_dp_equals_unsharded(pkg, oracle, ml, beta, precision):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from se_ml_b200.bp_gpu import nccl_unique_id
    world = 2
    uid = nccl_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_rank, args=(r, world, uid, ml, beta, precision, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = _collect(q, ps, 240)
    for p in ps:
        p.join(60)
    W, b, x, t = _data(world)
    # unsharded equivalent: global minibatch = rank 0's bunch followed by rank 1's bunch
    xg = np.concatenate([x[:, i * MS:(i + 1) * MS].reshape(world * MS, -1) for i in range(NB)])
    tg = np.concatenate([t[:, i * MS:(i + 1) * MS].reshape(world * MS, -1) for i in range(NB)])
    orc = oracle.OracleNet(LS, world * MS, 0.1, 0.9, 1e-5, beta, ml, W, b)
    lo, al = orc.train(xg, tg)
    Wo, bo = orc.weights()
    tol = 1e-3 if precision == 0 else 1e-4
    for r in res:
        for a, c in zip(r[1] + r[2], Wo + bo):
            assert np.linalg.norm(a - c) <= tol * max(np.linalg.norm(c), 1e-6)
        if ml:
            assert np.linalg.norm(r[3] - al[-1]) <= tol * np.linalg.norm(al[-1])     # alpha of the GLOBAL minibatch
            assert np.allclose(r[4], lo, rtol=5e-3)
    # both ranks hold identical weights
    for a, c in zip(res[0][1], res[1][1]):
        assert np.array_equal(a, c)


def test_two_gpu_dp_named_shape(pkg, oracle):
    """the named network, 128 frames per GPU (global minibatch 256): gradient tiles pushed to their owners over NVLink,
    owner-side update, shadow broadcast -- against the oracle on the unsharded minibatch"""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from se_ml_b200.bp_gpu import nccl_unique_id
    world, nb = 2, 3
    uid = nccl_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_rank, args=(r, world, uid, 1, 1.5, 0, q, NAMED, nb)) for r in range(world)]
    for p in ps:
        p.start()
    res = _collect(q, ps, 400)
    for p in ps:
        p.join(60)
    W, b, x, t = _data(world, NAMED, nb)
    xg = np.concatenate([x[:, i * MS:(i + 1) * MS].reshape(world * MS, -1) for i in range(nb)])
    tg = np.concatenate([t[:, i * MS:(i + 1) * MS].reshape(world * MS, -1) for i in range(nb)])
    orc = oracle.OracleNet(NAMED, world * MS, 0.1, 0.9, 1e-5, 1.5, 1, W, b)
    lo, al = orc.train(xg, tg)
    Wo, bo = orc.weights()
    for r in res:
        for a, c, w0 in zip(r[1] + r[2], Wo + bo, W + b):
            assert np.linalg.norm(a - c) <= 1e-3 * max(np.linalg.norm(c), 1e-6)
            assert np.linalg.norm((a - w0) - (c - w0)) <= 2e-3 * max(np.linalg.norm(c - w0), 1e-6)   # the update itself
        assert np.linalg.norm(r[3] - al[-1]) <= 1e-3 * np.linalg.norm(al[-1])
        assert np.allclose(r[4], lo, rtol=5e-3)
    for a, c in zip(res[0][1] + res[0][2], res[1][1] + res[1][2]):
        assert np.array_equal(a, c)
