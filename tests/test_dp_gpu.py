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
SMALL = [70, 96, 80, 33]
NAMED = [1799, 2048, 2048, 2048, 257]
CONFIG4 = [2827, 2048, 2048, 2048, 257]          # BASELINE.json configs[3]: ctx 11, global minibatch 1024


def _data(world, LS, MS, NB):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    rng = np.random.RandomState(13)
    W, b = O.init_weights(LS, seed=3)
    x = rng.randn(world, NB * MS, LS[0]).astype(np.float32)     # [rank][frames][dim]
    t = rng.randn(world, NB * MS, LS[-1]).astype(np.float32)
    return W, b, x, t


def _rank(rank, world, uid, ml, beta, precision, q, LS, MS, NB):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_pkg
    pkg = load_pkg()
    W, b, x, t = _data(world, LS, MS, NB)
    net = pkg.BP_GPU(0, rank, len(LS), LS, MS, 0.1, 0.9, 1e-5, W, b, beta, ml, precision=precision, world_size=world, rank=rank,
                     nccl_unique_id=uid)
    net.train(NB * MS, x[rank], t[rank])
    Wn, bn = net.returnWeights()
    q.put((rank, [w.copy() for w in Wn], [v.copy() for v in bn], net.alpha(), net.losses()))
    net.close()


def run_dp(world, LS, MS, NB, ml, beta, precision, timeout=400):
    """`world` processes (one per GPU) train NB global minibatches of world*MS frames; returns their results and the
    unsharded oracle's.  Global minibatch i = rank 0's bunch i, then rank 1's, ...."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from conftest import load_pkg
    load_pkg()
    from se_ml_b200.bp_gpu import nccl_unique_id
    uid = nccl_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_rank, args=(r, world, uid, ml, beta, precision, q, LS, MS, NB)) for r in range(world)]
    for p in ps:
        p.start()
    res = _collect(q, ps, timeout)
    for p in ps:
        p.join(60)
    W, b, x, t = _data(world, LS, MS, NB)
    xg = np.concatenate([x[:, i * MS:(i + 1) * MS].reshape(world * MS, -1) for i in range(NB)])
    tg = np.concatenate([t[:, i * MS:(i + 1) * MS].reshape(world * MS, -1) for i in range(NB)])
    orc = O.OracleNet(LS, world * MS, 0.1, 0.9, 1e-5, beta, ml, W, b)
    lo, al = orc.train(xg, tg)
    Wo, bo = orc.weights()
    return res, (W, b), (Wo, bo, al, lo)


def check_dp(res, init, ref, ml, tol, check_update=False):
    W, b = init
    Wo, bo, al, lo = ref
    for r in res:
        for a, c, w0 in zip(r[1] + r[2], Wo + bo, W + b):
            assert np.linalg.norm(a - c) <= tol * max(np.linalg.norm(c), 1e-6)
            if check_update:   # the update itself, not only the weights (which barely move in a few steps)
                assert np.linalg.norm((a - w0) - (c - w0)) <= 2 * tol * max(np.linalg.norm(c - w0), 1e-6)
        if ml:
            assert np.linalg.norm(r[3] - al[-1]) <= tol * np.linalg.norm(al[-1])     # alpha of the GLOBAL minibatch
        assert np.allclose(r[4], lo, rtol=5e-3)
    # every rank holds bit-identical weights and biases
    for r in res[1:]:
        for a, c in zip(res[0][1] + res[0][2], r[1] + r[2]):
            assert np.array_equal(a, c)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("ml,beta,precision", [(1, 1.5, 0), (1, 1.5, 1), (0, 2.0, 0)])
def test_dp_equals_unsharded(ml, beta, precision, world):
    """small ragged net, 128 frames per GPU: tensor path = factor exchange over NVLink peer memory (dp_factor.cuh),
    fp32 validation path = NCCL allreduce of sum|e|^beta and of the gradients"""
    res, init, ref = run_dp(world, SMALL, 128, 6, ml, beta, precision, 240)
    check_dp(res, init, ref, ml, 1e-3 if precision == 0 else 1e-4)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dp_named_shape(world):
    """the named network, 128 frames per GPU (global minibatch 128*world), against the oracle on the unsharded minibatch"""
    res, init, ref = run_dp(world, NAMED, 128, 3, 1, 1.5, 0)
    check_dp(res, init, ref, 1, 1e-3, check_update=True)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dp_config4(world):
    """BASELINE config 4: 2827-2048^3-257, global minibatch 1024 sharded as 1024/world frames per GPU"""
    res, init, ref = run_dp(world, CONFIG4, 1024 // world, 2, 1, 1.5, 0, 600)
    check_dp(res, init, ref, 1, 1e-3, check_update=True)


def test_dp_ragged_bunch():
    """bunch not a multiple of 128 (padding rows of the factor arena must stay zero) and MLflag=0, beta=1"""
    res, init, ref = run_dp(2, SMALL, 100, 5, 0, 1.0, 0, 240)
    check_dp(res, init, ref, 0, 1e-3)
