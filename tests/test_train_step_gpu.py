"""One training step / a few steps of the CUDA path against the C oracle (GPU only).

Tolerances are the north star's: outputs, gradients and alpha within 1e-3 relative (Frobenius) for the
tensor-core path (bf16x3, fp32 accumulate: measured ~1e-5) and 2e-5 for the fp32 validation path."""
import numpy as np
import pytest
from conftest import rel_err

pytestmark = pytest.mark.gpu


def make_case(O, layersizes, M, seed, beta_w=2.0):
    rng = np.random.RandomState(seed)
    W, b = O.init_weights(layersizes, seed=seed, beta=beta_w)
    b = [rng.uniform(-0.1, 0.1, x.size).astype(np.float32) for x in b]
    x = rng.randn(M, layersizes[0]).astype(np.float32)
    t = rng.randn(M, layersizes[-1]).astype(np.float32)
    return W, b, x, t


def run_both(pkg, O, layersizes, M, MLflag, beta, precision, nsteps=1, seed=3, lr=0.1, mom=0.9, wc=1e-5):
    W, b, x, t = make_case(O, layersizes, M * nsteps, seed)
    orc = O.OracleNet(layersizes, M, lr, mom, wc, beta, MLflag, W, b)
    net = pkg.BP_GPU(0, 0, len(layersizes), layersizes, M, lr, mom, wc, W, b, beta, MLflag, precision=precision)
    return orc, net, x, t


SMALL = [70, 96, 80, 33]


@pytest.mark.parametrize("precision,tol", [(1, 2e-5), (0, 1e-3)])
@pytest.mark.parametrize("MLflag,beta", [(1, 1.5), (1, 1.0), (0, 2.0), (0, 1.0), (1, 2.0)])
@pytest.mark.parametrize("M", [128, 50])
def test_single_step_tensors(pkg, oracle, precision, tol, MLflag, beta, M):
    O = oracle
    orc, net, x, t = run_both(pkg, O, SMALL, M, MLflag, beta, precision)
    orc.train_bunch(x, t)
    net.debug_step(x, t, apply_update=True)
    L = len(SMALL)
    assert rel_err(net.debug_read(0), orc.out(M)) < tol
    if MLflag == 1:
        assert rel_err(net.alpha(), orc.alpha()) < tol
    for l in range(1, L):
        assert rel_err(net.debug_read(1, l), orc.dedx(l, M)) < tol, "dedx layer %d" % l
        assert rel_err(net.debug_read(3, l), orc.grad(l)) < tol, "grad layer %d" % l
        if l < L - 1:
            assert rel_err(net.debug_read(2, l), orc.y(l, M)) < tol, "y layer %d" % l
    Wg, bg = net.returnWeights()
    Wo, bo = orc.weights()
    for l in range(L - 1):
        assert rel_err(Wg[l], Wo[l]) < tol
        assert rel_err(bg[l], bo[l]) < tol
        # the UPDATE itself (not just the weights, which barely move in one step) must agree
    assert abs(net.losses()[0] - orc.last_loss()) <= 1e-3 * abs(orc.last_loss())


def test_zero_error_branch(pkg, oracle):
    """e == 0 must give exactly zero gradient (DevFunc.cu:388-391, 479-482)"""
    O = oracle
    layersizes, M = [8, 6, 5], 128
    W, b, x, t = make_case(O, layersizes, M, 11)
    for MLflag, beta in ((1, 1.5), (0, 1.0)):
        net = pkg.BP_GPU(0, 0, 3, layersizes, M, 0.1, 0.9, 0.0, W, b, beta, MLflag, precision=1)
        out = net.forward(x)
        t2 = t.copy()
        t2[::3] = out[::3]          # exact hits
        net.debug_step(x, t2, apply_update=False)
        d = net.debug_read(1, 2)
        assert np.all(d[::3] == 0.0)
        assert np.all(d[1::3] != 0.0)


@pytest.mark.parametrize("precision,tol", [(1, 5e-5), (0, 1e-3)])
def test_multi_step_chunk(pkg, oracle, precision, tol):
    """ggd_train over a chunk (graph replay, device-side bunch counter, tail bunch dropped)"""
    O = oracle
    layersizes, M, nb = SMALL, 128, 19
    W, b, x, t = make_case(O, layersizes, M * nb + 37, 5)
    orc = O.OracleNet(layersizes, M, 0.05, 0.9, 1e-5, 1.5, 1, W, b)
    net = pkg.BP_GPU(0, 0, len(layersizes), layersizes, M, 0.05, 0.9, 1e-5, W, b, 1.5, 1, precision=precision)
    lo, al = orc.train(x, t)
    net.train(x.shape[0], x, t)
    lg = net.losses()
    assert len(lg) == nb == len(lo)
    assert np.max(np.abs(lg - lo) / np.abs(lo)) < 5e-3       # loss curve within 0.5 %
    assert rel_err(net.alpha(), al[-1]) < tol
    Wg, bg = net.returnWeights()
    Wo, bo = orc.weights()
    for l in range(len(layersizes) - 1):
        assert rel_err(Wg[l], Wo[l]) < tol
        assert rel_err(bg[l], bo[l]) < 5 * tol
    # CV metrics, partial last bunch included
    xc, tc = x[:300], t[:300]
    for f in ("cv_sqerr", "cv_abserr", "cv_loglik"):
        ref = getattr(orc, f)(xc, tc)
        got = {"cv_sqerr": net.CrossValid, "cv_abserr": net.CrossValiddB, "cv_loglik": net.CrossValid2}[f](300, xc, tc)
        assert abs(got - ref) <= 2e-3 * abs(ref), (f, got, ref)


def test_named_shape_step(pkg, oracle):
    """the named configuration: 1799-2048-2048-2048-257, bunch 128, MLflag=1, beta=1.5"""
    O = oracle
    layersizes, M = [1799, 2048, 2048, 2048, 257], 128
    W, b, x, t = make_case(O, layersizes, 2 * M, 9)
    orc = O.OracleNet(layersizes, M, 0.1, 0.9, 1e-5, 1.5, 1, W, b)
    net = pkg.BP_GPU(0, 0, 5, layersizes, M, 0.1, 0.9, 1e-5, W, b, 1.5, 1)
    orc.train_bunch(x[:M], t[:M])
    net.debug_step(x[:M], t[:M], apply_update=True)
    assert rel_err(net.debug_read(0), orc.out(M)) < 1e-3
    assert rel_err(net.alpha(), orc.alpha()) < 1e-3
    for l in range(1, 5):
        assert rel_err(net.debug_read(3, l), orc.grad(l)) < 1e-3, l
    # second step sees the updated weights
    orc.train_bunch(x[M:], t[M:])
    net.debug_step(x[M:], t[M:], apply_update=True)
    assert rel_err(net.debug_read(0), orc.out(M)) < 1e-3
    Wg, _ = net.returnWeights()
    Wo, _ = orc.weights()
    for l in range(4):
        assert rel_err(Wg[l], Wo[l]) < 1e-3


def test_fused_update_equals_unfused(pkg, oracle):
    """the fused gradient-GEMM + update kernel (TMEM -> smem -> TMA store) against the materialised-gradient path"""
    O = oracle
    layersizes, M, nb = [7 * 33, 200, 130, 33], 128, 9     # ragged: Kp = 256/256/192, not multiples of 128 everywhere
    W, b, x, t = make_case(O, layersizes, M * nb, 17)
    res = []
    for flags in (0, pkg.FLAG_UNFUSED_UPDATE):
        net = pkg.BP_GPU(0, 0, len(layersizes), layersizes, M, 0.1, 0.9, 1e-5, W, b, 1.5, 1, flags=flags)
        net.train(x.shape[0], x, t)
        res.append((net.returnWeights(), net.alpha(), net.losses()))
        net.close()
    (Wa, ba), aa, la = res[0]
    (Wb, bb), ab, lb = res[1]
    for u, v in zip(Wa + ba, Wb + bb):
        assert rel_err(u, v) < 1e-6
    assert rel_err(aa, ab) < 1e-6 and np.allclose(la, lb, rtol=1e-6)
    orc = O.OracleNet(layersizes, M, 0.1, 0.9, 1e-5, 1.5, 1, W, b)
    orc.train(x, t)
    for u, v in zip(Wa + ba, orc.weights()[0] + orc.weights()[1]):
        assert rel_err(u, v) < 1e-3


def test_named_shape_chunk_persistent(pkg, oracle):
    """ggd_train at the named shape: the production path (graph replay, persistent gradient+update launch for all
    layers, biases and the bunch counter) over several bunches against the C oracle"""
    O = oracle
    layersizes, M, nb = [1799, 2048, 2048, 2048, 257], 128, 5
    W, b, x, t = make_case(O, layersizes, M * nb, 21)
    orc = O.OracleNet(layersizes, M, 0.1, 0.9, 1e-5, 1.5, 1, W, b)
    lo, al = orc.train(x, t)
    net = pkg.BP_GPU(0, 0, 5, layersizes, M, 0.1, 0.9, 1e-5, W, b, 1.5, 1)
    net.train(x.shape[0], x, t)
    lg = net.losses()
    assert len(lg) == nb
    assert np.max(np.abs(lg - lo) / np.abs(lo)) < 5e-3
    assert rel_err(net.alpha(), al[-1]) < 1e-3
    Wg, bg = net.returnWeights()
    Wo, bo = orc.weights()
    for l in range(4):
        assert rel_err(Wg[l], Wo[l]) < 1e-3
        assert rel_err(bg[l], bo[l]) < 1e-3
        # the UPDATE itself must agree, not only the (barely moved) weights
        assert rel_err(Wg[l] - W[l], Wo[l] - W[l]) < 2e-3, l
        assert rel_err(bg[l] - b[l], bo[l] - b[l]) < 2e-3, l


@pytest.mark.parametrize("M,MLflag,beta", [(50, 1, 1.5), (100, 0, 2.0), (128, 1, 1.0)])
def test_chunk_ragged_bunch_production_path(pkg, oracle, M, MLflag, beta):
    """ggd_train with bunches below the 128-frame tile (rows >= bunch are padding) and ragged layer widths: the fused
    loss epilogue and the persistent gradient+update kernel must ignore the padding exactly"""
    O = oracle
    layersizes, nb = [7 * 33, 200, 130, 33], 7
    W, b, x, t = make_case(O, layersizes, M * nb + 11, 31)
    orc = O.OracleNet(layersizes, M, 0.1, 0.9, 1e-5, beta, MLflag, W, b)
    lo, al = orc.train(x, t)
    net = pkg.BP_GPU(0, 0, len(layersizes), layersizes, M, 0.1, 0.9, 1e-5, W, b, beta, MLflag)
    net.train(x.shape[0], x, t)
    lg = net.losses()
    assert len(lg) == nb
    assert np.max(np.abs(lg - lo) / np.abs(lo)) < 5e-3
    if MLflag == 1:
        assert rel_err(net.alpha(), al[-1]) < 1e-3
    Wg, bg = net.returnWeights()
    Wo, bo = orc.weights()
    for l in range(len(layersizes) - 1):
        assert rel_err(Wg[l], Wo[l]) < 1e-3
        assert rel_err(bg[l], bo[l]) < 1e-3
        assert rel_err(Wg[l] - W[l], Wo[l] - W[l]) < 2e-3, l


def test_device_loader_equals_host_loader(pkg, oracle):
    """ggd_train_raw (byte swap, z-score, context expansion, target selection and operand split on the GPU
    from the raw pfile records) must give BIT-identical training to ggd_train on the arrays expanded by the restated host
    loader, chunk by chunk, on the reference's bundled pfiles with finetune.pl's settings (ctx 7, offset 3, seed 27870775)"""
    import os
    from conftest import GOLDEN
    O = oracle
    fea, tg, nrm = (os.path.join(GOLDEN, f) for f in ("train_noisy.pfile", "train_clean.pfile", "train_noisy.norm"))
    layersizes, M, cache = [7 * 257, 96, 257], 128, 600
    W, b = O.init_weights(layersizes, seed=4)
    nets = [pkg.BP_GPU(0, 0, 3, layersizes, M, 0.1, 0.9, 1e-5, W, b, 1.5, 1) for _ in range(2)]
    loaders = [O.PfileLoader(fea, tg, nrm, 257, 7, 3, cache, 27870775) for _ in range(2)]
    sent_en = len(loaders[0].sent_end) - 1
    starts, total = loaders[0].chunk_info(0, sent_en)
    assert len(starts) >= 2
    for idx in range(len(starts)):
        ind, tgt = loaders[0].read_chunk(starts, total, sent_en, idx)
        frec, trec, first = loaders[1].read_chunk_raw(starts, total, sent_en, idx)
        assert first.size == ind.shape[0]
        nets[0].train(ind.shape[0], ind, tgt)
        nets[1].train_raw(frec, trec, first, 257, 7, 3, loaders[1].mean, loaders[1].dvar)
        assert np.array_equal(nets[0].losses(), nets[1].losses()), idx
    (Wa, ba), (Wb, bb) = nets[0].returnWeights(), nets[1].returnWeights()
    for u, v in zip(Wa + ba, Wb + bb):
        assert np.array_equal(u, v)
    assert np.array_equal(nets[0].alpha(), nets[1].alpha())
    # argument checking: a row that would read past the chunk is refused
    with pytest.raises(Exception):
        nets[1].train_raw(frec, trec, np.array([frec.shape[0] - 3], np.int32), 257, 7, 3, loaders[1].mean, loaders[1].dvar)


def test_reused_host_buffer_growing_chunks(pkg, oracle):
    """GGD_FLAG_PIN_HOST with a REUSED host buffer whose chunks grow (an epoch's chunks differ in size and come in shuffled
    order: BPtrain.cc reuses indata/targ): the pinned range must follow, and the piece-wise upload must stay correct"""
    O = oracle
    layersizes, M = [300, 64, 40], 128
    W, b, x, t = make_case(O, layersizes, M * 150, 41)     # 300*4*128*150 = 23 MB: above the 1 MB pinning threshold
    buf_x = np.zeros_like(x); buf_t = np.zeros_like(t)
    net = pkg.BP_GPU(0, 0, 3, layersizes, M, 0.05, 0.9, 1e-5, W, b, 1.5, 1, flags=pkg.FLAG_PIN_HOST)
    net.keep_pinned(buf_x, buf_t)                            # long-lived: the registration survives across calls
    orc = O.OracleNet(layersizes, M, 0.05, 0.9, 1e-5, 1.5, 1, W, b)
    pos = 0
    for nb in (20, 130):                                    # small chunk first, then a much larger one in the same buffer
        n = nb * M
        buf_x[:n] = x[pos:pos + n]; buf_t[:n] = t[pos:pos + n]
        net.train(n, buf_x[:n], buf_t[:n])
        lo, _ = orc.train(x[pos:pos + n], t[pos:pos + n])
        assert np.max(np.abs(net.losses() - lo) / np.abs(lo)) < 5e-3
        pos += n
    Wg, _ = net.returnWeights()
    Wo, _ = orc.weights()
    for l in range(2):
        assert rel_err(Wg[l], Wo[l]) < 1e-3


@pytest.mark.parametrize("M", [256, 512, 1024, 200])
@pytest.mark.parametrize("MLflag,beta", [(1, 1.5), (0, 2.0)])
def test_wide_bunch_sizes(pkg, oracle, M, MLflag, beta):
    """bunches above 128 frames: forward / backward over several row tiles, stand-alone loss kernel, and the WIDE persistent
    gradient+update kernel (dw_wide.cu: frames streamed in 32-frame blocks, 64..256-unit segments) + bias_wide_kernel,
    over several steps against the oracle (BP_GPU.cu:170-184 for any bunch size)"""
    O = oracle
    layersizes, nb = [7 * 33, 200, 330, 33], 4     # ragged: Kp = 256 / 256 / 384 (segments of 4 and 2 slabs), Np = 256 / 384 / 64
    W, b, x, t = make_case(O, layersizes, M * nb, 23)
    orc = O.OracleNet(layersizes, M, 0.1, 0.9, 1e-5, beta, MLflag, W, b)
    net = pkg.BP_GPU(0, 0, len(layersizes), layersizes, M, 0.1, 0.9, 1e-5, W, b, beta, MLflag)
    lo, al = orc.train(x, t)
    net.train(x.shape[0], x, t)
    assert np.max(np.abs(net.losses() - lo) / np.abs(lo)) < 5e-3
    if MLflag == 1:
        assert rel_err(net.alpha(), al[-1]) < 1e-3
    Wg, bg = net.returnWeights()
    Wo, bo = orc.weights()
    for l in range(len(layersizes) - 1):
        assert rel_err(Wg[l], Wo[l]) < 1e-3
        assert rel_err(Wg[l] - W[l], Wo[l] - W[l]) < 2e-3, "update of layer %d" % (l + 1)
        assert rel_err(bg[l], bo[l]) < 2e-3
    net.close()


def test_wide_kernel_at_128_equals_persist(pkg, oracle, monkeypatch):
    """GGD_DW_PERSIST=0 routes a 128-frame bunch through dw_wide.cu: same weights as dw_persist.cu (same products, fp32
    accumulation in a different order)"""
    O = oracle
    layersizes, M, nb = [7 * 33, 200, 130, 33], 128, 6
    W, b, x, t = make_case(O, layersizes, M * nb, 29)
    res = []
    for env in ("1", "0"):
        monkeypatch.setenv("GGD_DW_PERSIST", env)
        net = pkg.BP_GPU(0, 0, len(layersizes), layersizes, M, 0.1, 0.9, 1e-5, W, b, 1.5, 1)
        net.train(x.shape[0], x, t)
        res.append((net.returnWeights(), net.alpha(), net.losses()))
        net.close()
    (Wa, ba), aa, la = res[0]
    (Wb, bb), ab, lb = res[1]
    for u, v in zip(Wa + ba, Wb + bb):
        assert rel_err(u, v) < 1e-5
    assert rel_err(aa, ab) < 1e-5 and np.allclose(la, lb, rtol=1e-5)


@pytest.mark.parametrize("M", [1024, 256])
def test_config4_shape_single_gpu(pkg, oracle, M):
    """BASELINE config 4's network (ctx 11: 2827-2048^3-257) on ONE GPU with the whole global minibatch, against the oracle"""
    O = oracle
    layersizes = [2827, 2048, 2048, 2048, 257]
    W, b, x, t = make_case(O, layersizes, 2 * M, 31)
    orc = O.OracleNet(layersizes, M, 0.1, 0.9, 1e-5, 1.5, 1, W, b)
    net = pkg.BP_GPU(0, 0, 5, layersizes, M, 0.1, 0.9, 1e-5, W, b, 1.5, 1)
    lo, al = orc.train(x, t)
    net.train(x.shape[0], x, t)
    assert np.max(np.abs(net.losses() - lo) / np.abs(lo)) < 5e-3
    assert rel_err(net.alpha(), al[-1]) < 1e-3
    Wg, bg = net.returnWeights()
    Wo, bo = orc.weights()
    for l in range(4):
        assert rel_err(Wg[l], Wo[l]) < 1e-3
        assert rel_err(Wg[l] - W[l], Wo[l] - W[l]) < 2e-3, "update of layer %d" % (l + 1)
        assert rel_err(bg[l] - b[l], bo[l] - b[l]) < 2e-3
    net.close()


@pytest.mark.parametrize("precision,tol", [(1, 2e-5), (0, 1e-3)])
def test_enhance_inference_path(pkg, oracle, precision, tol):
    """ggd_enhance = Test_code/decode.m for one utterance (z-score, edge-replicated context of frame_expand.m, forward,
    de-normalisation) against the float64 numpy restatement; utterances shorter than the context and longer than a bunch"""
    O = oracle
    layersizes, M = [7 * 33, 96, 80, 33], 128
    W, b, _, _ = make_case(O, layersizes, M, 37)
    rng = np.random.RandomState(3)
    mean = rng.randn(33).astype(np.float32); dvar = rng.uniform(0.5, 2.0, 33).astype(np.float32)
    net = pkg.BP_GPU(0, 0, len(layersizes), layersizes, M, 0.1, 0.9, 1e-5, W, b, 1.5, 1, precision=precision)
    for T in (1, 2, 5, 128, 333):
        lps = (rng.randn(T, 33) * 3 + 10).astype(np.float32)
        got = net.enhance(lps, mean, dvar, 7)
        ref = O.enhance_ref(lps, W, b, layersizes, mean, dvar, 7)
        assert got.shape == ref.shape == (T, 33)
        assert rel_err(got, ref) < tol, T
    with pytest.raises(Exception):
        net.enhance(np.zeros((4, 33), np.float32), mean, dvar, 5)      # 5 x 33 != layersizes[0]


def test_cv_all_equals_three_calls(pkg, oracle):
    O = oracle
    layersizes, M = SMALL, 128
    W, b, x, t = make_case(O, layersizes, 300, 43)
    net = pkg.BP_GPU(0, 0, len(layersizes), layersizes, M, 0.1, 0.9, 1e-5, W, b, 1.5, 1)
    net.train(256, x[:256], t[:256])            # sets alpha for CrossValid2
    a = net.CrossValidAll(300, x, t)
    assert a == (net.CrossValid(300, x, t), net.CrossValiddB(300, x, t), net.CrossValid2(300, x, t))
