"""CPU tests: the oracle against the reference's golden vectors / known answers / the reference's own code."""
import ctypes as C
import math
import os
import numpy as np
import pytest
from conftest import GOLDEN, PKG_DIR, rel_err

NAMES = ("TEST_DR8_MPAM0_SX289", "TEST_DR8_MPAM0_SX379")
SEED = 27870775      # finetune.pl:31


# ---------------------------------------------------------------- LPS
@pytest.mark.parametrize("name,frames,samples", [(NAMES[0], 168, 43264), (NAMES[1], 156, 40192)])
def test_lps_oracle_bit_exact_on_reference_goldens(oracle, name, frames, samples):
    pcm = oracle.read_wav_pcm16(os.path.join(GOLDEN, name + ".wav"))
    hdr, gold = oracle.read_htk(os.path.join(GOLDEN, name + ".lps"))
    assert len(pcm) == samples and hdr == dict(nSamples=frames, sampPeriod=160000, sampSize=1028, parmKind=9)
    mine = oracle.lps_extract(pcm)
    assert mine.shape == (frames, 257)
    assert np.array_equal(mine.view(np.uint32), gold.view(np.uint32))


def test_lps_oracle_vs_reference_binary(oracle):
    from oracle import refcuda
    if not refcuda.available("Wav2LPS_be_ref"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.RandomState(1234)
    pcm = np.clip(np.round(rng.randn(16000 * 5) * 3000), -32768, 32767).astype(np.int16)
    ref, _ = refcuda.ref_wav2lps(pcm)
    assert np.array_equal(oracle.lps_extract(pcm).view(np.uint32), ref.view(np.uint32))


def test_lps_frame_count_and_floor(oracle):
    L = oracle.lib()
    assert [L.lps_oracle_nframes(n) for n in (0, 511, 512, 767, 768, 43264)] == [0, 0, 1, 1, 2, 168]
    z = oracle.lps_extract(np.zeros(1024, np.int16))
    assert z.shape == (3, 257) and np.all(z == -50.0)


def test_rfft_matches_numpy(oracle):
    rng = np.random.RandomState(0)
    x = rng.randn(512).astype(np.float32)
    y = x.copy()
    oracle.lib().lps_oracle_rfft(y.ctypes.data_as(oracle.PF), 512, 9)
    f = np.fft.rfft(x.astype(np.float64))
    assert np.allclose(y[:257], f.real, atol=2e-4)
    assert np.allclose(y[:256:-1], -f.imag[1:256], atol=2e-4) or np.allclose(y[:256:-1], f.imag[1:256], atol=2e-4)


# ---------------------------------------------------------------- loader
def _loader(oracle):
    return oracle.PfileLoader(os.path.join(GOLDEN, "train_noisy.pfile"), os.path.join(GOLDEN, "train_clean.pfile"),
                              os.path.join(GOLDEN, "train_noisy.norm"), 257, 7, 3, 102400, SEED)


def test_loader_known_answers(oracle):
    """numbers obtained with the reference's own Interface.cc on the bundled pfiles"""
    feats, ends, _ = oracle.read_pfile(os.path.join(GOLDEN, "train_noisy.pfile"))
    assert feats.shape == (1885, 257)
    assert list(np.diff(np.concatenate([[0], ends]))) == [146, 143, 247, 227, 168, 177, 192, 191, 190, 204]
    ld = _loader(oracle)
    st, tot = ld.chunk_info(0, 7)
    assert (len(st), tot) == (1, 1443)
    cst, ctot = ld.chunk_info(8, 9)
    assert (len(cst), ctot) == (1, 382)
    x, t = ld.read_chunk(st, tot, 7, 0)
    d = np.float64
    assert abs(x.astype(d).sum() - 362738.279) < 0.01 and abs((x.astype(d) ** 2).sum() - 2543536.044) < 0.05
    assert abs(t.astype(d).sum() - (-182261.062)) < 0.01 and abs((t.astype(d) ** 2).sum() - 451292.635) < 0.05
    assert np.allclose(x[0, :4], [-1.032592, -0.154630, 0.140260, 0.258554], atol=1e-6)
    assert abs(x[0, 771] - (-0.389783)) < 1e-6 and abs(t[0, 0] - (-0.555728)) < 1e-6


def test_norm_file_is_mean_and_reciprocal_std(oracle):
    feats, _, _ = oracle.read_pfile(os.path.join(GOLDEN, "train_noisy.pfile"))
    mean, dvar = oracle.read_norm(os.path.join(GOLDEN, "train_noisy.norm"), 257)
    assert np.allclose(mean, feats.astype(np.float64).mean(0), rtol=1e-5, atol=1e-5)
    assert np.allclose(dvar, 1.0 / feats.astype(np.float64).std(0), rtol=1e-4)


def _host_lib():
    so = os.path.join(PKG_DIR, "host", "libbphost.so")
    if not os.path.exists(so):
        pytest.skip("host/libbphost.so not built")
    L = C.CDLL(so)
    L.bph_create.restype = C.c_void_p
    L.bph_create.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.bph_destroy.argtypes = [C.c_void_p]
    L.bph_chunk_info.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.bph_shuffle.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
    L.bph_read_chunk.argtypes = [C.c_void_p, C.c_int, C.c_int, oracle_PF, oracle_PF]
    L.bph_read_chunk_raw.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.bph_read_chunk_raw_slice.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3
    L.bph_weights.restype = oracle_PF
    L.bph_weights.argtypes = [C.c_void_p, C.c_int]
    L.bph_write_weights.argtypes = [C.c_void_p]
    return L


oracle_PF = C.POINTER(C.c_float)


def _flags(tmp_path, init, **over):
    kw = dict(gpu_used=0, numlayers=4, layersizes="1799,32,16,257", bunchsize=128, MLflag=1, shapefactor=1.5, momentum=0.9,
              weightcost=0.00001, lrate=0.1, fea_dim=257, fea_context=7, traincache=102400, init_randem_seed=SEED, targ_offset=3,
              initwts_file=init, norm_file=os.path.join(GOLDEN, "train_noisy.norm"), fea_file=os.path.join(GOLDEN, "train_noisy.pfile"),
              targ_file=os.path.join(GOLDEN, "train_clean.pfile"), outwts_file=str(tmp_path / "out.wts"), log_file=str(tmp_path / "x.log"),
              train_sent_range="0-7", cv_sent_range="8-9", dropoutflag=0, visible_omit=0.1, hid_omit=0.1)
    kw.update(over)
    return kw


def _argv(kw):
    args = [b"BPtrain_Sigmoid"] + [("%s=%s" % (k, v)).encode() for k, v in kw.items()]
    return len(args), (C.c_char_p * len(args))(*args)


@pytest.mark.parametrize("traincache", [102400, 500])
def test_product_host_loader_matches_restatement_and_reference(oracle, tmp_path, traincache):
    """host/interface.cpp (product) vs the numpy restatement vs the reference's Interface.cc, bit for bit,
    including multi-chunk splitting (traincache=500) and the lrand48 chunk/sample shuffles"""
    L = _host_lib()
    ls = [1799, 32, 16, 257]
    W, b = oracle.init_weights(ls, seed=3)
    init = str(tmp_path / "init.wts")
    oracle.write_wts(init, ls, W, b)
    kw = _flags(tmp_path, init, traincache=traincache)
    h = L.bph_create(*_argv(kw))
    assert h
    for l in range(1, 4):
        assert np.array_equal(np.ctypeslib.as_array(L.bph_weights(h, l), shape=(W[l - 1].size,)), W[l - 1])
    nch, ns = C.c_int(), C.c_int()
    assert L.bph_chunk_info(h, b"0-7", 0, C.byref(nch), C.byref(ns)) == 0
    ld = oracle.PfileLoader(kw["fea_file"], kw["targ_file"], kw["norm_file"], 257, 7, 3, traincache, SEED)
    st, tot = ld.chunk_info(0, 7)
    assert (nch.value, ns.value) == (len(st), tot)
    order_np = oracle.rand_index(list(range(len(st))), ld.rng)
    order = (C.c_int * len(st))(*range(len(st)))
    L.bph_shuffle(h, order, len(st))
    assert list(order) == order_np
    ref = None
    from oracle import refcuda
    if refcuda.available("libref_interface.so"):
        kw2 = dict(kw); kw2["outwts_file"] = str(tmp_path / "ref.wts"); kw2["log_file"] = str(tmp_path / "ref.log")
        ref = refcuda.RefInterface(**kw2)
        assert ref.train_info("0-7") == (len(st), tot)
        assert ref.shuffle_chunks(len(st)) == order_np
    xin = np.zeros((traincache, 1799), np.float32); xt = np.zeros((traincache, 257), np.float32)
    for ci in order_np:
        n = L.bph_read_chunk(h, ci, 0, xin.ctypes.data_as(oracle_PF), xt.ctypes.data_as(oracle_PF))
        x, t = ld.read_chunk(st, tot, 7, ci)
        assert n == x.shape[0]
        assert np.array_equal(xin[:n].view(np.uint32), x.view(np.uint32)) and np.array_equal(xt[:n].view(np.uint32), t.view(np.uint32))
        if ref is not None:
            xr, tr = ref.read_chunk(ci, 1799, 257)
            assert xr.shape[0] == n and np.array_equal(xr.view(np.uint32), x.view(np.uint32)) and np.array_equal(tr.view(np.uint32), t.view(np.uint32))
    # CV chunks: no shuffle
    assert L.bph_chunk_info(h, b"8-9", 1, C.byref(nch), C.byref(ns)) == 0
    cst, ctot = ld.chunk_info(8, 9)
    assert (nch.value, ns.value) == (len(cst), ctot)
    n = L.bph_read_chunk(h, 0, 1, xin.ctypes.data_as(oracle_PF), xt.ctypes.data_as(oracle_PF))
    x, t = ld.read_chunk(cst, ctot, 9, 0, shuffle=False)
    assert n == x.shape[0] and np.array_equal(xin[:n], x) and np.array_equal(xt[:n], t)
    # weight file written by the product equals the oracle's writer (MAT-v4 layout, Interface.cc:489-514)
    assert L.bph_write_weights(h) == 0
    L.bph_destroy(h)
    oracle.write_wts(str(tmp_path / "chk.wts"), ls, W, b)
    assert open(kw["outwts_file"], "rb").read() == open(str(tmp_path / "chk.wts"), "rb").read()


def test_host_loader_error_paths(oracle, tmp_path):
    L = _host_lib()
    ls = [1799, 32, 16, 257]
    W, b = oracle.init_weights(ls, seed=3)
    init = str(tmp_path / "init.wts")
    oracle.write_wts(init, ls, W, b)
    assert not L.bph_create(*_argv(_flags(tmp_path, init, layersizes="1799,32,17,257")))      # node counts do not match
    assert "init weights node nums do not match" in open(str(tmp_path / "x.log")).read()
    assert not L.bph_create(*_argv(_flags(tmp_path, init, fea_context=5)))                      # 257*5 != 1799
    assert not L.bph_create(*_argv(_flags(tmp_path, init, fea_file="/nonexistent")))
    h = L.bph_create(*_argv(_flags(tmp_path, init)))
    a, c = C.c_int(), C.c_int()
    assert L.bph_chunk_info(h, b"3-12", 0, C.byref(a), C.byref(c)) != 0                         # sentence 12 does not exist
    assert L.bph_chunk_info(h, b"5", 0, C.byref(a), C.byref(c)) != 0                            # format error
    L.bph_destroy(h)


def test_file_format_round_trips(oracle, tmp_path):
    rng = np.random.RandomState(2)
    feats = rng.randn(57, 5).astype(np.float32)
    oracle.write_pfile(str(tmp_path / "a.pfile"), feats, [20, 30, 7])
    f2, ends, sid = oracle.read_pfile(str(tmp_path / "a.pfile"))
    assert np.array_equal(f2, feats) and list(ends) == [20, 50, 57] and list(sid[[0, 19, 20, 56]]) == [0, 0, 1, 2]
    ls = [6, 4, 3]
    W, b = oracle.init_weights(ls, seed=1)
    oracle.write_wts(str(tmp_path / "w.wts"), ls, W, b)
    W2, b2 = oracle.read_wts(str(tmp_path / "w.wts"), ls)
    assert all(np.array_equal(x, y) for x, y in zip(W + b, W2 + b2))
    raw = open(str(tmp_path / "w.wts"), "rb").read()
    assert np.frombuffer(raw[:20], "<i4").tolist() == [10, 4, 6, 0, 10] and raw[20:30] == b"weights12\0"
    oracle.write_htk(str(tmp_path / "x.lps"), feats[:, :3])
    hdr, f3 = oracle.read_htk(str(tmp_path / "x.lps"))
    assert hdr["nSamples"] == 57 and hdr["sampSize"] == 12 and np.array_equal(f3, feats[:, :3])


def test_rand48_matches_libc():
    from oracle.oracle import Rand48
    libc = C.CDLL("libc.so.6")
    libc.lrand48.restype = C.c_long
    libc.srand48(C.c_long(SEED))
    r = Rand48(SEED)
    assert [r.lrand48() for _ in range(1000)] == [libc.lrand48() for _ in range(1000)]


# ---------------------------------------------------------------- training step
def _numpy_step(ls, W, b, x, t, beta, ml):
    """independent float64 restatement of forward + loss gradient (for a finite-difference-free check of the oracle)"""
    M = x.shape[0]
    ys = [x.astype(np.float64)]
    for l in range(1, len(ls)):
        Wm = W[l - 1].astype(np.float64).reshape(ls[l - 1], ls[l])
        z = ys[-1] @ Wm + b[l - 1]
        ys.append(z if l == len(ls) - 1 else 1.0 / (1.0 + np.exp(-z)))
    e = ys[-1] - t
    s = (np.abs(e) ** beta).sum(0)
    alpha = (beta * s / M) ** (1.0 / beta)
    if ml:
        d = np.sign(e) * np.abs(e) ** (beta - 1) * beta / alpha ** beta / M
        loss = np.log(alpha).sum() + (np.abs(e / alpha) ** beta).sum() / M
    else:
        d = beta * np.sign(e) * np.abs(e) ** (beta - 1) / M
        loss = (np.abs(e) ** beta).sum() / M
    grads = []
    for l in range(len(ls) - 1, 0, -1):
        grads.append(ys[l - 1].T @ d)
        if l > 1:
            Wm = W[l - 1].astype(np.float64).reshape(ls[l - 1], ls[l])
            d = (d @ Wm.T) * ys[l - 1] * (1 - ys[l - 1])
    return ys[-1], alpha, loss, grads[::-1]


@pytest.mark.parametrize("ml,beta", [(1, 1.5), (1, 1.0), (0, 2.0), (0, 1.3)])
def test_oracle_step_against_float64_math(oracle, ml, beta):
    ls, M = [20, 16, 12, 9], 32
    rng = np.random.RandomState(4)
    W, b = oracle.init_weights(ls, seed=5)
    b = [rng.uniform(-0.2, 0.2, v.size).astype(np.float32) for v in b]
    x = rng.randn(M, ls[0]).astype(np.float32); t = rng.randn(M, ls[-1]).astype(np.float32)
    out, alpha, loss, grads = _numpy_step(ls, W, b, x, t, beta, ml)
    net = oracle.OracleNet(ls, M, 0.1, 0.9, 1e-5, beta, ml, W, b)
    net.train_bunch(x, t)
    assert rel_err(net.out(M), out) < 1e-6
    if ml:
        assert rel_err(net.alpha(), alpha) < 1e-6
    assert abs(net.last_loss() - loss) < 1e-5 * abs(loss)
    for l in range(1, len(ls)):
        assert rel_err(net.grad(l), grads[l - 1].reshape(-1)) < 2e-5
    # update rule incl. the reference's 1/M^2 net scaling (dedx carries 1/M, kernUpdatedelta divides by n again)
    Wn, bn = net.weights()
    for l in range(1, len(ls)):
        g = grads[l - 1].reshape(-1)
        want = W[l - 1] + (-0.1 * (g / M + 1e-5 * W[l - 1]))
        assert rel_err(Wn[l - 1] - W[l - 1], want - W[l - 1]) < 2e-3   # difference of float32 neighbours


def test_oracle_momentum_and_tail_bunch(oracle):
    ls, M = [10, 8, 5], 16
    rng = np.random.RandomState(8)
    W, b = oracle.init_weights(ls, seed=2)
    x = rng.randn(3 * M + 5, ls[0]).astype(np.float32); t = rng.randn(3 * M + 5, ls[-1]).astype(np.float32)
    net = oracle.OracleNet(ls, M, 0.1, 0.9, 0.0, 2.0, 0, W, b)
    losses, _ = net.train(x, t)
    assert len(losses) == 3                       # tail of 5 frames dropped (BP_GPU.cu:173-180)
    net2 = oracle.OracleNet(ls, M, 0.1, 0.9, 0.0, 2.0, 0, W, b)
    for i in range(3):
        net2.train_bunch(x[i * M:(i + 1) * M], t[i * M:(i + 1) * M])
    assert all(np.array_equal(a, c) for a, c in zip(net.weights()[0], net2.weights()[0]))
    # CV processes the partial bunch (BP_GPU.cu:203-218)
    assert net.forward(x).shape == (3 * M + 5, 5)


def test_oracle_sharded_equals_unsharded(oracle):
    ls, Mg = [24, 20, 11], 64
    rng = np.random.RandomState(6)
    W, b = oracle.init_weights(ls, seed=7)
    x = rng.randn(Mg, ls[0]).astype(np.float32); t = rng.randn(Mg, ls[-1]).astype(np.float32)
    for ml, beta in ((1, 1.5), (0, 2.0)):
        a = oracle.OracleNet(ls, Mg, 0.1, 0.9, 1e-5, beta, ml, W, b)
        a.train_bunch(x, t)
        for world in (2, 4, 8):
            s = oracle.OracleNet(ls, Mg // world, 0.1, 0.9, 1e-5, beta, ml, W, b)
            s.train_bunch_sharded(world, x, t)
            if ml:
                assert rel_err(s.alpha(), a.alpha()) < 1e-6
            for u, v in zip(s.weights()[0] + s.weights()[1], a.weights()[0] + a.weights()[1]):
                assert rel_err(u, v) < 1e-6


def test_gamma_polynomial(oracle):
    L = oracle.lib()
    for xv in (0.5, 2.0 / 3.0, 1.0, 1.5, 2.5, 3.7, 6.2):
        assert abs(L.ggd_oracle_gamma(xv) - math.gamma(xv)) < 2e-5 * math.gamma(xv)   # polynomial approximation
    assert L.ggd_oracle_gamma(-1.0) == 0.0


def test_cv_metrics_definitions(oracle):
    ls, M = [6, 5, 4], 8
    rng = np.random.RandomState(1)
    W, b = oracle.init_weights(ls, seed=1)
    x = rng.randn(19, 6).astype(np.float32); t = rng.randn(19, 4).astype(np.float32)
    net = oracle.OracleNet(ls, M, 0.1, 0.9, 0.0, 1.5, 1, W, b)
    net.train_bunch(x[:M], t[:M])
    out = net.forward(x).astype(np.float64)
    al = net.alpha().astype(np.float64)
    assert abs(net.cv_sqerr(x, t) - ((out - t) ** 2).sum()) < 1e-3
    assert abs(net.cv_abserr(x, t) - np.abs(out - t).sum() / 4) < 1e-3
    ll = 19 * 4 * math.log(1.5 / (2 * math.gamma(1 / 1.5))) - 19 * np.log(al).sum() - (np.abs((t - out) / al) ** 1.5).sum()
    assert abs(net.cv_loglik(x, t) - ll) < 2e-3 * abs(ll)


def test_raw_chunk_helper_matches_loader(oracle=None):
    """read_chunk_raw (inputs of the device-side loader) + the loader arithmetic in numpy == read_chunk, bit for bit"""
    from oracle import oracle as O
    fea, tg, nrm = (os.path.join(GOLDEN, f) for f in ("train_noisy.pfile", "train_clean.pfile", "train_noisy.norm"))
    la = O.PfileLoader(fea, tg, nrm, 257, 7, 3, 500, 27870775)
    lb = O.PfileLoader(fea, tg, nrm, 257, 7, 3, 500, 27870775)
    sent_en = len(la.sent_end) - 1
    starts, total = la.chunk_info(0, sent_en)
    for idx in range(len(starts)):
        ind, tgt = la.read_chunk(starts, total, sent_en, idx)
        frec, trec, first = lb.read_chunk_raw(starts, total, sent_en, idx)
        x = frec[:, 2:].copy().byteswap().view(np.float32)
        x = ((x - lb.mean).astype(np.float32) * lb.dvar).astype(np.float32)
        t = trec[:, 2:].copy().byteswap().view(np.float32)
        t = ((t - lb.mean).astype(np.float32) * lb.dvar).astype(np.float32)
        got_in = np.stack([x[f:f + 7].reshape(-1) for f in first])
        got_tg = np.stack([t[f + 3] for f in first])
        assert np.array_equal(got_in, ind) and np.array_equal(got_tg, tgt)


@pytest.mark.parametrize("traincache", [102400, 500])
def test_host_raw_chunk_reader(oracle, tmp_path, traincache):
    """host/interface.cpp: read_chunk_raw (what the drop-in executable feeds ggd_train_raw) against the loader restatement:
    same raw records, same shuffled row -> first-frame map, same random numbers consumed as the expanding reader"""
    L = _host_lib()
    ls = [1799, 32, 16, 257]
    W, b = oracle.init_weights(ls, seed=3)
    init = str(tmp_path / "init.wts")
    oracle.write_wts(init, ls, W, b)
    kw = _flags(tmp_path, init, traincache=traincache)
    h = L.bph_create(*_argv(kw))
    assert h
    nch, ns = C.c_int(), C.c_int()
    assert L.bph_chunk_info(h, b"0-7", 0, C.byref(nch), C.byref(ns)) == 0
    ld = oracle.PfileLoader(kw["fea_file"], kw["targ_file"], kw["norm_file"], 257, 7, 3, traincache, SEED)
    st, tot = ld.chunk_info(0, 7)
    assert (nch.value, ns.value) == (len(st), tot)
    order = list(range(len(st)))
    oracle.rand_index(order, ld.rng)
    idx = (C.c_int * len(st))(*range(len(st)))
    L.bph_shuffle(h, idx, len(st))
    assert list(idx) == order
    maxf = 4000
    fea = np.zeros((maxf, 259), np.uint32); tg = np.zeros((maxf, 259), np.uint32); first = np.zeros(traincache, np.int32)
    need = C.c_int()
    for ci in order:
        n = L.bph_read_chunk_raw(h, ci, fea.ctypes.data, tg.ctypes.data, first.ctypes.data, maxf, C.byref(need))
        fr, tr, f0 = ld.read_chunk_raw(st, tot, 7, ci)
        assert n == f0.size and need.value == fr.shape[0]
        assert np.array_equal(first[:n], f0)
        assert np.array_equal(fea[:need.value, 2:], fr[:, 2:]) and np.array_equal(tg[:need.value, 2:], tr[:, 2:])
    L.bph_destroy(h)


def test_frame_expand_matches_reference_loops():
    """oracle.frame_expand against a literal transcription of the loops of Test_code/frame_expand.m:6-25 (1-based indices)"""
    from oracle import oracle as O
    rng = np.random.RandomState(0)
    for T in (1, 2, 3, 9):
        f = rng.randn(T, 4)
        ctx = 7
        rows = []
        for t in range(1, T + 1):
            parts = []
            for c in range((ctx - 1) // 2, 0, -1):
                parts.append(f[0] if t - c <= 0 else f[t - c - 1])
            parts.append(f[t - 1])
            for c in range(1, (ctx - 1) // 2 + 1):
                parts.append(f[T - 1] if t + c >= T else f[t + c - 1])
            rows.append(np.concatenate(parts))
        assert np.array_equal(O.frame_expand(f, ctx), np.stack(rows))


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("traincache", [102400, 500])
def test_host_loader_record_slices(oracle, tmp_path, world, traincache):
    """data-parallel loader (host/interface.cpp read_chunk_raw_slice, ggd_raw_chunk::rec_frame0): every rank reads only its
    1/world of a chunk's pfile records; the slices tile the chunk exactly (what the library's all-gather reassembles) and
    every rank builds the SAME row -> first-frame map as the one-GPU loader (same lrand48 stream)"""
    L = _host_lib()
    ls = [1799, 32, 16, 257]
    W, b = oracle.init_weights(ls, seed=3)
    init = str(tmp_path / "init.wts")
    oracle.write_wts(init, ls, W, b)
    kw = _flags(tmp_path, init, traincache=traincache)
    hosts = [L.bph_create(*_argv(kw)) for _ in range(world + 1)]        # hosts[world] = the one-GPU loader
    assert all(hosts)
    nch, ns = C.c_int(), C.c_int()
    for h in hosts:
        assert L.bph_chunk_info(h, b"0-7", 0, C.byref(nch), C.byref(ns)) == 0
    maxf = 4096
    for ci in range(nch.value):
        full_f = np.zeros((maxf, 259), np.uint32); full_t = np.zeros((maxf, 259), np.uint32); full_first = np.zeros(traincache, np.int32)
        need = C.c_int()
        n = L.bph_read_chunk_raw(hosts[world], ci, full_f.ctypes.data, full_t.ctypes.data, full_first.ctypes.data, maxf, C.byref(need))
        assert n > 0
        S = -(-need.value // world)
        got_f = np.zeros_like(full_f); got_t = np.zeros_like(full_t)
        covered = 0
        for r in range(world):
            f = np.zeros((maxf, 259), np.uint32); t = np.zeros((maxf, 259), np.uint32); first = np.zeros(traincache, np.int32)
            nd, r0, nr = C.c_int(), C.c_int(), C.c_int()
            m = L.bph_read_chunk_raw_slice(hosts[r], ci, r, world, f.ctypes.data, t.ctypes.data, first.ctypes.data, maxf, C.byref(nd), C.byref(r0), C.byref(nr))
            assert m == n and nd.value == need.value
            assert r0.value == r * S and nr.value == max(0, min(S, need.value - r * S))
            assert np.array_equal(first[:n], full_first[:n])
            got_f[r0.value:r0.value + nr.value] = f[:nr.value]; got_t[r0.value:r0.value + nr.value] = t[:nr.value]
            covered += nr.value
        assert covered == need.value
        assert np.array_equal(got_f[:need.value], full_f[:need.value]) and np.array_equal(got_t[:need.value], full_t[:need.value])
    for h in hosts:
        L.bph_destroy(h)


def test_bench_reference_arm_uses_all_cores_under_torchrun_env():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the reference arm must still run the C oracle on every host core
    and report the thread count it really used"""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--workload", "mmse_1799x2048x3_257_b128"], env=env, capture_output=True, text=True, timeout=600)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    cores = len(os.sched_getaffinity(0))
    assert line["impl"] == "reference" and line["cpu_baseline"]["cores"] == cores and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0


def test_pfile_writer_reproduces_the_golden_pfile(oracle, tmp_path):
    """host/pfile_writer.cpp (the container around the LPS kernel's pfile records: feacat + pfile_concat,
    tools_pfile/pfile_noisy.pl:33,45): fed the records and sentence lengths of the reference's bundled train_noisy.pfile it
    must reproduce that file byte for byte (header text, records, sentence table); the .norm writer reproduces qnnorm's
    text layout (`vec N`, %g lines) and the bundled train_noisy.norm's numbers survive a round trip"""
    L = _host_lib()
    L.bph_write_pfile.argtypes = [C.c_char_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int]
    L.bph_write_norm.argtypes = [C.c_char_p, oracle_PF, oracle_PF, C.c_int]
    gold = os.path.join(GOLDEN, "train_noisy.pfile")
    blob = open(gold, "rb").read()
    feats, tail, sent = oracle.read_pfile(gold)
    nf = feats.shape[0]
    rec = np.frombuffer(blob[32768:32768 + nf * 259 * 4], np.uint32).copy()
    lens = np.diff(np.concatenate([[0], tail])).astype(np.int64)
    assert lens.sum() == nf and len(lens) == 10
    out = str(tmp_path / "mine.pfile")
    assert L.bph_write_pfile(out.encode(), rec.ctypes.data, nf, 257, lens.ctypes.data, len(lens)) == 0
    assert open(out, "rb").read() == blob
    # a frame count that does not match the sentence table is refused
    assert L.bph_write_pfile(out.encode(), rec.ctypes.data, nf - 1, 257, lens.ctypes.data, len(lens)) != 0
    mean, dvar = oracle.read_norm(os.path.join(GOLDEN, "train_noisy.norm"), 257)
    nout = str(tmp_path / "mine.norm")
    assert L.bph_write_norm(nout.encode(), mean.ctypes.data_as(oracle_PF), dvar.ctypes.data_as(oracle_PF), 257) == 0
    m2, d2 = oracle.read_norm(nout, 257)
    assert np.array_equal(m2, mean) and np.array_equal(d2, dvar)
    ref = str(tmp_path / "ref.norm")
    oracle.write_norm(ref, mean, dvar)
    assert open(nout).read() == open(ref).read()
    assert open(nout).read().split("\n")[:3] == open(os.path.join(GOLDEN, "train_noisy.norm")).read().split("\n")[:3]


def test_wav2pfile_fails_loudly_without_a_gpu(tmp_path):
    """no CPU fallback: the fused PCM -> pfile tool exits non-zero with the library's message when no CUDA device exists"""
    import subprocess
    exe = os.path.join(PKG_DIR, "host", "Wav2Pfile")
    if not os.path.exists(exe):
        pytest.skip("host/Wav2Pfile not built")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    raw = str(tmp_path / "a.raw")
    np.zeros(4000, np.int16).tofile(raw)
    p = subprocess.run([exe, "-o", str(tmp_path / "a.pfile"), raw], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "CUDA" in p.stderr


def test_enhance_tool_parses_files_and_fails_loudly_without_a_gpu(oracle, tmp_path):
    """host/Enhance_LPS (decode.m's network part): .norm and MAT-v4 .wts are parsed and validated before the GPU is touched;
    without a CUDA device it exits non-zero with the library's message (no CPU fallback)"""
    import subprocess
    exe = os.path.join(PKG_DIR, "host", "Enhance_LPS")
    if not os.path.exists(exe):
        pytest.skip("host/Enhance_LPS not built")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    ls = [7 * 257, 32, 257]
    W, b = oracle.init_weights(ls, seed=2)
    wts, nrm, lps = str(tmp_path / "m.wts"), str(tmp_path / "m.norm"), str(tmp_path / "a.lps")
    oracle.write_wts(wts, ls, W, b)
    mean, dvar = oracle.read_norm(os.path.join(GOLDEN, "train_noisy.norm"), 257)
    oracle.write_norm(nrm, mean, dvar)
    oracle.write_htk(lps, np.zeros((5, 257), np.float32))
    base = [exe, "-wts", wts, "-norm", nrm, lps, str(tmp_path / "o.lps")]
    p = subprocess.run(base + ["-layers", "1799,32,257"], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "CUDA" in p.stderr, p.stderr            # files accepted, then no device
    p = subprocess.run(base + ["-layers", "1799,33,257"], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "expected" in p.stderr                  # .wts shape disagrees with -layers
    p = subprocess.run(base + ["-layers", "1799,32,257", "-ctx", "5"], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "context" in p.stderr
