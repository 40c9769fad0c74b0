"""Parity against the UNMODIFIED reference CUDA code (oracle/_ref/, built from /root/reference by
oracle/build_ref.sh) running on the same B200: the strongest pin of both the C oracle and the CUDA path.
Skipped where oracle/_ref was not built."""
import os
import re
import subprocess
import numpy as np
import pytest
from conftest import rel_err, GOLDEN, PKG_DIR, ROOT

pytestmark = pytest.mark.gpu


def _case(O, ls, n, seed):
    rng = np.random.RandomState(seed)
    W, b = O.init_weights(ls, seed=seed)
    b = [rng.uniform(-0.1, 0.1, x.size).astype(np.float32) for x in b]
    return W, b, rng.randn(n, ls[0]).astype(np.float32), rng.randn(n, ls[-1]).astype(np.float32)


@pytest.mark.parametrize("MLflag,beta", [(1, 1.5), (1, 1.0), (0, 2.0)])
def test_chunk_vs_reference_cuda(pkg, oracle, MLflag, beta):
    from oracle import refcuda
    if not refcuda.available("libref_bpgpu.so"):
        pytest.skip("oracle/_ref/libref_bpgpu.so not built")
    O = oracle
    ls, M, nb = [7 * 33, 160, 96, 33], 128, 12
    W, b, x, t = _case(O, ls, nb * M + 11, 21)
    ref = refcuda.RefBPGPU(ls, M, 0.05, 0.9, 1e-5, beta, MLflag, W, b)
    ref.train(x, t)
    Wr, br = ref.weights()
    xc, tc = x[:333], t[:333]
    cv_ref = [ref.cv(k, xc, tc) for k in range(3 if MLflag == 1 else 2)]
    ref.close()
    orc = O.OracleNet(ls, M, 0.05, 0.9, 1e-5, beta, MLflag, W, b)
    orc.train(x, t)
    Wo, bo = orc.weights()
    cv_orc = [orc.cv_sqerr(xc, tc), orc.cv_abserr(xc, tc)] + ([orc.cv_loglik(xc, tc)] if MLflag == 1 else [])
    for precision, tol in ((1, 2e-4), (0, 1e-3)):
        net = pkg.BP_GPU(0, 0, len(ls), ls, M, 0.05, 0.9, 1e-5, W, b, beta, MLflag, precision=precision)
        net.train(x.shape[0], x, t)
        Wg, bg = net.returnWeights()
        cv_got = [net.CrossValid(333, xc, tc), net.CrossValiddB(333, xc, tc)] + ([net.CrossValid2(333, xc, tc)] if MLflag == 1 else [])
        for l in range(len(ls) - 1):
            assert rel_err(Wg[l], Wr[l]) < tol, ("W", l, precision)
            assert rel_err(bg[l], br[l]) < 5 * tol, ("b", l, precision)
        for g, r in zip(cv_got, cv_ref):
            assert abs(g - r) <= 2e-3 * abs(r), (cv_got, cv_ref)
        net.close()
    # the oracle itself against the reference (this is what pins the oracle)
    for l in range(len(ls) - 1):
        assert rel_err(Wo[l], Wr[l]) < 2e-4
        assert rel_err(bo[l], br[l]) < 1e-3
    for o, r in zip(cv_orc, cv_ref):
        assert abs(o - r) <= 1e-3 * abs(r)


def _log_values(path):
    txt = open(path).read()
    vals = {}
    for key, pat in (("sq", r"CV over\. squared error: ([-\d.eE+naninf]+)"), ("abs", r"square root squared error: ([-\d.eE+naninf]+)"),
                     ("ll", r"CV log likelihood: ([-\d.eE+naninf]+)"), ("samples", r"Training sentences have (\d+) chunks, (\d+) samples"),
                     ("cv_samples", r"CV sentences have (\d+) chunks, (\d+) samples")):
        m = re.search(pat, txt)
        if m:
            vals[key] = float(m.group(len(m.groups())))
    return vals


@pytest.mark.parametrize("MLflag,beta", [(1, 1.5), (0, 2.0)])
def test_cli_epoch_on_bundled_pfile(pkg, oracle, tmp_path, MLflag, beta):
    """One finetune.pl epoch (same flags) on the bundled pfiles: this repository's BPtrain_Sigmoid vs the
    reference's own binary; compares the written .wts and the CV lines of the log."""
    from oracle import refcuda
    O = oracle
    mine = os.path.join(PKG_DIR, "host", "BPtrain_Sigmoid")
    if not os.path.exists(mine):
        pytest.skip("host/BPtrain_Sigmoid not built")
    ls = [1799, 2048, 2048, 2048, 257]
    W, b = O.init_weights(ls, seed=4)
    init = str(tmp_path / "init.wts")
    O.write_wts(init, ls, W, b)

    def flags(tag):
        return ["gpu_used=0", "numlayers=5", "layersizes=1799,2048,2048,2048,257", "bunchsize=128", "MLflag=%d" % MLflag,
                "shapefactor=%g" % beta, "momentum=0.9", "weightcost=0.00001", "lrate=0.1", "fea_dim=257", "fea_context=7",
                "traincache=102400", "init_randem_seed=27870775", "targ_offset=3", "initwts_file=" + init,
                "norm_file=" + os.path.join(GOLDEN, "train_noisy.norm"), "fea_file=" + os.path.join(GOLDEN, "train_noisy.pfile"),
                "targ_file=" + os.path.join(GOLDEN, "train_clean.pfile"), "outwts_file=" + str(tmp_path / (tag + ".wts")),
                "log_file=" + str(tmp_path / (tag + ".log")), "train_sent_range=0-7", "cv_sent_range=8-9", "dropoutflag=0",
                "visible_omit=0.1", "hid_omit=0.1"]
    subprocess.run([mine] + flags("mine"), check=True, cwd=str(tmp_path), stdout=subprocess.DEVNULL)
    vm = _log_values(str(tmp_path / "mine.log"))
    assert vm["samples"] == 1443 and vm["cv_samples"] == 382
    Wm, bm = O.read_wts(str(tmp_path / "mine.wts"), ls)
    # oracle epoch through the numpy loader restatement
    ld = O.PfileLoader(os.path.join(GOLDEN, "train_noisy.pfile"), os.path.join(GOLDEN, "train_clean.pfile"),
                       os.path.join(GOLDEN, "train_noisy.norm"), 257, 7, 3, 102400, 27870775)
    st, tot = ld.chunk_info(0, 7)
    x, t = ld.read_chunk(st, tot, 7, 0)
    orc = O.OracleNet(ls, 128, 0.1, 0.9, 1e-5, beta, MLflag, W, b)
    orc.train(x, t)
    Wo, bo = orc.weights()
    cst, ctot = ld.chunk_info(8, 9)
    xc, tc = ld.read_chunk(cst, ctot, 9, 0, shuffle=False)
    for l in range(4):
        assert rel_err(Wm[l], Wo[l]) < 1e-3, l
    assert abs(vm["sq"] - orc.cv_sqerr(xc, tc) / ctot) <= 2e-3 * abs(vm["sq"])
    assert abs(vm["abs"] - orc.cv_abserr(xc, tc) / ctot) <= 2e-3 * abs(vm["abs"])
    if MLflag == 1:
        assert abs(vm["ll"] - orc.cv_loglik(xc, tc) / ctot) <= 2e-3 * abs(vm["ll"])
    # The reference's own pipeline on the same flags.  Its full binary cannot be used: BPtrain.cc's threadFetch
    # falls off the end of a non-void function, which g++ >= 8 compiles to a trap / fall-through (it hangs at -O2
    # and dies with SIGILL at -O0 on this toolchain).  So the reference's two halves, each compiled UNMODIFIED,
    # are driven in BPtrain.cc's order instead: Interface (loader, Interface.cc) + BP_GPU (device, BP_GPU.cu).
    if refcuda.available("libref_bpgpu.so") and refcuda.available("libref_interface.so"):
        kw = dict(f.split("=", 1) for f in flags("ref"))
        rif = refcuda.RefInterface(**kw)
        nch, ns = rif.train_info("0-7")
        assert (nch, ns) == (1, 1443)
        order = rif.shuffle_chunks(nch)
        rbp = refcuda.RefBPGPU(ls, 128, 0.1, 0.9, 1e-5, beta, MLflag, W, b)
        for ci in order:
            xr, tr = rif.read_chunk(ci, 1799, 257)
            rbp.train(xr, tr)
        Wr, br = rbp.weights()
        cch, cns = rif.cv_info("8-9")
        xcr, tcr = rif.read_chunk(0, 1799, 257, cv=True)
        vr = {"sq": rbp.cv(0, xcr, tcr) / cns, "abs": rbp.cv(1, xcr, tcr) / cns}
        if MLflag == 1:
            vr["ll"] = rbp.cv(2, xcr, tcr) / cns
        rbp.close()
        for l in range(4):
            assert rel_err(Wm[l], Wr[l]) < 1e-3, ("this repo vs reference CUDA", l)
            assert rel_err(Wo[l], Wr[l]) < 1e-3, ("oracle vs reference CUDA", l)
        for k in vr:
            assert abs(vm[k] - vr[k]) <= 2e-3 * abs(vr[k]), (k, vm, vr)
        print("reference CUDA:", vr, "ours:", vm)


def test_wav2lps_cli(pkg, oracle, tmp_path):
    """Wav2LPS_be drop-in writes the reference's HTK file byte-for-byte (up to last-place log differences)."""
    exe = os.path.join(PKG_DIR, "host", "Wav2LPS_be")
    if not os.path.exists(exe):
        pytest.skip("host/Wav2LPS_be not built")
    name = "TEST_DR8_MPAM0_SX289"
    pcm = oracle.read_wav_pcm16(os.path.join(GOLDEN, name + ".wav"))
    raw, out = str(tmp_path / "x.raw"), str(tmp_path / "x.lps")
    pcm.tofile(raw)
    subprocess.run([exe, "-F", "RAW", "-fs", "16", raw, out], check=True, stderr=subprocess.DEVNULL)
    a, b = open(out, "rb").read(), open(os.path.join(GOLDEN, name + ".lps"), "rb").read()
    assert len(a) == len(b) and a[:12] == b[:12]
    ha, fa = oracle.read_htk(out)
    hb, fb = oracle.read_htk(os.path.join(GOLDEN, name + ".lps"))
    assert ha == hb
    assert np.mean(fa.view(np.uint32) != fb.view(np.uint32)) < 1e-3
    assert np.all(np.abs(fa - fb) <= 1e-4 * np.maximum(np.abs(fb), 1.0))


@pytest.mark.parametrize("ngpu", [2, 4, 8])
def test_cli_multi_gpu_equals_single_gpu(pkg, oracle, tmp_path, ngpu):
    """the drop-in executable with gpu_used=0,...,N-1 (forked workers, frame-sharded: bunchsize stays the GLOBAL minibatch,
    every rank reads only its slice of the chunk's pfile records and the library all-gathers them over NVLink) writes the
    same .wts and CV lines as the one-GPU run with the same finetune.pl flags (tolerance 1e-3)"""
    import torch
    if torch.cuda.device_count() < ngpu:
        pytest.skip("needs %d GPUs" % ngpu)
    O = oracle
    mine = os.path.join(PKG_DIR, "host", "BPtrain_Sigmoid")
    if not os.path.exists(mine):
        pytest.skip("host/BPtrain_Sigmoid not built")
    ls = [1799, 2048, 2048, 2048, 257]
    W, b = O.init_weights(ls, seed=4)
    init = str(tmp_path / "init.wts")
    O.write_wts(init, ls, W, b)

    def flags(tag, gpus):
        return ["gpu_used=" + gpus, "numlayers=5", "layersizes=1799,2048,2048,2048,257", "bunchsize=128", "MLflag=1",
                "shapefactor=1.5", "momentum=0.9", "weightcost=0.00001", "lrate=0.1", "fea_dim=257", "fea_context=7",
                "traincache=102400", "init_randem_seed=27870775", "targ_offset=3", "initwts_file=" + init,
                "norm_file=" + os.path.join(GOLDEN, "train_noisy.norm"), "fea_file=" + os.path.join(GOLDEN, "train_noisy.pfile"),
                "targ_file=" + os.path.join(GOLDEN, "train_clean.pfile"), "outwts_file=" + str(tmp_path / (tag + ".wts")),
                "log_file=" + str(tmp_path / (tag + ".log")), "train_sent_range=0-7", "cv_sent_range=8-9", "dropoutflag=0",
                "visible_omit=0.1", "hid_omit=0.1"]
    subprocess.run([mine] + flags("one", "0"), check=True, cwd=str(tmp_path), stdout=subprocess.DEVNULL, timeout=300)
    subprocess.run([mine] + flags("many", ",".join(str(i) for i in range(ngpu))), check=True, cwd=str(tmp_path), stdout=subprocess.DEVNULL, timeout=300)
    v1, vn = _log_values(str(tmp_path / "one.log")), _log_values(str(tmp_path / "many.log"))
    assert vn["samples"] == v1["samples"] == 1443 and vn["cv_samples"] == v1["cv_samples"] == 382
    W1, b1 = O.read_wts(str(tmp_path / "one.wts"), ls)
    Wn, bn = O.read_wts(str(tmp_path / "many.wts"), ls)
    for l in range(4):
        assert rel_err(Wn[l], W1[l]) < 1e-3, l
        assert rel_err(Wn[l] - W[l], W1[l] - W[l]) < 2e-3, "update of layer %d" % (l + 1)
        assert rel_err(bn[l], b1[l]) < 2e-3, l
    for k in ("sq", "abs", "ll"):
        assert abs(vn[k] - v1[k]) <= 2e-3 * abs(v1[k]), k
