"""LPS kernel against the reference's golden wav/lps pairs and the C oracle (GPU only)."""
import os
import numpy as np
import pytest
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
NAMES = ("TEST_DR8_MPAM0_SX289", "TEST_DR8_MPAM0_SX379")


def ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64); bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai); bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def check_close(got, ref):
    """LPS_FLAG_EXACT: the kernel runs the reference's own butterfly network"""
    # north star: 1e-4 relative, applied as |a-b| <= 1e-4*max(|b|,1) ...
    assert np.all(np.abs(got - ref) <= 1e-4 * np.maximum(np.abs(ref), 1.0))
    # ... and in fact (almost) bit-exact:
    d = ulp_diff(got, ref)
    assert d.max() <= 1, "max ulp distance %d" % d.max()
    assert (d > 0).mean() < 1e-3, "fraction of last-place differences %g" % (d > 0).mean()


def check_close_fast(got, ref):
    """default kernel (register-resident radix-8 FFT, fp32) against the reference's fp32 split-radix network.
    North star: 1e-4 relative, applied per bin as |a-b| <= 1e-4*max(|b|,1) for every bin within 40 dB of its
    frame's strongest bin.  Two DIFFERENT fp32 FFTs cannot agree better than their own round-off, ~1e-7 of the frame's peak
    AMPLITUDE, in any bin: for a bin r dB below (peak - 40 dB) the bound is widened by the amplitude ratio 10^(r/20), i.e. the
    absolute spectral error stays below ~1e-5 of the peak amplitude.  (Measured: the deviations sit in the DC / Nyquist bins of
    frames where those are 90-100 dB below the peak; a float64 FFT deviates from the golden files by MORE, 4e-3, there --
    it is the reference's own round-off.)  Plus: 1e-5 in the Frobenius norm and 99.99 % of all bins inside the plain bound."""
    assert got.shape == ref.shape
    if not ref.size:
        return
    assert np.linalg.norm(got.astype(np.float64) - ref) <= 1e-5 * max(np.linalg.norm(ref.astype(np.float64)), 1e-30)
    err = np.abs(got - ref)
    plain = 1e-4 * np.maximum(np.abs(ref), 1.0)
    below = np.maximum(ref.max(axis=1, keepdims=True) - ref - np.log(1e4), 0.0)     # nepers of POWER below (peak - 40 dB)
    assert np.all(err <= plain * np.exp(0.5 * below)), "worst %.3g of the bound" % (err / (plain * np.exp(0.5 * below))).max()
    assert (err <= plain).mean() >= 0.9999, (err > plain).mean()


@pytest.mark.parametrize("name", NAMES)
def test_golden_pairs(pkg, oracle, name):
    pcm = oracle.read_wav_pcm16(os.path.join(GOLDEN, name + ".wav"))
    hdr, gold = oracle.read_htk(os.path.join(GOLDEN, name + ".lps"))
    ex = pkg.Wav2LPS(0)
    got = ex.extract(pcm, flags=pkg.FLAG_EXACT)
    assert got.shape == gold.shape == (hdr["nSamples"], 257)
    check_close(got, gold)
    # big-endian output = bytes of the reference's HTK payload
    be = ex.extract(pcm, flags=pkg.FLAG_BIG_ENDIAN | pkg.FLAG_EXACT)
    assert np.array_equal(be.view(">f4").astype(np.float32), got)
    # default (fast) kernel on the same pair
    fast = ex.extract(pcm)
    check_close_fast(fast, gold)
    assert np.array_equal(ex.extract(pcm, flags=pkg.FLAG_BIG_ENDIAN).view(">f4").astype(np.float32), fast)


def test_random_noise_vs_oracle(pkg, oracle):
    rng = np.random.RandomState(1234)
    pcm = np.clip(np.round(rng.randn(16000 * 20) * 3000), -32768, 32767).astype(np.int16)
    ex = pkg.Wav2LPS(0)
    ref = oracle.lps_extract(pcm)
    check_close(ex.extract(pcm, flags=pkg.FLAG_EXACT), ref)
    check_close_fast(ex.extract(pcm), ref)
    check_close_fast(ex.extract(pcm[1:]), oracle.lps_extract(pcm[1:]))       # odd sample offset inside a batch: see ragged test


def test_edge_cases(pkg, oracle):
    ex = pkg.Wav2LPS(0)
    assert pkg.lps_nframes(511) == 0 and pkg.lps_nframes(512) == 1 and pkg.lps_nframes(767) == 1 and pkg.lps_nframes(768) == 2
    assert ex.extract(np.zeros(100, np.int16)).shape == (0, 257)
    for fl in (0, pkg.FLAG_EXACT):
        z = ex.extract(np.zeros(1024, np.int16), flags=fl)   # silence -> floored at -50 (Wav2LogSpec_be.c:476-477)
        assert z.shape == (3, 257) and np.all(z == -50.0)
    full = np.full(600, -32768, np.int16)               # extreme amplitude, trailing partial hop dropped
    check_close(ex.extract(full, flags=pkg.FLAG_EXACT), oracle.lps_extract(full))
    # a constant signal is a pure DC line: every other bin is window leakage 150+ dB down, where fp32 FFTs differ by design
    got, ref = ex.extract(full), oracle.lps_extract(full)
    assert np.all(np.abs(got[:, :3] - ref[:, :3]) <= 1e-4 * np.maximum(np.abs(ref[:, :3]), 1.0))


def test_batch_ragged_and_zscore(pkg, oracle):
    rng = np.random.RandomState(5)
    lens = [0, 300, 512, 5000, 777, 16000, 256 * 9 + 1]
    pcm = np.clip(np.round(rng.randn(sum(lens)) * 2000), -32768, 32767).astype(np.int16)
    off = np.concatenate([[0], np.cumsum(lens)])
    ex = pkg.Wav2LPS(0)
    ref = np.concatenate([oracle.lps_extract(pcm[off[i]:off[i + 1]]) for i in range(len(lens))])
    got = ex.extract_batch(pcm, off, flags=pkg.FLAG_EXACT)
    assert got.shape == ref.shape
    check_close(got, ref)
    check_close_fast(ex.extract_batch(pcm, off), ref)    # utterances starting at odd sample offsets take the unaligned load path
    mean, dvar = oracle.read_norm(os.path.join(GOLDEN, "train_noisy.norm"), 257)
    ex.set_norm(mean, dvar)
    zs = ex.extract_batch(pcm, off, flags=pkg.FLAG_ZSCORE)
    want = ((ref - mean) * dvar).astype(np.float32)      # Interface.cc:763-764
    assert np.all(np.abs(zs - want) <= 1e-4 * np.maximum(np.abs(want), 1.0))


def test_pipelined_host_batch(pkg, oracle):
    """more frames than one pipeline piece (32 768): uploads, kernels and downloads of the pieces overlap on two streams"""
    rng = np.random.RandomState(77)
    lens = [256 * 20000 + 300, 0, 256 * 30000 + 17, 700, 256 * 25000]
    pcm = np.clip(np.round(rng.randn(sum(lens)) * 2500), -32768, 32767).astype(np.int16)
    off = np.concatenate([[0], np.cumsum(lens)])
    ex = pkg.Wav2LPS(0)
    got = ex.extract_batch(pcm, off)
    assert got.shape[0] == sum(pkg.lps_nframes(n) for n in lens) > 2 * 32768
    # spot-check utterance boundaries and piece boundaries against the oracle
    fo = np.concatenate([[0], np.cumsum([pkg.lps_nframes(n) for n in lens])])
    for u in (0, 2, 3, 4):
        n = min(lens[u], 256 * 40 + 512)
        ref = oracle.lps_extract(pcm[off[u]:off[u] + n])
        check_close_fast(got[fo[u]:fo[u] + ref.shape[0]], ref)
    for f in (32768, 65536):
        u = int(np.searchsorted(fo, f, side="right") - 1)
        s0 = off[u] + (f - fo[u] - 2) * 256
        ref = oracle.lps_extract(pcm[s0:s0 + 256 * 5 + 256])
        check_close_fast(got[f - 2:f - 2 + ref.shape[0]], ref)
    dev = ex.extract_batch(pcm, off, flags=pkg.FLAG_EXACT)
    check_close_fast(got, dev)


def test_pfile_records_and_norm_production(pkg, oracle, tmp_path):
    """the LPS kernel writes QuickNet pfile RECORDS (feacat, tools_pfile/pfile_noisy.pl:33) and accumulates the
    per-bin statistics of the .norm file (qnnorm, get_norm.pl:4) on the device.  Checks: the records wrapped in the pfile
    container are read back by the reference-format reader with the right sentence / frame indices and features; the norm
    equals the float64 numpy statistics (ddof 0); and qnnorm's own output is reproduced from the reference's golden pfile."""
    rng = np.random.RandomState(9)
    lens = [4000, 513, 9000, 256 * 7]
    pcm = np.clip(np.round(rng.randn(sum(lens)) * 2000), -32768, 32767).astype(np.int16)
    off = np.concatenate([[0], np.cumsum(lens)])
    ex = pkg.Wav2LPS(0)
    feats = ex.extract_batch(pcm, off, flags=pkg.FLAG_EXACT)
    ex.norm_reset()
    rec = ex.extract_batch(pcm, off, flags=pkg.FLAG_EXACT | pkg.FLAG_PFILE | pkg.FLAG_ACCUM_NORM)
    words = rec.view(">u4")
    nfr = [pkg.lps_nframes(n) for n in lens]
    assert rec.shape == (sum(nfr), 259)
    assert np.array_equal(words[:, 0], np.repeat(np.arange(len(lens)), nfr))
    assert np.array_equal(words[:, 1], np.concatenate([np.arange(n) for n in nfr]))
    assert np.array_equal(rec[:, 2:].view(">f4").astype(np.float32), feats)
    # the container around the records: header + records + sentence index tail, read back by the loader's format reader
    ref_path, my_path = str(tmp_path / "ref.pfile"), str(tmp_path / "mine.pfile")
    oracle.write_pfile(ref_path, feats, nfr)
    blob = open(ref_path, "rb").read()
    body = len(feats) * 259 * 4
    open(my_path, "wb").write(blob[:32768] + rec.tobytes() + blob[32768 + body:])
    assert open(my_path, "rb").read() == blob
    mean, dvar, n = ex.norm_finalize()
    assert n == len(feats)
    x = feats.astype(np.float64)
    assert np.allclose(mean, x.mean(0), rtol=1e-6) and np.allclose(dvar, 1.0 / x.std(0), rtol=1e-5)
    # qnnorm pinned by the reference's own files: statistics of the golden pfile against the golden .norm (6 printed digits)
    gfeat = oracle.read_pfile(os.path.join(GOLDEN, "train_noisy.pfile"))[0]
    gmean, gdvar = oracle.read_norm(os.path.join(GOLDEN, "train_noisy.norm"), 257)
    import torch
    d = torch.from_numpy(np.ascontiguousarray(gfeat)).cuda()
    m2, v2 = ex.norm_of_device_features(d.data_ptr(), gfeat.shape[0])
    assert np.allclose(m2, gmean, rtol=1e-5) and np.allclose(v2, gdvar, rtol=1e-5)
    # fast kernel: same records within the fast tolerance
    ex.norm_reset()
    rec2 = ex.extract_batch(pcm, off, flags=pkg.FLAG_PFILE | pkg.FLAG_ACCUM_NORM)
    check_close_fast(rec2[:, 2:].view(">f4").astype(np.float32), feats)
    m3, v3, _ = ex.norm_finalize()
    assert np.allclose(m3, mean, rtol=1e-5) and np.allclose(v3, dvar, rtol=1e-4)
