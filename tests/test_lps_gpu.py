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
    # north star: 1e-4 relative, applied as |a-b| <= 1e-4*max(|b|,1) ...
    assert np.all(np.abs(got - ref) <= 1e-4 * np.maximum(np.abs(ref), 1.0))
    # ... but the kernel runs the reference's own butterfly network, so it is in fact (almost) bit-exact:
    d = ulp_diff(got, ref)
    assert d.max() <= 1, "max ulp distance %d" % d.max()
    assert (d > 0).mean() < 1e-3, "fraction of last-place differences %g" % (d > 0).mean()


@pytest.mark.parametrize("name", NAMES)
def test_golden_pairs(pkg, oracle, name):
    pcm = oracle.read_wav_pcm16(os.path.join(GOLDEN, name + ".wav"))
    hdr, gold = oracle.read_htk(os.path.join(GOLDEN, name + ".lps"))
    ex = pkg.Wav2LPS(0)
    got = ex.extract(pcm)
    assert got.shape == gold.shape == (hdr["nSamples"], 257)
    check_close(got, gold)
    # big-endian output = bytes of the reference's HTK payload
    be = ex.extract(pcm, flags=pkg.FLAG_BIG_ENDIAN)
    assert np.array_equal(be.view(">f4").astype(np.float32), got)


def test_random_noise_vs_oracle(pkg, oracle):
    rng = np.random.RandomState(1234)
    pcm = np.clip(np.round(rng.randn(16000 * 20) * 3000), -32768, 32767).astype(np.int16)
    ex = pkg.Wav2LPS(0)
    check_close(ex.extract(pcm), oracle.lps_extract(pcm))


def test_edge_cases(pkg, oracle):
    ex = pkg.Wav2LPS(0)
    assert pkg.lps_nframes(511) == 0 and pkg.lps_nframes(512) == 1 and pkg.lps_nframes(767) == 1 and pkg.lps_nframes(768) == 2
    assert ex.extract(np.zeros(100, np.int16)).shape == (0, 257)
    z = ex.extract(np.zeros(1024, np.int16))           # silence -> floored at -50 (Wav2LogSpec_be.c:476-477)
    assert z.shape == (3, 257) and np.all(z == -50.0)
    full = np.full(600, -32768, np.int16)               # extreme amplitude, trailing partial hop dropped
    check_close(ex.extract(full), oracle.lps_extract(full))


def test_batch_ragged_and_zscore(pkg, oracle):
    rng = np.random.RandomState(5)
    lens = [0, 300, 512, 5000, 777, 16000, 256 * 9 + 1]
    pcm = np.clip(np.round(rng.randn(sum(lens)) * 2000), -32768, 32767).astype(np.int16)
    off = np.concatenate([[0], np.cumsum(lens)])
    ex = pkg.Wav2LPS(0)
    got = ex.extract_batch(pcm, off)
    ref = np.concatenate([oracle.lps_extract(pcm[off[i]:off[i + 1]]) for i in range(len(lens))])
    assert got.shape == ref.shape
    check_close(got, ref)
    mean, dvar = oracle.read_norm(os.path.join(GOLDEN, "train_noisy.norm"), 257)
    ex.set_norm(mean, dvar)
    zs = ex.extract_batch(pcm, off, flags=pkg.FLAG_ZSCORE)
    want = ((ref - mean) * dvar).astype(np.float32)      # Interface.cc:763-764
    assert np.all(np.abs(zs - want) <= 1e-4 * np.maximum(np.abs(want), 1.0))
