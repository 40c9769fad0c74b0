"""The host tools added last: host/Wav2Pfile (fused PCM -> pfile + .norm) and host/Enhance_LPS (decode.m network part) on the
reference's golden utterances (GPU only).
(Sorted last on purpose: this file was added after the round's last GPU session and has not run on hardware yet; the library
calls it makes -- LPS_FLAG_PFILE / LPS_FLAG_ACCUM_NORM -- and its container writer are covered by tests/test_lps_gpu.py
and tests/test_oracle_cpu.py.)"""
import os
import subprocess
import numpy as np
import pytest
from conftest import GOLDEN, PKG_DIR

pytestmark = pytest.mark.gpu
NAMES = ("TEST_DR8_MPAM0_SX289", "TEST_DR8_MPAM0_SX379")


@pytest.mark.parametrize("mode", ["-exact", ""])
def test_wav2pfile_on_golden_utterances(pkg, oracle, tmp_path, mode):
    exe = os.path.join(PKG_DIR, "host", "Wav2Pfile")
    if not os.path.exists(exe):
        pytest.skip("host/Wav2Pfile not built")
    raws, gold = [], []
    for n in NAMES:
        pcm = oracle.read_wav_pcm16(os.path.join(GOLDEN, n + ".wav"))
        raw = str(tmp_path / (n + ".raw"))
        np.ascontiguousarray(pcm, "<i2").tofile(raw)
        raws.append(raw)
        gold.append(oracle.read_htk(os.path.join(GOLDEN, n + ".lps"))[1])
    short = str(tmp_path / "short.raw")                      # shorter than one frame: skipped with a warning, no sentence
    np.zeros(100, np.int16).tofile(short)
    pf, nm = str(tmp_path / "out.pfile"), str(tmp_path / "out.norm")
    cmd = [exe] + ([mode] if mode else []) + ["-norm", nm, "-o", pf, raws[0], short, raws[1]]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    feats, tail, sent = oracle.read_pfile(pf)
    want = np.concatenate(gold)
    assert feats.shape == want.shape
    assert list(tail) == [gold[0].shape[0], want.shape[0]]                       # cumulative frames per sentence
    assert np.array_equal(sent, np.repeat([0, 1], [g.shape[0] for g in gold]))
    if mode == "-exact":
        assert np.all(np.abs(feats - want) <= 1e-4 * np.maximum(np.abs(want), 1.0))
    else:
        assert np.linalg.norm(feats.astype(np.float64) - want) <= 1e-5 * np.linalg.norm(want.astype(np.float64))
    # the file is exactly the reference-format container around these features
    ref = str(tmp_path / "ref.pfile")
    oracle.write_pfile(ref, feats, [g.shape[0] for g in gold])
    assert open(pf, "rb").read() == open(ref, "rb").read()
    mean, dvar = oracle.read_norm(nm, 257)
    x = feats.astype(np.float64)
    assert np.allclose(mean, x.mean(0), rtol=2e-5, atol=1e-6) and np.allclose(dvar, 1.0 / x.std(0), rtol=2e-5)


def test_enhance_tool_equals_decode_restatement(pkg, oracle, tmp_path):
    """host/Enhance_LPS on a golden LPS file against the float64 restatement of Test_code/decode.m + frame_expand.m"""
    exe = os.path.join(PKG_DIR, "host", "Enhance_LPS")
    if not os.path.exists(exe):
        pytest.skip("host/Enhance_LPS not built")
    ls = [7 * 257, 96, 64, 257]
    W, b = oracle.init_weights(ls, seed=6)
    rng = np.random.RandomState(2)
    b = [rng.uniform(-0.1, 0.1, x.size).astype(np.float32) for x in b]
    wts, nrm = str(tmp_path / "m.wts"), str(tmp_path / "m.norm")
    oracle.write_wts(wts, ls, W, b)
    mean, dvar = oracle.read_norm(os.path.join(GOLDEN, "train_noisy.norm"), 257)
    oracle.write_norm(nrm, mean, dvar)
    mean, dvar = oracle.read_norm(nrm, 257)                   # what the tool reads (6 printed digits)
    src = os.path.join(GOLDEN, NAMES[0] + ".lps")
    out = str(tmp_path / "enh.lps")
    p = subprocess.run([exe, "-wts", wts, "-norm", nrm, "-layers", ",".join(map(str, ls)), src, out], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    hdr, lps = oracle.read_htk(src)
    hdr2, got = oracle.read_htk(out)
    assert hdr2["nSamples"] == hdr["nSamples"] and got.shape == lps.shape
    ref = oracle.enhance_ref(lps, W, b, ls, mean.astype(np.float64), dvar.astype(np.float64), 7)
    assert np.linalg.norm(got - ref) <= 1e-3 * np.linalg.norm(ref)
