"""The fused PCM -> pfile (+ .norm) tool host/Wav2Pfile on the reference's two golden utterances (GPU only).
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
