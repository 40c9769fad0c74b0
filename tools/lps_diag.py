"""where does the fast LPS kernel deviate from the exact one? (tuning / debugging aid)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
from oracle import oracle as O
pkg = load_pkg()
ex = pkg.Wav2LPS(0)


def report(name, fast, exact):
    err = np.abs(fast - exact) / np.maximum(np.abs(exact), 1.0)
    print("==", name, "frames", fast.shape, "max err %.3e" % err.max(), "frac>1e-4 %.2e" % (err > 1e-4).mean(), "frac>1e-5 %.2e" % (err > 1e-5).mean())
    idx = np.argsort(err.ravel())[::-1][:8]
    for i in idx:
        f, k = divmod(int(i), 257)
        print("   frame %6d bin %3d fast %.6f exact %.6f err %.3e   frame max %.3f" % (f, k, fast[f, k], exact[f, k], err[f, k], exact[f].max()))


rng = np.random.RandomState(1234)
pcm = np.clip(np.round(rng.randn(16000 * 20) * 3000), -32768, 32767).astype(np.int16)
report("noise", ex.extract(pcm), ex.extract(pcm, flags=pkg.FLAG_EXACT))
report("noise[1:]", ex.extract(pcm[1:]), ex.extract(pcm[1:], flags=pkg.FLAG_EXACT))
report("noise[1:] copy", ex.extract(pcm[1:].copy()), ex.extract(pcm[1:].copy(), flags=pkg.FLAG_EXACT))
for n in ("TEST_DR8_MPAM0_SX289", "TEST_DR8_MPAM0_SX379"):
    p = O.read_wav_pcm16(os.path.join(ROOT, "tests", "golden", n + ".wav"))
    report(n, ex.extract(p), ex.extract(p, flags=pkg.FLAG_EXACT))
rng = np.random.RandomState(77)
lens = [256 * 20000 + 300, 0, 256 * 30000 + 17, 700, 256 * 25000]
pcm = np.clip(np.round(rng.randn(sum(lens)) * 2500), -32768, 32767).astype(np.int16)
off = np.concatenate([[0], np.cumsum(lens)])
report("pipelined", ex.extract_batch(pcm, off), ex.extract_batch(pcm, off, flags=pkg.FLAG_EXACT))
