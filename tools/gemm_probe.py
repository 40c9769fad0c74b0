"""GPU probe: warm timing and per-CTA phase timeline of the tcgen05 GEMM for the shapes of the named network."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_pkg
pkg = load_pkg()
L = pkg.load_library()
PF = C.POINTER(C.c_float)
L.ggd_debug_gemm_timed.argtypes = [C.c_int] * 7 + [PF, PF, PF, C.c_int, PF, C.POINTER(C.c_ulonglong), C.c_int]
SLOTS = ["entry", "setup", "issue0", "land0", "landN", "mmaN", "acc", "xchg", "epi", "exit"]


def run(name, a_mn, b_mn, I, J, R, bn, splits, reps=50):
    rng = np.random.RandomState(0)
    A = rng.randn(*((R, I) if a_mn else (I, R))).astype(np.float32)
    B = rng.randn(*((R, J) if b_mn else (J, R))).astype(np.float32)
    D = np.zeros((I, J), np.float32)
    ctas = splits * (-(-J // bn)) * (-(-I // 128))
    tr = np.zeros((ctas, 16), np.uint64)
    ms = C.c_float()
    rc = L.ggd_debug_gemm_timed(a_mn, b_mn, I, J, R, bn, splits, A.ctypes.data_as(PF), B.ctypes.data_as(PF), D.ctypes.data_as(PF),
                                reps, C.byref(ms), tr.ctypes.data_as(C.POINTER(C.c_ulonglong)), ctas)
    if rc != 0:
        print(name, "ERROR", L.ggd_last_error().decode()); return
    ref = (A.T if a_mn else A).astype(np.float64) @ (B.T if b_mn else B).astype(np.float64).T
    err = np.linalg.norm(D - ref) / np.linalg.norm(ref)
    t = tr[:, :10].astype(np.float64)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1000.0
    line = " ".join("%s=%.1f/%.1f" % (SLOTS[k], np.median(rel[:, k]), rel[:, k].max()) for k in range(10))
    flops = 2.0 * I * J * R
    print("%-28s bn=%3d S=%d ctas=%3d  %.2f us (%.1f TF alg)  err=%.1e | median/max us since first entry: %s" %
          (name, bn, splits, ctas, ms.value * 1e3, flops / (ms.value * 1e-3) / 1e12, err, line), flush=True)


if __name__ == "__main__":
    for bn, ss in ((128, (8, 4, 2, 1)), (64, (4, 2, 1))):
        for s in ss:
            run("fwd 128x2048x2048 (K,MN)", 0, 1, 128, 2048, 2048, bn, s)
    for bn, ss in ((128, (8, 4)), (64, (4,))):
        for s in ss:
            run("dX 128x2048x2048 (K,K)", 0, 0, 128, 2048, 2048, bn, s)
    for bn in (128, 64):
        run("dW 2048x2048x128 (MN,MN)", 1, 1, 2048, 2048, 128, bn, 1)
    run("fwd top 128x257x2048", 0, 1, 128, 257, 2048, 64, 4)
    run("fwd L1 128x2048x1799", 0, 1, 128, 2048, 1799, 128, 8)
    run("fwd M=1024", 0, 1, 1024, 2048, 2048, 128, 1)
    run("fwd M=1024 S2", 0, 1, 1024, 2048, 2048, 128, 2)
