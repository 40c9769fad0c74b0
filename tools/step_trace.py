"""GPU probe: device-side timeline of one training step of the named network (per launch: first CTA entry, last CTA exit,
median phase times).  Shows kernel durations and inter-kernel gaps without event/launch overhead."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_pkg
from oracle import oracle as O
pkg = load_pkg()
L = pkg.load_library()
PF = C.POINTER(C.c_float)
L.ggd_debug_trace_step.argtypes = [C.c_void_p, PF, PF, C.c_int, PF, C.c_int, C.POINTER(C.c_int)]
ls, M = [1799, 2048, 2048, 2048, 257], int(sys.argv[1]) if len(sys.argv) > 1 else 128
W, b = O.init_weights(ls, seed=1)
rng = np.random.RandomState(0)
x = rng.randn(M, ls[0]).astype(np.float32); t = rng.randn(M, ls[-1]).astype(np.float32)
for fused in (1, 0):
    net = pkg.BP_GPU(0, 0, 5, ls, M, 0.1, 0.9, 1e-5, W, b, 1.5, 1)
    out = np.zeros((32, 12), np.float32); n = C.c_int()
    rc = L.ggd_debug_trace_step(net.h, x.ctypes.data_as(PF), t.ctypes.data_as(PF), fused, out.ctypes.data_as(PF), 32, C.byref(n))
    assert rc == 0, L.ggd_last_error()
    print("fused=%d  kind ctas | first-entry last-exit (us) dur | medians since own entry: setup wdreq land0 landN acc staged epi exit" % fused)
    names = {0: "fwd", 2: "dx ", 3: "dw ", 9: "dwu"}
    for r in out[:n.value]:
        print("  %s %4d | %7.1f %7.1f %6.1f | %s" % (names[int(r[0])], int(r[1]), r[2], r[3], r[3] - r[2], " ".join("%5.1f" % v for v in r[4:])))
    net.close()
