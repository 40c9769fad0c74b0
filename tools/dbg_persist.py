"""Stage-by-stage exercise of the production training path at the named shape (each stage prints before the next starts);
run under `timeout` on the GPU box when a pipeline change needs bisecting."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_pkg
pkg = load_pkg()
ls, bunch = [1799, 2048, 2048, 2048, 257], 128
rng = np.random.RandomState(1)
W = [rng.uniform(-0.05, 0.05, ls[i] * ls[i + 1]).astype(np.float32) for i in range(4)]
b = [np.zeros(ls[i + 1], np.float32) for i in range(4)]
net = pkg.BP_GPU(0, 0, 5, ls, bunch, 0.1, 0.9, 1e-5, W, b, 1.5, 1)
dev = torch.device("cuda", 0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
d_in = torch.randn(nb * bunch, ls[0], device=dev)
d_tg = torch.randn(nb * bunch, ls[-1], device=dev)
net.reserve(nb * bunch)
for n in (1, 3, 15, 16, 17, 32, nb):
    t0 = time.time()
    net.train_device(n * bunch, d_in.data_ptr(), d_tg.data_ptr())
    print("train_device", n, "bunches ok", "%.1f ms" % ((time.time() - t0) * 1e3), net.stats()["device_ms"], flush=True)
kt = net.profile_kernels(16 * bunch, d_in.data_ptr(), d_tg.data_ptr())
print("profile ok", {k: v for k, v in kt.items() if isinstance(v, dict) and v["launches"]}, flush=True)
