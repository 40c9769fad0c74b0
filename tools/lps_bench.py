"""LPS extraction throughput only (device-resident, e2e with pinned host buffers): the `lps` object of the bench line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from __graft_entry__ import load_pkg  # noqa: E402

pkg = load_pkg()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
print(json.dumps(bench.bench_lps(pkg, torch, dev, bench.measured_peaks(), "--cpu" in sys.argv)))
