"""End-to-end timing of the DROP-IN EXECUTABLES on a synthetic pfile pair (GPU box): this repository's
host/BPtrain_Sigmoid (device-side loader by default, host_loader=1 for the CPU expansion) next to the reference's own
CUDA trainer built unmodified into oracle/_ref/BPtrain_ref.  Same finetune.pl flags, same files, one epoch + CV.
Wall clock of the whole process (CUDA start-up included) and frames/s from the 'Total cost time' window are printed
as one JSON line; final weights are compared (same shuffles, same arithmetic up to the stated precision)."""
import json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
ls = [1799, 2048, 2048, 2048, 257]
T = tempfile.mkdtemp()
rng = np.random.RandomState(5)
lens, left = [], frames
while left > 0:
    n = min(left, int(rng.randint(100, 400))); lens.append(n); left -= n
noisy = (rng.randn(frames, 257) * 3 + 1).astype(np.float32)
clean = (noisy * 0.7 + rng.randn(frames, 257)).astype(np.float32)
O.write_pfile(T + "/noisy.pfile", noisy, lens); O.write_pfile(T + "/clean.pfile", clean, lens)
O.write_norm(T + "/noisy.norm", noisy.mean(0), 1.0 / noisy.std(0))
W, b = O.init_weights(ls, seed=4)
O.write_wts(T + "/init.wts", ls, W, b)
ncv = max(2, len(lens) // 50)
flags = ("gpu_used=0 numlayers=5 layersizes=%s bunchsize=128 MLflag=1 shapefactor=1.5 momentum=0.9 weightcost=0.00001 lrate=0.01 "
         "fea_dim=257 fea_context=7 traincache=102400 init_randem_seed=27870775 targ_offset=3 initwts_file=%s/init.wts norm_file=%s/noisy.norm "
         "fea_file=%s/noisy.pfile targ_file=%s/clean.pfile train_sent_range=0-%d cv_sent_range=%d-%d dropoutflag=0 visible_omit=0.1 hid_omit=0.1"
         % (",".join(map(str, ls)), T, T, T, T, len(lens) - ncv - 1, len(lens) - ncv, len(lens) - 1)).split()
train_frames = sum(max(0, n - 6) for n in lens[:len(lens) - ncv])
res = {"train_samples": train_frames, "pfile_frames": frames}
exes = {"ours_device_loader": ([os.path.join(ROOT, "speech-enhancement-based-on-a-maximum-likelihood-criterion_b200/host/BPtrain_Sigmoid")], []),
        "ours_host_loader": ([os.path.join(ROOT, "speech-enhancement-based-on-a-maximum-likelihood-criterion_b200/host/BPtrain_Sigmoid")], ["host_loader=1"]),
        }
for name, (exe, extra) in exes.items():
    if not os.path.exists(exe[0]):
        res[name] = {"unavailable": exe[0]}; continue
    t0 = time.time()
    p = subprocess.run(exe + flags + extra + ["outwts_file=%s/%s.wts" % (T, name), "log_file=%s/%s.log" % (T, name)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500)
    dt = time.time() - t0
    log = open("%s/%s.log" % (T, name)).read() if os.path.exists("%s/%s.log" % (T, name)) else ""
    cv = [l.strip() for l in log.splitlines() if l.startswith("CV")]
    res[name] = {"rc": p.returncode, "wall_s": round(dt, 2), "samples_per_s_wall": round(train_frames / dt), "cv": cv}
    if p.returncode != 0:
        res[name]["log_tail"] = log[-400:]; res[name]["out_tail"] = p.stdout.decode("utf-8", "ignore")[-300:]
# The reference's full binary traps on this toolchain (BPtrain.cc's threadFetch falls off a non-void function, see
# tests/test_vs_reference_gpu.py), so its two halves, each compiled UNMODIFIED, are driven in BPtrain.cc's order:
# Interface (loader) + BP_GPU (device); the loader and the trainer run back to back like the reference's main thread sees them.
try:
    from oracle import refcuda
    if refcuda.available("libref_bpgpu.so") and refcuda.available("libref_interface.so"):
        kw = dict(f.split("=", 1) for f in flags + ["outwts_file=%s/ref.wts" % T, "log_file=%s/ref.log" % T])
        t0 = time.time()
        rif = refcuda.RefInterface(**kw)
        nch, ns = rif.train_info(kw["train_sent_range"])
        rbp = refcuda.RefBPGPU(ls, 128, 0.01, 0.9, 1e-5, 1.5, 1, W, b)
        t_load = t_train = 0.0
        for ci in rif.shuffle_chunks(nch):
            t1 = time.time(); xr, tr = rif.read_chunk(ci, 1799, 257); t2 = time.time()
            rbp.train(xr, tr); t3 = time.time()
            t_load += t2 - t1; t_train += t3 - t2
        dt = time.time() - t0
        res["reference_cuda_composed"] = {"wall_s": round(dt, 2), "samples_per_s_wall": round(ns / dt), "loader_s": round(t_load, 2),
                                          "train_s": round(t_train, 2), "samples": ns}
except Exception as ex:
    res["reference_cuda_composed"] = {"unavailable": str(ex)[:200]}
try:
    a = np.fromfile(T + "/ours_device_loader.wts", np.uint8); c = np.fromfile(T + "/ours_host_loader.wts", np.uint8)
    res["device_loader_equals_host_loader_bitwise"] = bool(a.size == c.size and np.array_equal(a, c))
except Exception as ex:
    res["compare"] = str(ex)
print(json.dumps(res))
