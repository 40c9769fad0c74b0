#!/usr/bin/env python
"""bench.py -- throughput of the BPtrain_Sigmoid training step (frames/s, fwd+bwd+update, GGD-ML).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU-runnable path (C oracle)

A "step" is one bunch: `bunch` frames through forward, GGD loss gradient, backward and momentum-SGD update
(BP_GPU::train_bunch_single, BP_GPU.cu:308-440).  Workload at every N: the named network
1799-2048-2048-2048-257, 128 frames per GPU per step (weak scaling: the global minibatch is 128*N and
alpha / the gradients are allreduced), MLflag=1, shapefactor=1.5, synthetic N(0,1) frames.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (layersizes, MLflag, shapefactor, frames per GPU per step)
    "ggd_ml_1799x2048x3_257_b128": ([1799, 2048, 2048, 2048, 257], 1, 1.5, 128),
    "mmse_1799x2048x3_257_b128": ([1799, 2048, 2048, 2048, 257], 0, 2.0, 128),
    "ggd_ml_2827x2048x3_257_g1024": ([2827, 2048, 2048, 2048, 257], 1, 1.5, None),   # config 4: global 1024 / N
}
LR, MOM, WC = 0.1, 0.9, 1e-5


def flops_per_frame(ls):
    P = sum(ls[i] * ls[i + 1] for i in range(len(ls) - 1))
    P1 = ls[0] * ls[1]
    return 2 * P + 2 * P + 2 * (P - P1)          # fwd + dW + dX


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


def gemm_probe(pkg, peaks):
    """forward-layer GEMM (K = N = 2048) of gemm_tc.cu at 128 / 1024 / 4096 frames, warm, CUDA events inside the library"""
    import ctypes as C
    L = pkg.load_library()
    PF = C.POINTER(C.c_float)
    L.ggd_debug_gemm_timed.argtypes = [C.c_int] * 7 + [PF, PF, PF, C.c_int, PF, C.POINTER(C.c_ulonglong), C.c_int]
    out = {}
    rng = np.random.RandomState(0)
    for M, bn, splits in ((128, 64, 4), (1024, 128, 1), (4096, 128, 1)):
        A = rng.randn(M, 2048).astype(np.float32); B = rng.randn(2048, 2048).astype(np.float32)
        D = np.zeros((M, 2048), np.float32)
        ms = C.c_float()
        rc = L.ggd_debug_gemm_timed(0, 1, M, 2048, 2048, bn, splits, A.ctypes.data_as(PF), B.ctypes.data_as(PF), D.ctypes.data_as(PF),
                                    20, C.byref(ms), None, 0)
        if rc != 0:
            raise RuntimeError(L.ggd_last_error().decode())
        alg = 2.0 * M * 2048 * 2048 / (ms.value * 1e-3) / 1e12
        out["fwd_%dx2048x2048" % M] = {"us": ms.value * 1e3, "tflops_algorithmic": alg, "tflops_executed_bf16x3": 3 * alg,
                                        "frac_of_bf16_sustained_executed": 3 * alg / peaks["bf16_sus"]}
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.skip = gpu, [], False, 0
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def wait_first(self, timeout=3.0):
        """nvidia-smi needs ~0.5-1 s to print its first row: without this a short run (--steps 20) ends before any sample exists"""
        t0 = time.time()
        while not self.rows and time.time() - t0 < timeout and self.is_alive():
            time.sleep(0.02)
        self.skip = len(self.rows)      # rows before this point were sampled on an idle GPU: not part of the record

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in (self.rows[self.skip:] or self.rows):
            try:
                sm.append(float(r[0])); smax = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def bench_lps(pkg, torch, dev, peaks, with_cpu):
    """LPS extraction: one hour of int16 noise clip(round(N(0, 3000^2))) (224 999 frames; 115 MB in, 231 MB out:
    larger than the 126 MB L2), device-resident and e2e."""
    n = 16000 * 3600
    g = torch.Generator(device=dev); g.manual_seed(1234)
    pcm = torch.clamp(torch.round(torch.randn(n, device=dev, generator=g) * 3000.0), -32768, 32767).to(torch.int16)
    ex = pkg.Wav2LPS(dev.index or 0)
    nf = pkg.lps_nframes(n)
    out = torch.empty(nf, 257, device=dev, dtype=torch.float32)
    off = [0, n]
    for _ in range(3):
        ex.extract_batch_device(pcm.data_ptr(), off, out.data_ptr())
    ms = []
    for _ in range(10):
        ex.extract_batch_device(pcm.data_ptr(), off, out.data_ptr())
        ms.append(ex.last_kernel_ms())
    kms = float(np.median(ms))
    res = {"frames": nf, "value": nf / (kms * 1e-3), "unit": "frames/s", "kernel_ms": kms,
           "roofline": {"bound": "hbm", "achieved": 1540.0 * nf / (kms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": 1540.0 * nf / (kms * 1e-3) / 1e9 / peaks["hbm"],
                        "note": "1 540 algorithmic bytes per frame (256 int16 in + 257 fp32 out); default register-resident radix-8 FFT kernel"}}
    hp = torch.empty(n, dtype=torch.int16).pin_memory(); hp.copy_(pcm.cpu())
    ho = torch.empty(nf, 257, dtype=torch.float32).pin_memory()
    h, feats = hp.numpy(), ho.numpy()
    ex.extract(h, out=feats)                 # warm: staging buffers of the handle are sized on first use
    t0 = time.perf_counter()
    ex.extract(h, out=feats)                 # pinned host PCM -> device -> kernel -> pinned host features
    dt = time.perf_counter() - t0
    res["e2e"] = {"value": nf / dt, "unit": "frames/s", "h2d_bytes": int(h.nbytes), "d2h_bytes": int(feats.nbytes)}
    if with_cpu:
        try:
            res["cpu_baseline"] = lps_cpu_baseline(pkg, h[:16000 * 60])
        except Exception as ex2:
            res["cpu_baseline"] = {"unavailable": str(ex2)[:200]}
    ex.close()
    return res


def lps_cpu_baseline(pkg, sample):
    """the reference's Wav2LPS_be (built unmodified into oracle/_ref, its own -O0 flags and -O2), one
    process per host core on 60 s of the same noise each, aggregate frames/s; falls back to the C oracle port."""
    import tempfile
    from oracle import oracle as O, refcuda
    cores = host_cores()
    nf = pkg.lps_nframes(len(sample))
    out = {"unit": "frames/s", "cores": cores}
    bins = [("O2", "Wav2LPS_be_ref"), ("O0", "Wav2LPS_be_ref_O0")]
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    have = [(tag, os.path.join(ref_dir, b)) for tag, b in bins if os.path.exists(os.path.join(ref_dir, b))]
    if not have:
        t0 = time.perf_counter(); O.lps_extract(sample); dtc = time.perf_counter() - t0
        out.update({"value": nf / dtc, "cores": 1, "kind": "port", "sample": "60 s of noise through oracle/lps_oracle.c, one thread"})
        return out
    with tempfile.TemporaryDirectory() as td:
        raw = os.path.join(td, "in.raw")
        sample.astype("<i2").tofile(raw)
        for tag, exe in have:
            def run(n):
                ps = [subprocess.Popen([exe, "-F", "RAW", "-fs", "16", raw, os.path.join(td, "o%d.lps" % i)],
                                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in range(n)]
                t0 = time.perf_counter()
                for q in ps:
                    q.wait()
                return time.perf_counter() - t0
            t1 = run(1)
            tn = run(cores)
            out["per_core_%s" % tag] = nf / t1
            out["all_cores_%s" % tag] = cores * nf / tn
    out["value"] = out.get("all_cores_O2", out.get("all_cores_O0"))
    out["kind"] = "reference"
    out["sample"] = "60 s of the same noise per process through oracle/_ref/Wav2LPS_be_ref (reference sources, -O2; -O0 = its own makefile flags), %d processes at once, incl. file I/O" % cores
    return out


class quiet_stdout:
    """Redirects the C-level stdout (the reference library printf()s) so that the JSON line stays the only output."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null); os.close(self.saved)


def omp_threads(want=None):
    """Sets (before the oracle library is first loaded) and reports the OpenMP thread count the C oracle really runs with.
    torch.distributed.run exports OMP_NUM_THREADS=1 to every rank: without this the CPU arm would silently run on one core."""
    import ctypes
    if want:
        os.environ["OMP_NUM_THREADS"] = str(want)
    try:
        g = ctypes.CDLL("libgomp.so.1")
        if want:
            g.omp_set_num_threads(int(want))
        return int(g.omp_get_max_threads())
    except OSError:
        return 1


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_net_inputs(ls, seed=1):
    from oracle import oracle as O
    return O.init_weights(ls, seed=seed, beta=2.0)


def run_reference(args, ls, ml, beta, bunch):
    """The reference's own CPU-runnable implementation of the path = the plain-C oracle (there is no CPU trainer in
    the reference; oracle/ggd_oracle.c restates BP_GPU::train_bunch_single), OpenMP over the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = omp_threads(host_cores())
    from oracle import oracle as O
    O.build()
    W, b = make_net_inputs(ls)
    rng = np.random.RandomState(1)
    net = O.OracleNet(ls, bunch, LR, MOM, WC, beta, ml, W, b)
    x = rng.randn(bunch, ls[0]).astype(np.float32)
    t = rng.randn(bunch, ls[-1]).astype(np.float32)
    for _ in range(max(1, min(args.warmup, 2))):
        net.train_bunch(x, t)
    t0 = time.time()
    done = 0
    budget = 90.0
    while done < args.steps and (time.time() - t0) < budget:
        net.train_bunch(x, t)
        done += 1
    dt = time.time() - t0
    val = done * bunch / dt
    cores = threads
    line = {"impl": "reference", "metric": "train_frames_per_sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "timed_steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "bunch": bunch, "layersizes": ls, "MLflag": ml, "shapefactor": beta,
                       "note": "the reference has no CPU trainer and no multi-GPU path: this arm is its training step restated in C "
                               "(oracle/ggd_oracle.c) on the host cores, one bunch of %d frames per step at every N" % bunch},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "%d bunches of %d frames through oracle/ggd_oracle.c (OpenMP, omp_get_max_threads() = %d)" % (done, bunch, cores)},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def dp_parity_leg(pkg, torch, dist, world, rank, local_rank, uid_fn):
    """OUTSIDE the timed region: `world` ranks train a few sharded steps and rank 0 compares weights, alpha and the loss
    trace with the C oracle on the UNSHARDED minibatch (tolerance 1e-3 relative), plus bit-identity of the
    weights across ranks.  Two cases: a small ragged net (3 steps) and the named net (2 steps), 128 frames per GPU."""
    from oracle import oracle as O
    res = {"ok": True, "max_rel": 0.0, "bit_identical": True, "cases": []}
    for name, ls, nb in (("small_70x96x80x33", [70, 96, 80, 33], 3), ("named_1799x2048x3_257", [1799, 2048, 2048, 2048, 257], 2)):
        M = 128
        rng = np.random.RandomState(13)
        W, b = O.init_weights(ls, seed=3)
        x = rng.randn(world, nb * M, ls[0]).astype(np.float32)
        t = rng.randn(world, nb * M, ls[-1]).astype(np.float32)
        net = pkg.BP_GPU(0, local_rank, len(ls), ls, M, LR, MOM, WC, W, b, 1.5, 1, world_size=world, rank=rank, nccl_unique_id=uid_fn())
        net.train(nb * M, x[rank], t[rank])
        Wn, bn = net.returnWeights()
        alpha, losses = net.alpha(), net.losses()
        net.close()
        flat = np.concatenate([w.ravel() for w in Wn + bn]).astype(np.float32)
        mine = torch.from_numpy(flat.view(np.int32).copy()).cuda()
        lo_, hi_ = mine.clone(), mine.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo_, hi_))
        case = {"net": name, "steps": nb, "global_minibatch": M * world, "bit_identical_across_ranks": same}
        if rank == 0:
            xg = np.concatenate([x[:, i * M:(i + 1) * M].reshape(world * M, -1) for i in range(nb)])
            tg = np.concatenate([t[:, i * M:(i + 1) * M].reshape(world * M, -1) for i in range(nb)])
            orc = O.OracleNet(ls, world * M, LR, MOM, WC, 1.5, 1, W, b)
            lo, al = orc.train(xg, tg)
            Wo, bo = orc.weights()
            rel = lambda a, c: float(np.linalg.norm(np.asarray(a, np.float64) - c) / max(np.linalg.norm(np.asarray(c, np.float64)), 1e-30))
            case["rel_weights"] = max(rel(a, c) for a, c in zip(Wn + bn, Wo + bo))
            case["rel_update"] = max(rel(a - w0, c - w0) for a, c, w0 in zip(Wn, Wo, W))
            case["rel_alpha"] = rel(alpha, al[-1])
            case["rel_loss"] = float(np.max(np.abs(losses - lo) / np.abs(lo)))
            worst = max(case["rel_weights"], case["rel_alpha"], 0.5 * case["rel_update"])
            res["max_rel"] = max(res["max_rel"], worst)
            res["ok"] = res["ok"] and worst <= 1e-3 and case["rel_loss"] <= 5e-3
        res["bit_identical"] = res["bit_identical"] and same
        res["ok"] = res["ok"] and same
        res["cases"].append(case)
    return res


def config4_record(pkg, torch, dist, world, rank, local_rank, uid_fn, peaks, steps=100):
    """BASELINE.json configs[3]: 2827-2048^3-257 (ctx 11), GLOBAL minibatch 1024 sharded as 1024/world frames per GPU
    (strong scaling: the update runs replicated over all 1024 frames on every rank).  Device-resident, CUDA events."""
    ls, ml, beta, _ = WORKLOADS["ggd_ml_2827x2048x3_257_g1024"]
    bunch = 1024 // world
    dev = torch.device("cuda", local_rank)
    W, b = make_net_inputs(ls)
    net = pkg.BP_GPU(0, local_rank, len(ls), ls, bunch, LR, MOM, WC, W, b, beta, ml, world_size=world, rank=rank,
                     nccl_unique_id=uid_fn() if world > 1 else None)
    g = torch.Generator(device=dev); g.manual_seed(11 + rank)
    n = steps * bunch
    d_in = torch.randn(n, ls[0], device=dev, generator=g); d_tg = torch.randn(n, ls[-1], device=dev, generator=g)
    net.reserve(n)
    net.train_device(min(n, 16 * bunch), d_in.data_ptr(), d_tg.data_ptr())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    net.train_device(n, d_in.data_ptr(), d_tg.data_ptr())
    ms = net.stats()["device_ms"]
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    kt = net.profile_kernels(min(16, steps) * bunch, d_in.data_ptr(), d_tg.data_ptr())
    net.close()
    P = sum(ls[i] * ls[i + 1] for i in range(len(ls) - 1)) + sum(ls[1:])
    upd_ms = kt["dw_update"]["ms"] / max(kt["steps"], 1)
    fpf = flops_per_frame(ls)
    return {"workload": "ggd_ml_2827x2048x3_257_g1024", "frames_per_gpu_per_step": bunch, "global_minibatch": 1024, "steps": steps,
            "scaling": "strong", "value": steps * 1024 / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms / steps,
            "tensor_frac_whole_step": fpf * 1024 * steps / (ms * 1e-3) / 1e12 / peaks["bf16_sus"] / world,
            "roofline": {"bound": "hbm", "kernel": "dw_wide_kernel (dW over all 1024 frames + momentum update, replicated per rank)",
                         "achieved": 16.0 * P / (upd_ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                         "frac": 16.0 * P / (upd_ms * 1e-3) / 1e9 / peaks["hbm"], "traffic": None,
                         "tflops_executed_bf16x3": 3 * 2.0 * P * 1024 / (upd_ms * 1e-3) / 1e12,
                         "note": "16 B/param algorithmic; at 1024 frames the kernel is bound by the tensor pipe / operand fabric, not HBM"}}


def chain_trace(pkg, net, ls, bunch, torch, dev):
    """Device-side timeline of ONE training step in stream order with PDL (per launch: first CTA entry -> last CTA exit,
    globaltimer): kernel durations without the launch / event overhead of the per-launch event pairs."""
    import ctypes as C
    L = pkg.load_library()
    PF = C.POINTER(C.c_float)
    L.ggd_debug_trace_step.argtypes = [C.c_void_p, PF, PF, C.c_int, PF, C.c_int, C.POINTER(C.c_int)]
    rng = np.random.RandomState(0)
    x = rng.randn(bunch, ls[0]).astype(np.float32); t = rng.randn(bunch, ls[-1]).astype(np.float32)
    out = np.zeros((32, 12), np.float32); n = C.c_int()
    rc = L.ggd_debug_trace_step(net.h, x.ctypes.data_as(PF), t.ctypes.data_as(PF), 1, out.ctypes.data_as(PF), 32, C.byref(n))
    if rc != 0:
        raise RuntimeError(L.ggd_last_error().decode())
    rows = out[:n.value]
    kinds = {0: "fwd_gemm", 2: "dx_gemm", 3: "dw_gemm"}
    per = {}
    for r in rows:
        k = kinds.get(int(r[0]), "other")
        per.setdefault(k, []).append(float(r[3] - r[2]))
    return {"launch_us": {k: [round(v, 2) for v in vs] for k, vs in per.items()},
            "class_us": {k: float(sum(vs)) for k, vs in per.items()},
            "chain_us": float(rows[:, 3].max() - rows[:, 2].min()) if len(rows) else None,
            "how": "ggd_debug_trace_step: globaltimer stamps of every CTA, stream order with programmatic dependent launch"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=800)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ggd_ml_1799x2048x3_257_b128", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-lps", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the data-parallel parity leg (N > 1)")
    ap.add_argument("--no-config4", action="store_true", help="skip the config-4 sub-record")
    args = ap.parse_args()
    ls, ml, beta, bunch = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if bunch is None:
        bunch = 1024 // max(world, 1)
    if args.impl == "reference":
        return run_reference(args, ls, ml, beta, bunch)

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    from se_ml_b200.bp_gpu import nccl_unique_id
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    uid = None

    def new_uid():
        box = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        uid = new_uid()
    threads = omp_threads(host_cores())      # the C oracle (parity leg, CPU baseline) uses the host cores, not torchrun's OMP_NUM_THREADS=1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, b = make_net_inputs(ls)
    prec = 0 if args.precision == "bf16x3" else 1
    net = pkg.BP_GPU(0, local_rank, len(ls), ls, bunch, LR, MOM, WC, W, b, beta, ml, precision=prec, world_size=world, rank=rank,
                     nccl_unique_id=uid)
    # warm-up: at least 32 steps so that BOTH step graphs (16-step and single-step) have been uploaded and replayed once
    K, Wm = args.steps, max(args.warmup, 39)      # 2 x 16-step graph + 1 x 4-step graph + 3 single-step graphs: every graph uploaded
    nfr = K * bunch
    g = torch.Generator(device=dev); g.manual_seed(1 + rank)
    d_in = torch.randn(max(nfr, Wm * bunch), ls[0], device=dev, generator=g)        # synthetic frames, resident in HBM
    d_tg = torch.randn(max(nfr, Wm * bunch), ls[-1], device=dev, generator=g)
    # ---- warm-up (staging for the full chunk and the CUDA graphs are set up before anything is timed)
    net.reserve(max(nfr, Wm * bunch))
    sampler = ClockSampler(local_rank); sampler.start()      # clocks / throttle reasons from the warm-up on (the GPU is under load from here)
    sampler.wait_first()
    net.train_device(Wm * bunch, d_in.data_ptr(), d_tg.data_ptr())
    barrier()
    # ---- timed region: K steps, inputs already resident (800 steps: 842 MB of frames + 101 MB of weight state vs the 126 MB L2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    net.train_device(nfr, d_in.data_ptr(), d_tg.data_ptr())
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    st = net.stats()
    dev_ms = st["device_ms"]
    launches = int(st["launches"])
    tmax = torch.tensor([max(ms, dev_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    if ms < 150.0:
        time.sleep(0.15)          # a short timed region fits between two 100 ms samples: take the one right after it
    clocks = sampler.finish()
    value = K * bunch * world / (ms * 1e-3)

    resident_bytes = int(d_in.numel() * 4 + d_tg.numel() * 4)
    # ---- e2e: the public call with HOST (pinned) buffers: H2D of the chunk + steps + D2H of the loss trace
    e2e = None
    if not args.no_e2e:
        nho = max(nfr, Wm * bunch)
        h_in = torch.randn(nho, ls[0], generator=torch.Generator().manual_seed(7 + rank)).pin_memory()
        h_tg = torch.randn(nho, ls[-1], generator=torch.Generator().manual_seed(8 + rank)).pin_memory()
        xin, xtg = h_in.numpy(), h_tg.numpy()
        net.train(Wm * bunch, xin[:Wm * bunch], xtg[:Wm * bunch])
        barrier()
        t0 = time.perf_counter()
        net.train(nfr, xin[:nfr], xtg[:nfr])
        _ = net.losses()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        s2 = net.stats()
        e2e = {"value": K * bunch * world / float(tt.item()), "unit": "frames/s",
               "h2d_bytes_per_step": int(s2["h2d_bytes"] // K), "d2h_bytes_per_step": int(max(s2["d2h_bytes"], 8 * K) // K),
               "h2d_ms": s2["h2d_ms"], "device_ms": s2["device_ms"]}
        del h_in, h_tg
        # the drop-in's own path: PAGEABLE caller buffers (new float[] in Interface.cc:476-480) registered once by the library
        # (GGD_FLAG_PIN_HOST), one GPU only: the second call over the same buffers is what an epoch's chunks see
        if world == 1 and prec == 0:
            try:
                npg = min(nfr, 200 * bunch)
                rs = np.random.RandomState(9)
                pin_x = rs.standard_normal((npg, ls[0])).astype(np.float32); pin_t = rs.standard_normal((npg, ls[-1])).astype(np.float32)
                net2 = pkg.BP_GPU(0, local_rank, len(ls), ls, bunch, LR, MOM, WC, W, b, beta, ml, flags=pkg.FLAG_PIN_HOST)
                net2.keep_pinned(pin_x, pin_t)
                net2.train(npg, pin_x, pin_t)             # registers the buffers, sizes the staging, uploads the graphs
                t0 = time.perf_counter()
                net2.train(npg, pin_x, pin_t)
                _ = net2.losses()
                dt2 = time.perf_counter() - t0
                net2.close()
                e2e["pageable_pin_host"] = {"value": npg / dt2, "unit": "frames/s", "frames": npg,
                                            "what": "numpy (pageable) buffers + GGD_FLAG_PIN_HOST, second call over the same buffers"}
            except Exception as ex:
                e2e["pageable_pin_host"] = {"unavailable": str(ex)[:200]}

    # ---- per-kernel times (CUDA events around every launch, no graph) -> roofline of the dominant kernel
    peaks = measured_peaks()
    n_prof = min(64, d_in.shape[0] // bunch)          # bunches available in the resident chunk (small --steps)
    kt = net.profile_kernels(n_prof * bunch, d_in.data_ptr(), d_tg.data_ptr())
    steps_p = kt["steps"]
    per_step = {k: (kt[k]["ms"] / steps_p, kt[k]["launches"] // max(steps_p, 1)) for k in kt if isinstance(kt[k], dict)}
    fpf = flops_per_frame(ls)
    P = sum(ls[i] * ls[i + 1] for i in range(len(ls) - 1)); P1 = ls[0] * ls[1]
    Preal = P + sum(ls[1:])
    gemm_flops = {"fwd_gemm": 2 * P * bunch, "dw_gemm": 2 * P * bunch, "dx_gemm": 2 * (P - P1) * bunch}
    kern = {}
    # In-stream device timeline of one step (N = 1): per-launch durations without launch / event overhead.  The per-launch event
    # pairs of ggd_profile_kernels break the programmatic-dependent-launch overlap and add ~7 us per launch, so they overstate
    # the GEMM classes; where the trace exists it supplies the roofline numerators and the event-pair times stay as `ms_event_pair`.
    trace = None
    if world == 1 and prec == 0:
        try:
            trace = chain_trace(pkg, net, ls, bunch, torch, dev)
        except Exception as ex:
            trace = {"unavailable": str(ex)[:200]}
    tr_us = (trace or {}).get("class_us", {})
    for k, (msk, n) in per_step.items():
        if n == 0:
            continue
        ent = {"ms_per_step": msk, "launches_per_step": n, "timing": "CUDA-event pair around every launch (no graph)"}
        if k in tr_us and tr_us[k] > 0:
            ent["ms_event_pair"] = msk
            ent["ms_per_step"] = msk = tr_us[k] * 1e-3
            ent["timing"] = "device globaltimer, first CTA entry -> last CTA exit per launch, stream order with PDL"
        if k in gemm_flops:
            ent["tflops"] = gemm_flops[k] / (msk * 1e-3) / 1e12
            ent["frac_of_bf16_sustained"] = ent["tflops"] / peaks["bf16_sus"]
        if k == "dw_update":
            ent["gbs"] = 16.0 * Preal / (msk * 1e-3) / 1e9      # real (unpadded) weights + biases; W and delta once each way
            ent["tflops"] = 2.0 * P * bunch * world / (msk * 1e-3) / 1e12     # the update runs over the WHOLE minibatch on every rank
        if k == "update":
            ent["gbs"] = 20.0 * Preal / (msk * 1e-3) / 1e9          # read W, delta, g; write W, delta (+4 B shadows not counted)
        if k == "factor_push":
            fbytes = 4.0 * bunch * (sum((u + 63) // 64 * 64 for u in ls[:-1]) + sum((u + 63) // 64 * 64 for u in ls[1:]))
            ent["nvlink_bytes_out_per_step"] = fbytes * (world - 1)
        if k == "loss":
            ent["gbs"] = (3 * 4 * bunch * 257 + 1028) / (msk * 1e-3) / 1e9
        kern[k] = ent
    # ---- rooflines.  Every kernel class gets one (GEMMs: tensor pipe with SURVEY 8d's FLOPs; update: HBM with 16 B/param).
    # The headline `roofline` is the kernel with the largest share of the step: the forward and dE/dx launches are one kernel
    # template (gemm_tc_kernel), so they are taken together.
    roofs = {}
    prof = lambda name: os.path.join(ROOT, "profiles", name)
    named = ls == [1799, 2048, 2048, 2048, 257] and bunch == 128
    for k, e in kern.items():
        if k in gemm_flops:
            roofs[k] = {"bound": "tensor", "kernel": "gemm_tc_kernel (%s, %d launches/step)" % (k, e["launches_per_step"]), "achieved": e["tflops"],
                        "peak": peaks["bf16_sus"], "unit": "TFLOP/s", "frac": e["tflops"] / peaks["bf16_sus"], "traffic": None,
                        "peak_source": peaks["src"] + " bf16 sustained; algorithmic FLOPs (the pipe executes 3x: bf16x3)", "timing": e["timing"]}
            for cand in ("r02_%s_traffic.json" % k, "r01h_gemm_traffic.json" if k == "fwd_gemm" else ""):
                if cand and named and os.path.exists(prof(cand)):
                    gj = json.load(open(prof(cand)))
                    roofs[k]["traffic"] = gj["traffic_bytes_per_launch"]          # DRAM bytes of one launch (ncu --set full)
                    roofs[k]["traffic_source"] = gj["source"]
                    if "l2_to_sm_bytes" in gj:
                        roofs[k]["l2_to_sm_bytes_per_launch"] = gj["l2_to_sm_bytes"]
                    break
        elif "gbs" in e and k in ("dw_update", "update"):
            roofs[k] = {"bound": "hbm", "kernel": k, "achieved": e["gbs"], "peak": peaks["hbm"], "unit": "GB/s", "frac": e["gbs"] / peaks["hbm"],
                        "traffic": None, "peak_source": peaks["src"], "timing": e["timing"]}
            if k == "dw_update":
                roofs[k]["kernel"] = ("dw_persist_kernel" if (world == 1 and bunch <= 128) else "dw_wide_kernel") + \
                    " (dW GEMM + momentum update of all layers; 16 B/param algorithmic)"
                for cand in ("r02_dw_update_traffic.json", "r01h_dw_persist_traffic.json"):
                    if named and world == 1 and os.path.exists(prof(cand)):
                        tj = json.load(open(prof(cand)))
                        roofs[k]["traffic"] = tj["traffic_bytes_per_launch"]
                        roofs[k]["traffic_source"] = tj["source"]
                        break
                roofs[k]["algorithmic_bytes_per_launch"] = 16 * Preal
    roof = None
    if "fwd_gemm" in kern and "dx_gemm" in kern:
        t_chain = kern["fwd_gemm"]["ms_per_step"] + kern["dx_gemm"]["ms_per_step"]
        if trace and trace.get("chain_us"):
            t_chain = trace["chain_us"] * 1e-3      # wall time of the chain: consecutive launches overlap under PDL
        t_other = max((e["ms_per_step"] for k, e in kern.items() if k not in ("fwd_gemm", "dx_gemm")), default=0.0)
        total = sum(e["ms_per_step"] for e in kern.values())
        if t_chain >= t_other:
            tf = (gemm_flops["fwd_gemm"] + gemm_flops["dx_gemm"]) / (t_chain * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (forward + dE/dx chain, %d launches/step)" % (kern["fwd_gemm"]["launches_per_step"] + kern["dx_gemm"]["launches_per_step"]),
                    "achieved": tf, "peak": peaks["bf16_sus"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_sus"],
                    "traffic": roofs["fwd_gemm"].get("traffic"), "traffic_note": "DRAM bytes of ONE 128x2048x2048 forward launch (ncu --set full)",
                    "peak_source": roofs["fwd_gemm"]["peak_source"], "timing": kern["fwd_gemm"]["timing"],
                    "executed_frac_bf16x3": 3 * tf / peaks["bf16_sus"], "share_of_kernel_time": t_chain / max(total, 1e-12),
                    "note": "at %d frames per launch the chain is latency / weight-stream bound: 2*%d FLOP per 4-byte weight against a machine "
                            "balance of ~210 FLOP/B; gemm_tensor_probe has the same kernel at tensor-bound sizes" % (bunch, bunch)}
        else:
            dom = max((k for k in roofs if k not in ("fwd_gemm", "dx_gemm")), key=lambda k: kern[k]["ms_per_step"])
            roof = dict(roofs[dom]); roof["share_of_kernel_time"] = kern[dom]["ms_per_step"] / max(total, 1e-12)

    # ---- CPU baseline on rank 0 at N=1: the C oracle on a bounded sample of the same workload
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as O
        orc = O.OracleNet(ls, bunch, LR, MOM, WC, beta, ml, W, b)
        xs = d_in[:bunch].cpu().numpy(); ts = d_tg[:bunch].cpu().numpy()
        orc.train_bunch(xs, ts)
        t0 = time.time(); n = 0
        while n < 8 and time.time() - t0 < 20.0:
            orc.train_bunch(xs, ts); n += 1
        dtc = time.time() - t0
        cpu = {"value": n * bunch / dtc, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "%d bunches of %d frames through oracle/ggd_oracle.c (OpenMP, omp_get_max_threads() = %d)" % (n, bunch, threads)}

    # ---- the reference's own CUDA trainer (BP_GPU.cu + DevFunc.cu + cuBLAS, built unmodified into oracle/_ref) on the
    #      same GPU, same shapes: reported next to ours, not the optimisation target
    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import refcuda
            if refcuda.available("libref_bpgpu.so"):
                nrb = min(200, d_in.shape[0] // bunch)
                nref = nrb * bunch
                xr = d_in[:nref].cpu().numpy(); tr = d_tg[:nref].cpu().numpy()
                with quiet_stdout():
                    ref = refcuda.RefBPGPU(ls, bunch, LR, MOM, WC, beta, ml, W, b, gpu=local_rank)
                    ref.train(xr[:min(8, nrb) * bunch], tr[:min(8, nrb) * bunch])
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    ref.train(xr, tr)
                    torch.cuda.synchronize()
                    dtr = time.perf_counter() - t0
                    ref.close()
                ref_cuda = {"value": nref / dtr, "unit": "frames/s", "ms_per_step": 1e3 * dtr / nrb,
                            "what": "reference BP_GPU::train (fp32 cuBLAS SGEMM, 53 launches/bunch), %d bunches incl. its H2D copy" % nrb}
        except Exception as ex:     # the baseline is optional; never let it break the measurement
            ref_cuda = {"unavailable": str(ex)[:200]}

    # ---- tensor-pipe evidence: the same tcgen05 GEMM kernel on the forward shape at bunch sizes where it is not
    #      latency-bound (the 128-frame step cannot leave the HBM floor); algorithmic and executed
    #      (bf16x3 = 3 MMAs per product) TFLOP/s against the measured bf16 peak
    probe = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            probe = gemm_probe(pkg, peaks)
        except Exception as ex:
            probe = {"unavailable": str(ex)[:200]}

    # ---- data-parallel parity leg and the config-4 record (outside the timed region; every rank takes part)
    dp_parity = None
    if world > 1 and not args.no_parity:
        try:
            dp_parity = dp_parity_leg(pkg, torch, dist, world, rank, local_rank, new_uid)
        except Exception as ex:
            dp_parity = {"ok": False, "error": str(ex)[:300]}
    cfg4 = None
    if not args.no_config4 and args.workload != "ggd_ml_2827x2048x3_257_g1024":
        try:
            cfg4 = config4_record(pkg, torch, dist, world, rank, local_rank, new_uid, peaks)
        except Exception as ex:
            cfg4 = {"error": str(ex)[:300]}

    # ---- LPS front end (second half of the metric): frames/s of the extraction kernel on synthetic 16 kHz noise
    lps = None
    if rank == 0 and not args.no_lps:
        lps = bench_lps(pkg, torch, dev, peaks, world == 1 and not args.no_cpu_baseline)

    if rank == 0:
        line = {"metric": "train_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16x3 (bf16 hi+lo operands, fp32 TMEM accumulate, fp32 master weights)" if prec == 0 else "f32",
                "data": "synthetic",
                "config": {"workload": args.workload, "layersizes": ls, "MLflag": ml, "shapefactor": beta, "frames_per_gpu_per_step": bunch,
                           "global_minibatch": bunch * world,
                           "l2": "every step reads %d new input bytes and streams the %.0f MB of weights + momentum; resident chunk %.0f MB vs the 126 MB L2%s"
                                 % (bunch * (ls[0] + ls[-1]) * 4, 8.0 * Preal / 1e6, resident_bytes / 1e6,
                                    "" if resident_bytes + 8 * Preal > 126e6 else " (short run: raise --steps for a chunk larger than L2)"),
                           "parallelism": "dp%d (frames sharded; sum|e|^beta and the gradient FACTORS exchanged over NVLink peer memory, update replicated)" % world},
                "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roof, "roofline_all": roofs, "cpu_baseline": cpu,
                "kernels": kern, "step_trace": trace, "dp_parity": dp_parity, "config4": cfg4, "gemm_tensor_probe": probe, "flops_per_frame": fpf, "reference_cuda": ref_cuda, "lps": lps,
                "tensor_frac_whole_step": (fpf * bunch * K / (ms * 1e-3) / 1e12) / peaks["bf16_sus"]}
        print(json.dumps(line), flush=True)
    net.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
