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
        self.gpu, self.rows, self.stop_flag = gpu, [], False
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

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
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
            from oracle import oracle as O, refcuda
            sample = h[:16000 * 60]
            if refcuda.available("Wav2LPS_be_ref"):
                with quiet_stdout():
                    _, dtc = refcuda.ref_wav2lps(sample)
                kind = "reference"
            else:
                t0 = time.perf_counter(); O.lps_extract(sample); dtc = time.perf_counter() - t0
                kind = "port"
            res["cpu_baseline"] = {"value": pkg.lps_nframes(len(sample)) / dtc, "unit": "frames/s", "cores": 1, "kind": kind,
                                   "sample": "60 s of the same noise through %s (one process, incl. file I/O)" % ("oracle/_ref/Wav2LPS_be_ref (-O2)" if kind == "reference" else "oracle/lps_oracle.c")}
        except Exception as ex2:
            res["cpu_baseline"] = {"unavailable": str(ex2)[:200]}
    ex.close()
    return res


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


def make_net_inputs(ls, seed=1):
    from oracle import oracle as O
    return O.init_weights(ls, seed=seed, beta=2.0)


def run_reference(args, ls, ml, beta, bunch):
    """The reference's own CPU-runnable implementation of the path = the plain-C oracle (there is no CPU trainer in
    the reference; oracle/ggd_oracle.c restates BP_GPU::train_bunch_single), OpenMP over the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
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
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": "train_frames_per_sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "timed_steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "bunch": bunch, "layersizes": ls, "MLflag": ml, "shapefactor": beta},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "%d bunches of %d frames through oracle/ggd_oracle.c (OpenMP, %d threads)" % (done, bunch, cores)},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        box = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, b = make_net_inputs(ls)
    prec = 0 if args.precision == "bf16x3" else 1
    net = pkg.BP_GPU(0, local_rank, len(ls), ls, bunch, LR, MOM, WC, W, b, beta, ml, precision=prec, world_size=world, rank=rank,
                     nccl_unique_id=uid)
    K, Wm = args.steps, max(args.warmup, 3)
    nfr = K * bunch
    g = torch.Generator(device=dev); g.manual_seed(1 + rank)
    d_in = torch.randn(max(nfr, Wm * bunch), ls[0], device=dev, generator=g)        # synthetic frames, resident in HBM
    d_tg = torch.randn(max(nfr, Wm * bunch), ls[-1], device=dev, generator=g)
    # ---- warm-up (staging for the full chunk and the CUDA graphs are set up before anything is timed)
    net.reserve(max(nfr, Wm * bunch))
    net.train_device(Wm * bunch, d_in.data_ptr(), d_tg.data_ptr())
    barrier()
    # ---- timed region: K steps, inputs already resident; the input set (737 MB + weights) exceeds the 126 MB L2
    sampler = ClockSampler(local_rank); sampler.start()
    time.sleep(0.25)
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
    clocks = sampler.finish()
    value = K * bunch * world / (ms * 1e-3)

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

    # ---- per-kernel times (CUDA events around every launch, no graph) -> roofline of the dominant kernel
    peaks = measured_peaks()
    n_prof = min(64, d_in.shape[0] // bunch)          # bunches available in the resident chunk (small --steps)
    kt = net.profile_kernels(n_prof * bunch, d_in.data_ptr(), d_tg.data_ptr())
    steps_p = kt["steps"]
    per_step = {k: (kt[k]["ms"] / steps_p, kt[k]["launches"] // max(steps_p, 1)) for k in kt if isinstance(kt[k], dict)}
    fpf = flops_per_frame(ls)
    P = sum(ls[i] * ls[i + 1] for i in range(len(ls) - 1)); P1 = ls[0] * ls[1]
    gemm_flops = {"fwd_gemm": 2 * P * bunch, "dw_gemm": 2 * P * bunch, "dx_gemm": 2 * (P - P1) * bunch}
    kern = {}
    for k, (msk, n) in per_step.items():
        if n == 0:
            continue
        ent = {"ms_per_step": msk, "launches_per_step": n}
        if k in gemm_flops:
            ent["tflops"] = gemm_flops[k] / (msk * 1e-3) / 1e12
            ent["frac_of_bf16_sustained"] = ent["tflops"] / peaks["bf16_sus"]
        if k == "dw_update":
            Ppad = kt["param_elems"]      # padded weights + biases; the fused kernel streams W and delta once each way
            ent["gbs"] = 16.0 * Ppad / (msk * 1e-3) / 1e9
            ent["tflops"] = gemm_flops["dw_gemm"] / (msk * 1e-3) / 1e12
        if k == "update":
            if world > 1:
                # push-model owner update (reduce_update_kernel): each rank updates 1/world of the parameters: read W, delta and `world`
                # partial gradient tiles, write W, delta and the local copy of the shadows; the NVLink bytes are not HBM bytes of this rank
                ent["gbs"] = (16.0 + 4.0 * world + 4.0) * kt["param_elems"] / world / (msk * 1e-3) / 1e9
                ent["nvlink_gbs_out"] = 4.0 * kt["param_elems"] * (world - 1) / world / (msk * 1e-3) / 1e9
            else:
                ent["gbs"] = 20.0 * kt["param_elems"] / (msk * 1e-3) / 1e9     # read W, delta, g; write W, delta (+4 B shadows not counted)
        if k == "dw_gemm" and world > 1:
            ent["nvlink_gbs_out"] = 4.0 * kt["param_elems"] * (world - 1) / world / (msk * 1e-3) / 1e9   # gradient tiles pushed to their owners
        if k == "loss":
            ent["gbs"] = (3 * 4 * bunch * 257 + 1028) / (msk * 1e-3) / 1e9
        kern[k] = ent
    # ---- rooflines.  Every kernel class gets one (GEMMs: tensor pipe with SURVEY 8d's FLOPs; update: HBM with 16 B/param); the
    # headline `roofline` is the kernel SYMBOL with the largest share of the step (the forward class is two symbols: L-2 sigmoid
    # launches + 1 output-layer launch), which is what the committed ncu launch list shows as its top line.
    def symbol_share(k):
        e = kern[k]
        return e["ms_per_step"] * ((e["launches_per_step"] - 1) / e["launches_per_step"] if k == "fwd_gemm" and e["launches_per_step"] > 1 else 1.0)
    roofs = {}
    for k, e in kern.items():
        if k in gemm_flops:
            roofs[k] = {"bound": "tensor", "kernel": "gemm_tc_kernel (%s, %d launches/step)" % (k, e["launches_per_step"]), "achieved": e["tflops"],
                        "peak": peaks["bf16_sus"], "unit": "TFLOP/s", "frac": e["tflops"] / peaks["bf16_sus"], "traffic": None,
                        "peak_source": peaks["src"] + " bf16 sustained; algorithmic FLOPs (the pipe executes 3x: bf16x3)",
                        "traffic_source": None,
                        "note": "128-frame GEMMs are bound by the weight stream, not the tensor pipe: 2*128 FLOP per 4-byte weight = 64 FLOP/B against "
                                "a machine balance of ~210 FLOP/B; see gemm_tensor_probe for the same kernel at tensor-bound sizes"}
            gp = os.path.join(ROOT, "profiles", "r01h_gemm_traffic.json")
            if k == "fwd_gemm" and os.path.exists(gp) and ls == [1799, 2048, 2048, 2048, 257]:
                gj = json.load(open(gp))
                roofs[k]["traffic"] = gj["traffic_bytes_per_launch"]          # DRAM bytes of one 128x2048x2048 launch (ncu --set full)
                roofs[k]["traffic_source"] = gj["source"]
                roofs[k]["l2_to_sm_bytes_per_launch"] = gj["l2_to_sm_bytes"]
        elif "gbs" in e and k in ("dw_update", "update"):
            roofs[k] = {"bound": "hbm", "kernel": k, "achieved": e["gbs"], "peak": peaks["hbm"], "unit": "GB/s", "frac": e["gbs"] / peaks["hbm"],
                        "traffic": None, "peak_source": peaks["src"]}
            if k == "dw_update":
                roofs[k]["kernel"] = "dw_persist_kernel (dW GEMM + momentum update of all layers; 16 B/param algorithmic)"
                tp = os.path.join(ROOT, "profiles", "r01h_dw_persist_traffic.json")
                if os.path.exists(tp) and ls == [1799, 2048, 2048, 2048, 257]:
                    tj = json.load(open(tp))
                    roofs[k]["traffic"] = tj["traffic_bytes_per_launch"]
                    roofs[k]["traffic_source"] = tj["source"]
                roofs[k]["algorithmic_bytes_per_launch"] = 16 * kt["param_elems"]
    dom = max(roofs, key=symbol_share) if roofs else None
    roof = dict(roofs[dom]) if dom else None
    if roof is not None:
        roof["share_of_kernel_time"] = symbol_share(dom) / max(sum(e["ms_per_step"] for e in kern.values()), 1e-12)

    # ---- CPU baseline on rank 0 at N=1: the C oracle on a bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        orc = O.OracleNet(ls, bunch, LR, MOM, WC, beta, ml, W, b)
        xs = d_in[:bunch].cpu().numpy(); ts = d_tg[:bunch].cpu().numpy()
        orc.train_bunch(xs, ts)
        t0 = time.time(); n = 0
        while n < 8 and time.time() - t0 < 20.0:
            orc.train_bunch(xs, ts); n += 1
        dtc = time.time() - t0
        cpu = {"value": n * bunch / dtc, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "%d bunches of %d frames through oracle/ggd_oracle.c (OpenMP)" % (n, bunch)}

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
                           "global_minibatch": bunch * world, "l2": "inputs (737 MB/chunk) and weight state (250 MB) exceed the 126 MB L2",
                           "parallelism": "dp%d (frame-sharded; allreduce of sum|e|^beta and of the gradients)" % world},
                "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roof, "roofline_all": roofs, "cpu_baseline": cpu,
                "kernels": kern, "gemm_tensor_probe": probe, "flops_per_frame": fpf, "reference_cuda": ref_cuda, "lps": lps,
                "tensor_frac_whole_step": (fpf * bunch * K / (ms * 1e-3) / 1e12) / peaks["bf16_sus"]}
        print(json.dumps(line), flush=True)
    net.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
