"""BASELINE config 5: LPS extraction + one GGD-ML epoch of the DROP-IN EXECUTABLE on a synthetic pfile pair, on N GPUs.

    python tools/config5.py --gpus 8 --frames 22500000 [--hours 100] [--out gpurun_out/r02_config5.json]

Stage 1 (LPS): `hours` of synthetic 16 kHz int16 noise, utterance-sharded over N processes (one per GPU), each through the
public host-buffer call lps_extract_batch (pinned host PCM -> device -> kernel -> pinned host features); wall clock of the
slowest rank.  No collective: utterances are independent.
Stage 2 (epoch): host/BPtrain_Sigmoid with finetune.pl's flags and gpu_used=0,...,N-1 (frame-sharded, bunchsize = GLOBAL
minibatch) on a synthetic pfile pair of `frames` frames in /dev/shm: wall clock of the whole process (CUDA start-up, NCCL
set-up, pfile reading, epoch, weight file, CV) and of the 'Total cost time' window.
The pfile pair is synthetic (one block of N(1,3^2) features repeated with fresh sentence / frame indices): the arithmetic of the
epoch does not depend on the values, and writing 2 x 23 GB of fresh random numbers would dominate the run.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "speech-enhancement-based-on-a-maximum-likelihood-criterion_b200")


def lps_rank(rank, world, hours, q):
    os.environ["CUDA_VISIBLE_DEVICES"] = str(rank)
    import torch
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    # this rank's share of the utterances: `hours`/world of audio as 10 s utterances
    n_utt = max(1, int(hours * 360 / world))
    utt = 160000
    rng = np.random.RandomState(1234 + rank)
    block = np.clip(np.round(rng.randn(utt * 36) * 3000), -32768, 32767).astype(np.int16)     # 6 min of noise, tiled
    ex = pkg.Wav2LPS(0)
    piece = 360                     # utterances per call (1 h): bounded host / device buffers
    hp = torch.empty(piece * utt, dtype=torch.int16).pin_memory()
    hp.numpy()[:] = np.tile(block, piece // 36)
    off = np.arange(piece + 1, dtype=np.int64) * utt
    nf_piece = piece * pkg.lps_nframes(utt)
    ho = torch.empty(nf_piece, 257, dtype=torch.float32).pin_memory()
    L = pkg.load_library()
    import ctypes as C
    PS, PL, PF = C.POINTER(C.c_int16), C.POINTER(C.c_long), C.POINTER(C.c_float)
    pcm, out = hp.numpy(), ho.numpy()
    def call(nu):
        n = C.c_long()
        rc = L.lps_extract_batch(ex.h, pcm.ctypes.data_as(PS), off.ctypes.data_as(PL), nu, out.ctypes.data_as(PF), 0, C.byref(n))
        assert rc == 0, L.lps_last_error()
        return n.value
    call(piece)                     # warm: staging buffers sized
    q.put(("ready", rank))
    done, frames, kms = 0, 0, 0.0
    t0 = time.perf_counter()
    while done < n_utt:
        nu = min(piece, n_utt - done)
        frames += call(nu); kms += ex.last_kernel_ms(); done += nu
    dt = time.perf_counter() - t0
    q.put(("done", rank, frames, dt, kms))


def stage_lps(world, hours):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=lps_rank, args=(r, world, hours, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = []
    while len(res) < world:
        m = q.get(timeout=900)
        if m[0] == "done":
            res.append(m)
    for p in ps:
        p.join(60)
    frames = sum(m[2] for m in res); wall = max(m[3] for m in res); kms = max(m[4] for m in res)
    return {"hours_of_audio": hours, "frames": frames, "wall_s_slowest_rank": round(wall, 3), "frames_per_s_wall": frames / wall,
            "kernel_ms_slowest_rank": round(kms, 2), "frames_per_s_kernels_only": frames / (kms * 1e-3),
            "what": "lps_extract_batch, pinned host PCM in / pinned host features out, 1 h of 10 s utterances per call, default (fast) kernel"}


_GEN = {}


def _write_slice(job):
    name, p0, p1 = job
    g = _GEN
    blk = g["blk"]
    rec = np.zeros((blk, 259), ">i4")
    rec[:, 2:] = g[name]
    fd = os.open("%s/%s.pfile" % (g["T"], name), os.O_WRONLY)
    for q0 in range(p0, p1, blk):
        n = min(blk, p1 - q0)
        rec[:n, 0] = g["sent"][q0:q0 + n]; rec[:n, 1] = g["fidx"][q0:q0 + n]
        os.pwrite(fd, rec[:n].tobytes(), 32768 + q0 * 259 * 4)
    os.close(fd)
    return p1 - p0


def make_pfiles(T, frames, workers=16):
    from oracle import oracle as O
    rng = np.random.RandomState(5)
    blk = 1 << 18
    noisy = (rng.randn(blk, 257) * 3 + 1).astype(np.float32)
    clean = (noisy * 0.7 + rng.randn(blk, 257).astype(np.float32)).astype(np.float32)
    lens, left = [], frames
    lr = rng.randint(100, 400, size=frames // 100 + 10)
    for n in lr:
        n = int(min(left, n)); lens.append(n); left -= n
        if left == 0:
            break
    sent = np.repeat(np.arange(len(lens), dtype=np.int32), lens)
    first = np.concatenate([[0], np.cumsum(lens)])[:-1]
    fidx = (np.arange(frames, dtype=np.int64) - np.repeat(first, lens)).astype(np.int32)
    tail = np.concatenate([[0], np.cumsum(lens)]).astype(">i4").tobytes()
    for name in ("noisy", "clean"):
        hdr = ("-pfile_header version 0 size 32768\n-num_sentences %d\n-num_frames %d\n-first_feature_column 2\n-num_features 257\n"
               "-first_label_column 259\n-num_labels 0\n-format dd%s\n-data size %d offset 0 ndim 2 nrow %d ncol 259\n"
               "-sent_table_data size %d offset %d ndim 1\n-end\n") % (len(lens), frames, "f" * 257, frames * 259, frames, len(lens) + 1, frames * 259)
        with open("%s/%s.pfile" % (T, name), "wb") as f:
            f.write(hdr.encode("ascii").ljust(32768, b"\0"))
            f.truncate(32768 + frames * 259 * 4)
            f.seek(32768 + frames * 259 * 4)
            f.write(tail)
    # the records: slices written in parallel by forked workers (they inherit the arrays)
    _GEN.update({"T": T, "blk": blk, "sent": sent, "fidx": fidx, "noisy": noisy.astype(">f4").view(">i4"), "clean": clean.astype(">f4").view(">i4")})
    per = (frames + workers // 2 - 1) // (workers // 2)
    jobs = [(name, p0, min(frames, p0 + per)) for name in ("noisy", "clean") for p0 in range(0, frames, per)]
    with mp.get_context("fork").Pool(workers) as pool:
        pool.map(_write_slice, jobs)
    O.write_norm(T + "/noisy.norm", noisy.mean(0), 1.0 / noisy.std(0))
    return lens


def stage_epoch(world, frames, bunch, cache):
    from oracle import oracle as O
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    T = tempfile.mkdtemp(dir=base)
    t0 = time.time()
    try:
        lens = make_pfiles(T, frames)
    except OSError as ex:            # /dev/shm too small for 2 x 23 GB: fall back to half the frames and say so
        for f in os.listdir(T):
            os.remove(os.path.join(T, f))
        frames //= 2
        sys.stderr.write("pfile generation failed (%s): retrying with %d frames\n" % (ex, frames))
        lens = make_pfiles(T, frames)
    gen_s = time.time() - t0
    ls = [1799, 2048, 2048, 2048, 257]
    W, b = O.init_weights(ls, seed=4)
    O.write_wts(T + "/init.wts", ls, W, b)
    ncv = max(2, min(200, len(lens) // 50))
    flags = ("gpu_used=%s numlayers=5 layersizes=%s bunchsize=%d MLflag=1 shapefactor=1.5 momentum=0.9 weightcost=0.00001 lrate=0.01 "
             "fea_dim=257 fea_context=7 traincache=%d init_randem_seed=27870775 targ_offset=3 initwts_file=%s/init.wts norm_file=%s/noisy.norm "
             "fea_file=%s/noisy.pfile targ_file=%s/clean.pfile train_sent_range=0-%d cv_sent_range=%d-%d dropoutflag=0 visible_omit=0.1 hid_omit=0.1 "
             "outwts_file=%s/out.wts log_file=%s/train.log"
             % (",".join(map(str, range(world))), ",".join(map(str, ls)), bunch, cache, T, T, T, T, len(lens) - ncv - 1, len(lens) - ncv, len(lens) - 1, T, T)).split()
    train = sum(max(0, n - 6) for n in lens[:len(lens) - ncv])
    exe = os.path.join(PKG, "host", "BPtrain_Sigmoid")
    t0 = time.time()
    p = subprocess.run([exe] + flags, stdout=subprocess.PIPE, stderr=subprocess.STDOUT if not os.environ.get("GGD_CLI_TIMING") else None, timeout=3000)
    wall = time.time() - t0
    log = open(T + "/train.log").read() if os.path.exists(T + "/train.log") else ""
    cost = [l for l in log.splitlines() if l.startswith("Total cost time")]
    cv = [l.strip() for l in log.splitlines() if l.startswith("CV")]
    res = {"rc": p.returncode, "pfile_frames": frames, "pfile_bytes_each": 32768 + frames * 259 * 4, "pfile_dir": base or "tmp", "pfile_generation_s": round(gen_s, 1),
           "train_samples": train, "gpus": world, "global_minibatch": bunch, "traincache": cache, "wall_s_whole_process": round(wall, 2),
           "samples_per_s_wall": train / wall, "total_cost_time_line": cost[-1] if cost else None, "cv": cv,
           "what": "host/BPtrain_Sigmoid, finetune.pl flags, gpu_used=0..N-1 (forked workers, frame-sharded), device-side loader, one epoch + weight file + CV"}
    if cost:
        try:
            res["samples_per_s_train_window"] = train / float(cost[-1].split(":")[1].split()[0])
        except Exception:
            pass
    if p.returncode != 0:
        res["out_tail"] = p.stdout.decode("utf-8", "ignore")[-600:]; res["log_tail"] = log[-600:]
    for f in os.listdir(T):
        os.remove(os.path.join(T, f))
    os.rmdir(T)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--frames", type=int, default=5600000)
    ap.add_argument("--hours", type=float, default=25.0)
    ap.add_argument("--bunch", type=int, default=0, help="GLOBAL minibatch (default 128 per GPU)")
    ap.add_argument("--cache", type=int, default=200000)
    ap.add_argument("--skip-lps", action="store_true")
    ap.add_argument("--gen-only", action="store_true", help="only generate (and time) the synthetic pfile pair")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.gen_only:
        T = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        t0 = time.time(); lens = make_pfiles(T, a.frames)
        from oracle import oracle as O
        f = O.read_pfile(T + "/clean.pfile")
        print("generated", a.frames, "frames in %.1f s" % (time.time() - t0), f[0].shape, bool(np.isfinite(f[0]).all()), len(lens))
        for x in os.listdir(T):
            os.remove(os.path.join(T, x))
        os.rmdir(T)
        return
    res = {"config": "BASELINE.json configs[4] (synthetic 100 h-scale pfile, LPS extraction + GGD-ML epoch)", "gpus": a.gpus}
    if not a.skip_lps:
        res["lps"] = stage_lps(a.gpus, a.hours)
    res["epoch"] = stage_epoch(a.gpus, a.frames, a.bunch or 128 * a.gpus, a.cache)
    s = json.dumps(res)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
