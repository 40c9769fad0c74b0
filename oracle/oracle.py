"""ctypes wrapper around the CPU oracle (oracle/libggd_oracle.so) plus numpy restatements of the
reference's host-side file formats and loader.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (the *_b200 package) never imports this.

Reference citations (relative to /root/reference):
  pfile reader / chunking / context expansion / shuffle   Train_code_ML_GGD/Interface.cc:519-838, 975-1024
  norm file                                                Interface.cc:374-399
  MAT-v4 weight file                                       Interface.cc:430-467 (read), 484-516 (write)
  HTK feature file                                         Feature_prepare/SourceCode_Wav2LogSpec_be/fileio.c:187-243
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
PF = C.POINTER(C.c_float)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libggd_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.ggd_oracle_create.restype = C.c_void_p
        L.ggd_oracle_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_float, C.c_float, C.c_float,
                                        C.c_float, C.c_int, C.POINTER(PF), C.POINTER(PF)]
        L.ggd_oracle_destroy.argtypes = [C.c_void_p]
        L.ggd_oracle_train_bunch.argtypes = [C.c_void_p, C.c_int, PF, PF]
        L.ggd_oracle_train_bunch_sharded.argtypes = [C.c_void_p, C.c_int, C.c_int, PF, PF]
        L.ggd_oracle_train.restype = C.c_int
        L.ggd_oracle_train.argtypes = [C.c_void_p, C.c_int, PF, PF, PF, PF]
        L.ggd_oracle_forward.argtypes = [C.c_void_p, C.c_int, PF, PF]
        for f in ("ggd_oracle_cv_sqerr", "ggd_oracle_cv_abserr", "ggd_oracle_cv_loglik"):
            getattr(L, f).restype = C.c_float
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, PF, PF]
        L.ggd_oracle_gamma.restype = C.c_float
        L.ggd_oracle_gamma.argtypes = [C.c_float]
        for f in ("ggd_oracle_W", "ggd_oracle_b", "ggd_oracle_dW", "ggd_oracle_dedx", "ggd_oracle_grad", "ggd_oracle_y"):
            getattr(L, f).restype = PF
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        for f in ("ggd_oracle_alpha", "ggd_oracle_out"):
            getattr(L, f).restype = PF
            getattr(L, f).argtypes = [C.c_void_p]
        L.ggd_oracle_dp_colsum.argtypes = [C.c_void_p, C.c_int, C.c_int, PF, PF, PF]
        L.ggd_oracle_dp_backward.argtypes = [C.c_void_p, C.c_int, C.c_int, PF, PF, PF, PF]
        L.ggd_oracle_dp_update.argtypes = [C.c_void_p, C.c_int]
        L.ggd_oracle_bias_grad.restype = PF
        L.ggd_oracle_bias_grad.argtypes = [C.c_void_p, C.c_int]
        L.ggd_oracle_last_loss.restype = C.c_float
        L.ggd_oracle_last_loss.argtypes = [C.c_void_p]
        L.lps_oracle_nframes.restype = C.c_long
        L.lps_oracle_nframes.argtypes = [C.c_long]
        L.lps_oracle_extract.restype = C.c_long
        L.lps_oracle_extract.argtypes = [C.POINTER(C.c_int16), C.c_long, PF]
        L.lps_oracle_rfft.argtypes = [PF, C.c_int, C.c_int]
        L.lps_oracle_hamming.argtypes = [PF]
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(PF)


class OracleNet:
    """Mirror of the reference BP_GPU object (BP_GPU.h:45-70) on the CPU oracle."""

    def __init__(self, layersizes, bunchsize, lrate, momentum, weightcost, shapefactor, MLflag, W, b):
        self.L = lib()
        self.layersizes = list(layersizes)
        self.bunchsize = bunchsize
        n = len(layersizes)
        ls = (C.c_int * n)(*layersizes)
        self._keep = [np.ascontiguousarray(w, dtype=np.float32) for w in W] + \
                     [np.ascontiguousarray(x, dtype=np.float32) for x in b]
        Wp = (PF * n)()
        bp = (PF * n)()
        for l in range(1, n):
            Wp[l] = _fp(self._keep[l - 1])
            bp[l] = _fp(self._keep[n - 1 + l - 1])
        self.h = self.L.ggd_oracle_create(n, ls, bunchsize, lrate, momentum, weightcost, shapefactor, MLflag, Wp, bp)

    def close(self):
        if self.h:
            self.L.ggd_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def D(self):
        return self.layersizes[-1]

    def train_bunch(self, x, t):
        x = np.ascontiguousarray(x, np.float32); t = np.ascontiguousarray(t, np.float32)
        self.L.ggd_oracle_train_bunch(self.h, x.shape[0], _fp(x), _fp(t))

    def train_bunch_sharded(self, world, x, t):
        x = np.ascontiguousarray(x, np.float32); t = np.ascontiguousarray(t, np.float32)
        assert x.shape[0] % world == 0
        self.L.ggd_oracle_train_bunch_sharded(self.h, world, x.shape[0] // world, _fp(x), _fp(t))

    def train(self, x, t):
        x = np.ascontiguousarray(x, np.float32); t = np.ascontiguousarray(t, np.float32)
        nb_max = x.shape[0] // self.bunchsize + 1
        losses = np.zeros(nb_max, np.float32)
        alphas = np.zeros((nb_max, self.D), np.float32)
        nb = self.L.ggd_oracle_train(self.h, x.shape[0], _fp(x), _fp(t), _fp(losses), _fp(alphas))
        return losses[:nb], alphas[:nb]

    def forward(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros((x.shape[0], self.D), np.float32)
        self.L.ggd_oracle_forward(self.h, x.shape[0], _fp(x), _fp(out))
        return out

    def cv_sqerr(self, x, t):
        x = np.ascontiguousarray(x, np.float32); t = np.ascontiguousarray(t, np.float32)
        return float(self.L.ggd_oracle_cv_sqerr(self.h, x.shape[0], _fp(x), _fp(t)))

    def cv_abserr(self, x, t):
        x = np.ascontiguousarray(x, np.float32); t = np.ascontiguousarray(t, np.float32)
        return float(self.L.ggd_oracle_cv_abserr(self.h, x.shape[0], _fp(x), _fp(t)))

    def cv_loglik(self, x, t):
        x = np.ascontiguousarray(x, np.float32); t = np.ascontiguousarray(t, np.float32)
        return float(self.L.ggd_oracle_cv_loglik(self.h, x.shape[0], _fp(x), _fp(t)))

    def _arr(self, p, n):
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def weights(self):
        n = len(self.layersizes)
        W = [self._arr(self.L.ggd_oracle_W(self.h, l), self.layersizes[l] * self.layersizes[l - 1]) for l in range(1, n)]
        b = [self._arr(self.L.ggd_oracle_b(self.h, l), self.layersizes[l]) for l in range(1, n)]
        return W, b

    def alpha(self):
        return self._arr(self.L.ggd_oracle_alpha(self.h), self.D)

    def out(self, M):
        return self._arr(self.L.ggd_oracle_out(self.h), M * self.D).reshape(M, self.D)

    def dedx(self, l, M):
        return self._arr(self.L.ggd_oracle_dedx(self.h, l), M * self.layersizes[l]).reshape(M, self.layersizes[l])

    def grad(self, l):
        return self._arr(self.L.ggd_oracle_grad(self.h, l), self.layersizes[l] * self.layersizes[l - 1])

    def y(self, l, M):
        return self._arr(self.L.ggd_oracle_y(self.h, l), M * self.layersizes[l]).reshape(M, self.layersizes[l])

    # --- frame-sharded data parallelism, one rank's view (used by the gloo tests)
    def dp_colsum(self, Mg, x, t):
        x = np.ascontiguousarray(x, np.float32); t = np.ascontiguousarray(t, np.float32)
        cs = np.zeros(self.D, np.float32)
        self.L.ggd_oracle_dp_colsum(self.h, x.shape[0], Mg, _fp(x), _fp(t), _fp(cs))
        return cs

    def dp_backward(self, Mg, x, t, colsum_global, colsum_local):
        x = np.ascontiguousarray(x, np.float32); t = np.ascontiguousarray(t, np.float32)
        g = np.ascontiguousarray(colsum_global, np.float32); l = np.ascontiguousarray(colsum_local, np.float32)
        self.L.ggd_oracle_dp_backward(self.h, x.shape[0], Mg, _fp(x), _fp(t), _fp(g), _fp(l))

    def grad_views(self):
        """writable numpy views of the weight / bias gradient buffers (to allreduce in place)"""
        n = len(self.layersizes)
        gw = [np.ctypeslib.as_array(self.L.ggd_oracle_grad(self.h, l), shape=(self.layersizes[l] * self.layersizes[l - 1],)) for l in range(1, n)]
        gb = [np.ctypeslib.as_array(self.L.ggd_oracle_bias_grad(self.h, l), shape=(self.layersizes[l],)) for l in range(1, n)]
        return gw, gb

    def dp_update(self, Mg):
        self.L.ggd_oracle_dp_update(self.h, Mg)

    def last_loss(self):
        return float(self.L.ggd_oracle_last_loss(self.h))


# ---------------------------------------------------------------------------------------------
# LPS
def frame_expand(feature, context_num):
    """Test_code/frame_expand.m:6-25: per frame t the frames t-c..t+c, the first / last frame replicated at the edges"""
    T = feature.shape[0]
    half = (context_num - 1) // 2
    idx = np.clip(np.arange(T)[:, None] + np.arange(-half, half + 1)[None, :], 0, T - 1)
    return feature[idx].reshape(T, -1)


def enhance_ref(lps, W, b, layersizes, mean, dvar, context_num):
    """Test_code/decode.m:28-62 restated in float64 (MATLAB arithmetic): z-score, frame_expand, sigmoid layers, linear
    output, de-normalisation.  W[l] in the .wts order (index = out + in*cur).  (decode.m round-trips the expanded input
    through an 8-digit ASCII file, input_lsp.txt; that quantisation, ~1e-8 relative, is not reproduced.)"""
    x = (np.asarray(lps, np.float64) - mean) * dvar
    y = frame_expand(x, context_num)
    L = len(layersizes)
    for l in range(1, L):
        Wm = np.asarray(W[l - 1], np.float64).reshape(layersizes[l - 1], layersizes[l])     # [in][out]
        y = y @ Wm + np.asarray(b[l - 1], np.float64)
        if l < L - 1:
            y = 1.0 / (1.0 + np.exp(-y))
    return y / dvar + mean


def lps_extract(pcm):
    pcm = np.ascontiguousarray(pcm, np.int16)
    L = lib()
    nf = L.lps_oracle_nframes(len(pcm))
    out = np.zeros((max(nf, 0), 257), np.float32)
    if nf > 0:
        L.lps_oracle_extract(pcm.ctypes.data_as(C.POINTER(C.c_int16)), len(pcm), _fp(out))
    return out


def read_wav_pcm16(path):
    """The bundled wavs are plain 44-byte-header PCM16 mono 16 kHz: strip the header."""
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE"
    return np.frombuffer(raw[44:], dtype="<i2").copy()


def read_htk(path):
    """HTK big-endian feature file: 12-byte header {nSamples, sampPeriod, sampSize, parmKind} (fileio.c:187-211)."""
    raw = open(path, "rb").read()
    n, period = np.frombuffer(raw[:8], ">i4")
    size, kind = np.frombuffer(raw[8:12], ">i2")
    data = np.frombuffer(raw[12:], ">f4").astype(np.float32).reshape(-1, size // 4)
    return dict(nSamples=int(n), sampPeriod=int(period), sampSize=int(size), parmKind=int(kind)), data


def write_htk(path, feats, period=160000, kind=9):
    feats = np.asarray(feats, np.float32)
    with open(path, "wb") as f:
        f.write(np.array([feats.shape[0], period], ">i4").tobytes())
        f.write(np.array([feats.shape[1] * 4, kind], ">i2").tobytes())
        f.write(feats.astype(">f4").tobytes())


# ---------------------------------------------------------------------------------------------
# POSIX drand48 family (Interface.cc:411 srand48, :982 lrand48)
class Rand48:
    A = 0x5DEECE66D
    Cc = 0xB
    MASK = (1 << 48) - 1

    def __init__(self, seed):
        self.x = ((seed & 0xFFFFFFFF) << 16) | 0x330E

    def lrand48(self):
        self.x = (self.A * self.x + self.Cc) & self.MASK
        return self.x >> 17


def rand_index(vec, rng):
    """Interface::GetRandIndex, Interface.cc:975-986 (in place)."""
    n = len(vec)
    for i in range(n - 1):
        idx = rng.lrand48() % (n - i)
        vec[idx], vec[n - 1 - i] = vec[n - 1 - i], vec[idx]
    return vec


# ---------------------------------------------------------------------------------------------
# pfile / norm / wts
PFILE_HEADER = 32768


def read_pfile(path):
    """Returns (feats [F][dim] float32, sent_end [N] cumulative frame counts) (Interface.cc:519-586, 988-1024)."""
    raw = open(path, "rb").read()
    hdr = raw[:PFILE_HEADER].decode("ascii", "ignore")

    def get(name):
        p = hdr.index(name) + len(name)
        return int(hdr[p:].split()[0])
    ns, nf, dim = get("-num_sentences"), get("-num_frames"), get("-num_features")
    rec = np.frombuffer(raw[PFILE_HEADER:PFILE_HEADER + nf * (2 + dim) * 4], ">i4").reshape(nf, 2 + dim)
    feats = rec[:, 2:].copy().view(">f4").astype(np.float32)
    tail = np.frombuffer(raw[PFILE_HEADER + nf * (2 + dim) * 4 + 4:][:ns * 4], ">i4").astype(np.int64)
    return feats, tail, rec[:, 0].astype(np.int64)


def write_pfile(path, feats, sent_lens):
    """Writes a QuickNet pfile with the fields the reference reader uses (header keys, records, index tail)."""
    feats = np.asarray(feats, np.float32)
    nf, dim = feats.shape
    assert sum(sent_lens) == nf
    hdr = ("-pfile_header version 0 size 32768\n-num_sentences %d\n-num_frames %d\n-first_feature_column 2\n"
           "-num_features %d\n-first_label_column %d\n-num_labels 0\n-format dd%s\n-data size %d offset 0 ndim 2 nrow %d ncol %d\n"
           "-sent_table_data size %d offset %d ndim 1\n-end\n") % (
        len(sent_lens), nf, dim, 2 + dim, "f" * dim, nf * (2 + dim), nf, 2 + dim, len(sent_lens) + 1, nf * (2 + dim))
    rec = np.zeros((nf, 2 + dim), ">i4")
    pos = 0
    for s, n in enumerate(sent_lens):
        rec[pos:pos + n, 0] = s
        rec[pos:pos + n, 1] = np.arange(n)
        pos += n
    rec[:, 2:] = feats.astype(">f4").view(">i4")
    tail = np.concatenate([[0], np.cumsum(sent_lens)]).astype(">i4")
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii").ljust(PFILE_HEADER, b"\0"))
        f.write(rec.tobytes())
        f.write(tail.tobytes())


def read_norm(path, dim):
    """Interface.cc:385-396: skip a line, dim means, skip a line, dim reciprocal stds (atof per line)."""
    lines = open(path).read().split("\n")
    mean = np.array([float(x) for x in lines[1:1 + dim]], np.float64).astype(np.float32)
    dvar = np.array([float(x) for x in lines[2 + dim:2 + 2 * dim]], np.float64).astype(np.float32)
    return mean, dvar


def write_norm(path, mean, dvar):
    with open(path, "w") as f:
        f.write("vec %d\n" % len(mean))
        for v in mean:
            f.write("%g\n" % v)
        f.write("vec %d\n" % len(dvar))
        for v in dvar:
            f.write("%g\n" % v)


def write_wts(path, layersizes, W, b):
    """MAT-v4 little-endian, Interface.cc:484-516. W[l-1] flat in `out + in*rows` order."""
    with open(path, "wb") as f:
        for i in range(1, len(layersizes)):
            name = ("weights%d%d" % (i, i + 1)).encode() + b"\0"
            f.write(np.array([10, layersizes[i], layersizes[i - 1], 0, len(name)], "<i4").tobytes())
            f.write(name)
            f.write(np.asarray(W[i - 1], "<f4").tobytes())
            name = ("bias%d" % (i + 1)).encode() + b"\0"
            f.write(np.array([10, 1, layersizes[i], 0, len(name)], "<i4").tobytes())
            f.write(name)
            f.write(np.asarray(b[i - 1], "<f4").tobytes())


def read_wts(path, layersizes):
    """Interface.cc:442-464."""
    raw = open(path, "rb").read()
    pos = 0
    W, b = [], []
    for i in range(1, len(layersizes)):
        st = np.frombuffer(raw[pos:pos + 20], "<i4"); pos += 20 + int(st[4])
        assert st[1] == layersizes[i] and st[2] == layersizes[i - 1], "init weights node nums do not match"
        n = layersizes[i] * layersizes[i - 1]
        W.append(np.frombuffer(raw[pos:pos + 4 * n], "<f4").copy()); pos += 4 * n
        st = np.frombuffer(raw[pos:pos + 20], "<i4"); pos += 20 + int(st[4])
        assert st[2] == layersizes[i] and st[1] == 1, "init bias node nums do not match"
        b.append(np.frombuffer(raw[pos:pos + 4 * layersizes[i]], "<f4").copy()); pos += 4 * layersizes[i]
    return W, b


def init_weights(layersizes, seed=1, beta=2.0):
    """U(+-beta*sqrt(6)/sqrt(n_i+n_j)) weights, zero biases (the distribution of
    pretraining_weights/Gen_rand_net.cpp; generator here is numpy's, not libc rand())."""
    rng = np.random.RandomState(seed)
    W, b = [], []
    for i in range(1, len(layersizes)):
        r = beta * np.sqrt(6.0) / np.sqrt(layersizes[i] + layersizes[i - 1])
        W.append(rng.uniform(-r, r, layersizes[i] * layersizes[i - 1]).astype(np.float32))
        b.append(np.zeros(layersizes[i], np.float32))
    return W, b


class PfileLoader:
    """numpy restatement of Interface::{get_pfile_info,get_chunk_info,Readchunk,Readchunk_cv}
    (Interface.cc:519-972): chunking, z-score with the noisy-speech mean/dVar on BOTH streams
    (:760-766, :804-810), context expansion (:778-785), per-sample lrand48 shuffle (:750-754)."""

    def __init__(self, fea_file, targ_file, norm_file, fea_dim, fea_context, targ_offset, traincache, seed, out_dim=None):
        self.feats, self.sent_end, _ = read_pfile(fea_file)
        self.targs, tend, _ = read_pfile(targ_file)
        assert np.array_equal(self.sent_end, tend)
        self.mean, self.dvar = read_norm(norm_file, fea_dim)
        self.fea_dim, self.ctx, self.off, self.cache = fea_dim, fea_context, targ_offset, traincache
        self.out_dim = out_dim or self.targs.shape[1]
        self.rng = Rand48(seed)

    def chunk_info(self, sent_st, sent_en):
        """get_chunk_info (Interface.cc:588-651). Returns (chunk_frame_st list, total_samples)."""
        cur_frame_id = 0 if sent_st == 0 else int(self.sent_end[sent_st - 1])
        starts = [cur_frame_id]
        cur_chunk_frames = 0
        for s in range(sent_st, sent_en + 1):
            inc = int(self.sent_end[s]) - cur_frame_id
            cur_frame_id = int(self.sent_end[s])
            lost = self.ctx - 1 if inc >= self.ctx else inc
            cur_chunk_frames += inc - lost
            while cur_chunk_frames >= self.cache:
                nxt = cur_frame_id - (cur_chunk_frames - self.cache)
                starts.append(nxt)
                cur_chunk_frames = (cur_frame_id - nxt - self.ctx + 1) if (cur_frame_id - nxt > self.ctx - 1) else 0
        total = (len(starts) - 1) * self.cache + cur_chunk_frames
        return starts, total

    def read_chunk(self, starts, total, sent_en, idx, shuffle=True):
        """Readchunk / Readchunk_cv (Interface.cc:719-838 / 841-958)."""
        last = idx == len(starts) - 1
        if last:
            need = int(self.sent_end[sent_en]) - starts[idx]
            samples = total - self.cache * idx
        else:
            samples = self.cache
            need = starts[idx + 1] - starts[idx]
        order = list(range(samples))
        if shuffle:
            rand_index(order, self.rng)
        order = np.asarray(order)
        f0 = starts[idx]
        x = (self.feats[f0:f0 + need] - self.mean) * self.dvar
        rep = -(-self.out_dim // self.fea_dim)
        t = (self.targs[f0:f0 + need] - np.tile(self.mean, rep)[:self.out_dim]) * np.tile(self.dvar, rep)[:self.out_dim]
        x = x.astype(np.float32); t = t.astype(np.float32)
        ind = np.zeros((samples, self.fea_dim * self.ctx), np.float32)
        tg = np.zeros((samples, self.out_dim), np.float32)
        cur_sent = int(np.searchsorted(self.sent_end, f0, side="right"))
        processed, cur_frame_id, cur_sample = 0, f0, 0
        while processed != need:
            if self.sent_end[cur_sent] > need + f0:
                n = need - processed
            else:
                n = int(self.sent_end[cur_sent]) - cur_frame_id
            for j in range(0, n - self.ctx + 1):
                if cur_sample >= samples:
                    break
                row = order[cur_sample]
                ind[row] = x[processed + j:processed + j + self.ctx].reshape(-1)
                tg[row] = t[processed + j + self.off]
                cur_sample += 1
            cur_frame_id = int(self.sent_end[cur_sent])
            cur_sent += 1
            processed += n
        return ind, tg

    def read_chunk_raw(self, starts, total, sent_en, idx, shuffle=True):
        """The inputs of the device-side loader (ggd_train_raw) for the same chunk: raw big-endian records of the chunk's
        frames and, per (shuffled) net-input row, the first context frame inside the chunk.  Consumes the SAME random
        numbers as read_chunk, so a loader created with the same seed yields the same row order."""
        last = idx == len(starts) - 1
        if last:
            need = int(self.sent_end[sent_en]) - starts[idx]
            samples = total - self.cache * idx
        else:
            samples = self.cache
            need = starts[idx + 1] - starts[idx]
        order = list(range(samples))
        if shuffle:
            rand_index(order, self.rng)
        f0 = starts[idx]
        first = np.zeros(samples, np.int32)
        cur_sent = int(np.searchsorted(self.sent_end, f0, side="right"))
        processed, cur_frame_id, cur_sample = 0, f0, 0
        while processed != need:
            if self.sent_end[cur_sent] > need + f0:
                n = need - processed
            else:
                n = int(self.sent_end[cur_sent]) - cur_frame_id
            for j in range(0, n - self.ctx + 1):
                if cur_sample >= samples:
                    break
                first[order[cur_sample]] = processed + j
                cur_sample += 1
            cur_frame_id = int(self.sent_end[cur_sent])
            cur_sent += 1
            processed += n

        def records(x):
            rec = np.zeros((need, 2 + x.shape[1]), ">u4")
            rec[:, 2:] = x[f0:f0 + need].astype(">f4").view(">u4")
            return rec.view(np.uint32)          # the bytes as they lie in the pfile, read as native words
        return records(self.feats), records(self.targs), first
