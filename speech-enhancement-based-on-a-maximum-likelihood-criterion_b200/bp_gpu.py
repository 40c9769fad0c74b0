"""Python mirror of the reference's BP_GPU class (Train_code_ML_GGD/BP_GPU.h:45-70) over the C ABI
of libggd_b200.so (include/ggd_train.h).  Method names, argument meaning and the weight layout
(index = out + in*cur_units) are the reference's."""
import ctypes as C
import os
import numpy as np

PF = C.POINTER(C.c_float)
MAXLAYER = 10
PREC_BF16X3, PREC_FP32_SIMT = 0, 1
FLAG_UNFUSED_UPDATE, FLAG_NO_GRAPH, FLAG_KEEP_DEBUG, FLAG_PIN_HOST = 1, 2, 4, 8
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class GGDError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("numlayers", C.c_int), ("layersizes", C.c_int * MAXLAYER), ("bunchsize", C.c_int),
                ("lrate", C.c_float), ("momentum", C.c_float), ("weightcost", C.c_float), ("shapefactor", C.c_float),
                ("MLflag", C.c_int), ("dropoutflag", C.c_int), ("visible_omit", C.c_float), ("hid_omit", C.c_float),
                ("gpu", C.c_int), ("seed", C.c_int), ("precision", C.c_int), ("world_size", C.c_int), ("rank", C.c_int),
                ("nccl_unique_id", C.c_void_p), ("flags", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("device_ms", C.c_double), ("h2d_ms", C.c_double), ("launches", C.c_longlong), ("steps", C.c_longlong),
                ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong)]


class KernelTimes(C.Structure):
    _fields_ = [("ms", C.c_double * 16), ("launches", C.c_longlong * 16), ("steps", C.c_longlong), ("param_elems", C.c_longlong)]


KERNEL_CLASSES = ("fwd_gemm", "loss", "dx_gemm", "dw_gemm", "bias_grad", "allreduce", "update", "advance", "split", "dw_update", "factor_push")


def library_path():
    return os.path.join(_HERE, "libggd_b200.so")


def load_library():
    """Loads libggd_b200.so; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is None:
        p = library_path()
        if not os.path.exists(p):
            raise GGDError("libggd_b200.so is not built (run %s/build.sh); there is no CPU fallback" % _HERE)
        L = C.CDLL(p, mode=C.RTLD_GLOBAL)
        L.ggd_last_error.restype = C.c_char_p
        L.ggd_version.restype = C.c_char_p
        L.ggd_create.argtypes = [C.POINTER(_Config), C.POINTER(PF), C.POINTER(PF), C.POINTER(C.c_void_p)]
        L.ggd_destroy.argtypes = [C.c_void_p]
        L.ggd_train.argtypes = [C.c_void_p, C.c_int, PF, PF]
        L.ggd_reserve.argtypes = [C.c_void_p, C.c_int]
        L.ggd_release_host.argtypes = [C.c_void_p, C.c_void_p]
        L.ggd_train_raw.argtypes = [C.c_void_p, C.c_void_p]
        L.ggd_train_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        for f in ("ggd_cv_sqerr", "ggd_cv_abserr", "ggd_cv_loglik"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, PF, PF, PF]
        L.ggd_forward.argtypes = [C.c_void_p, C.c_int, PF, PF]
        L.ggd_cv_all.argtypes = [C.c_void_p, C.c_int, PF, PF, PF]
        L.ggd_enhance.argtypes = [C.c_void_p, C.c_int, PF, C.c_int, C.c_int, PF, PF, PF]
        L.ggd_get_weights.argtypes = [C.c_void_p, C.POINTER(PF), C.POINTER(PF)]
        L.ggd_get_alpha.argtypes = [C.c_void_p, PF]
        L.ggd_get_losses.argtypes = [C.c_void_p, PF, C.c_int, C.POINTER(C.c_int)]
        L.ggd_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.ggd_profile_kernels.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(KernelTimes)]
        L.ggd_debug_step.argtypes = [C.c_void_p, C.c_int, PF, PF, C.c_int]
        L.ggd_debug_read.argtypes = [C.c_void_p, C.c_int, C.c_int, PF]
        L.ggd_debug_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, PF, PF, PF]
        L.ggd_nccl_unique_id.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(PF)


def _f32(a):
    a = np.asarray(a)
    if a.dtype != np.float32 or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=np.float32)
    return a


class RawChunk(C.Structure):
    """ggd_raw_chunk (include/ggd_train.h)"""
    _fields_ = [("fea_records", C.POINTER(C.c_uint)), ("targ_records", C.POINTER(C.c_uint)), ("n_frames", C.c_int), ("n_samples", C.c_int),
                ("sample_first_frame", C.POINTER(C.c_int)), ("fea_dim", C.c_int), ("fea_context", C.c_int), ("targ_offset", C.c_int),
                ("mean", C.POINTER(C.c_float)), ("dvar", C.POINTER(C.c_float)), ("rec_frame0", C.c_int), ("rec_frames", C.c_int)]


class BP_GPU:
    """BP_GPU(random_seed, GPU_selected, numlayers, layersizes, bunchsize, lrate, momentum, weightcost,
    weights, bias, shapefactor, MLflag, dropoutflag, visible_omit, hid_omit)  -- BP_GPU.cu:9-113.
    `weights` / `bias` are lists indexed 1..numlayers-1 like the reference's arrays (index 0 unused/None),
    or lists of numlayers-1 arrays."""

    def __init__(self, random_seed, GPU_selected, numlayers, layersizes, bunchsize, lrate, momentum, weightcost,
                 weights, bias, shapefactor, MLflag, dropoutflag=0, visible_omit=0.0, hid_omit=0.0,
                 precision=PREC_BF16X3, flags=0, world_size=1, rank=0, nccl_unique_id=None):
        self.L = load_library()
        self.numlayers = int(numlayers)
        self.layersizes = [int(x) for x in layersizes[:numlayers]]
        self.bunchsize = int(bunchsize)
        self.flags = int(flags)
        self._long_lived = {}
        if len(weights) == numlayers - 1:
            weights = [None] + list(weights)
            bias = [None] + list(bias)
        cfg = _Config()
        cfg.numlayers = numlayers
        for i, v in enumerate(self.layersizes):
            cfg.layersizes[i] = v
        cfg.bunchsize = bunchsize
        cfg.lrate, cfg.momentum, cfg.weightcost, cfg.shapefactor = lrate, momentum, weightcost, shapefactor
        cfg.MLflag, cfg.dropoutflag, cfg.visible_omit, cfg.hid_omit = MLflag, dropoutflag, visible_omit, hid_omit
        cfg.gpu, cfg.seed, cfg.precision, cfg.flags = GPU_selected, random_seed, precision, flags
        cfg.world_size, cfg.rank = world_size, rank
        self._uid = None
        if nccl_unique_id is not None:
            self._uid = C.create_string_buffer(bytes(nccl_unique_id), 128)
            cfg.nccl_unique_id = C.cast(self._uid, C.c_void_p)
        self._w = [None] + [_f32(weights[l]).reshape(-1) for l in range(1, numlayers)]
        self._b = [None] + [_f32(bias[l]).reshape(-1) for l in range(1, numlayers)]
        for l in range(1, numlayers):
            assert self._w[l].size == self.layersizes[l] * self.layersizes[l - 1], "weights[%d] size" % l
            assert self._b[l].size == self.layersizes[l], "bias[%d] size" % l
        Wp, bp = (PF * MAXLAYER)(), (PF * MAXLAYER)()
        for l in range(1, numlayers):
            Wp[l], bp[l] = _fp(self._w[l]), _fp(self._b[l])
        self.h = C.c_void_p()
        self._ck(self.L.ggd_create(C.byref(cfg), Wp, bp, C.byref(self.h)))

    def _ck(self, rc):
        if rc != 0:
            raise GGDError("libggd_b200 error %d: %s" % (rc, self.L.ggd_last_error().decode()))

    def close(self):
        if getattr(self, "h", None):
            self.L.ggd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference surface ---------------------------------------------------------------
    def train(self, n_frames, in_, targ):
        a, t = _f32(in_), _f32(targ)
        self._ck(self.L.ggd_train(self.h, n_frames, _fp(a), _fp(t)))
        if self.flags & FLAG_PIN_HOST:
            # the library registered (pinned) these buffers by address; a temporary made by _f32(), or any array the caller
            # does not keep for the lifetime of the handle, must not stay registered after it is freed
            for given, used in ((in_, a), (targ, t)):
                if used.ctypes.data not in self._long_lived:
                    self._ck(self.L.ggd_release_host(self.h, C.c_void_p(used.ctypes.data)))

    def keep_pinned(self, *arrays):
        """Declares host arrays long-lived (the caller keeps them allocated and unchanged in address until close()):
        with FLAG_PIN_HOST their registration then survives across train() calls, like the reference's chunk buffers."""
        for a in arrays:
            self._long_lived[a.ctypes.data] = a

    def train_raw(self, fea_records, targ_records, sample_first_frame, fea_dim, fea_context, targ_offset, mean, dvar):
        """Device-side loader (ggd_train_raw): raw big-endian pfile records (uint32 [frames][2+dim]) + the first context
        frame of every (shuffled) net-input row; byte swap, z-score, context expansion and target selection run on the GPU
        (the arithmetic of Interface::Readchunk, Interface.cc:735-838)."""
        fr = np.ascontiguousarray(fea_records, dtype=np.uint32)
        tr = np.ascontiguousarray(targ_records, dtype=np.uint32)
        first = np.ascontiguousarray(sample_first_frame, dtype=np.int32)
        mean, dvar = _f32(mean), _f32(dvar)
        c = RawChunk()
        c.fea_records = fr.ctypes.data_as(C.POINTER(C.c_uint)); c.targ_records = tr.ctypes.data_as(C.POINTER(C.c_uint))
        c.n_frames = fr.shape[0]; c.n_samples = first.size
        c.sample_first_frame = first.ctypes.data_as(C.POINTER(C.c_int))
        c.fea_dim, c.fea_context, c.targ_offset = fea_dim, fea_context, targ_offset
        c.mean, c.dvar = _fp(mean), _fp(dvar)
        self._ck(self.L.ggd_train_raw(self.h, C.byref(c)))

    def _cv(self, fn, n_frames, in_, targ):
        in_, targ = _f32(in_), _f32(targ)
        r = C.c_float()
        self._ck(fn(self.h, n_frames, _fp(in_), _fp(targ), C.cast(C.byref(r), PF)))
        return r.value

    def CrossValid(self, n_frames, in_, targ):
        return self._cv(self.L.ggd_cv_sqerr, n_frames, in_, targ)

    def CrossValiddB(self, n_frames, in_, targ):
        return self._cv(self.L.ggd_cv_abserr, n_frames, in_, targ)

    def CrossValid2(self, n_frames, in_, targ):
        return self._cv(self.L.ggd_cv_loglik, n_frames, in_, targ)

    def CrossValidAll(self, n_frames, in_, targ):
        """(CrossValid, CrossValiddB, CrossValid2) from one forward pass"""
        in_, targ = _f32(in_), _f32(targ)
        r = (C.c_float * 3)()
        self._ck(self.L.ggd_cv_all(self.h, n_frames, _fp(in_), _fp(targ), C.cast(r, PF)))
        return r[0], r[1], r[2]

    def enhance(self, lps, mean, dvar, fea_context):
        """Test_code/decode.m for one utterance: raw LPS frames [T][fea_dim] -> enhanced LPS [T][layersizes[-1]]"""
        lps, mean, dvar = _f32(lps), _f32(mean), _f32(dvar)
        out = np.zeros((lps.shape[0], self.layersizes[-1]), np.float32)
        self._ck(self.L.ggd_enhance(self.h, lps.shape[0], _fp(lps), lps.shape[1], fea_context, _fp(mean), _fp(dvar), _fp(out)))
        return out

    def returnWeights(self):
        n = self.numlayers
        W = [None] + [np.zeros(self.layersizes[l] * self.layersizes[l - 1], np.float32) for l in range(1, n)]
        b = [None] + [np.zeros(self.layersizes[l], np.float32) for l in range(1, n)]
        Wp, bp = (PF * MAXLAYER)(), (PF * MAXLAYER)()
        for l in range(1, n):
            Wp[l], bp[l] = _fp(W[l]), _fp(b[l])
        self._ck(self.L.ggd_get_weights(self.h, Wp, bp))
        return W[1:], b[1:]

    def reserve(self, n_frames):
        self._ck(self.L.ggd_reserve(self.h, int(n_frames)))

    # ---- additions (same handle) ---------------------------------------------------------------
    def train_device(self, n_frames, d_in_ptr, d_targ_ptr):
        self._ck(self.L.ggd_train_device(self.h, n_frames, C.c_void_p(d_in_ptr), C.c_void_p(d_targ_ptr)))

    def profile_kernels(self, n_frames, d_in_ptr, d_targ_ptr):
        kt = KernelTimes()
        self._ck(self.L.ggd_profile_kernels(self.h, n_frames, C.c_void_p(d_in_ptr), C.c_void_p(d_targ_ptr), C.byref(kt)))
        out = {"steps": kt.steps, "param_elems": kt.param_elems}
        for i, k in enumerate(KERNEL_CLASSES):
            out[k] = {"ms": kt.ms[i], "launches": kt.launches[i]}
        return out

    def forward(self, in_):
        in_ = _f32(in_)
        out = np.zeros((in_.shape[0], self.layersizes[-1]), np.float32)
        self._ck(self.L.ggd_forward(self.h, in_.shape[0], _fp(in_), _fp(out)))
        return out

    def alpha(self):
        a = np.zeros(self.layersizes[-1], np.float32)
        self._ck(self.L.ggd_get_alpha(self.h, _fp(a)))
        return a

    def losses(self):
        n = C.c_int()
        self._ck(self.L.ggd_get_losses(self.h, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), np.float32)
        self._ck(self.L.ggd_get_losses(self.h, _fp(out), n.value, C.byref(n)))
        return out[:n.value]

    def stats(self):
        s = Stats()
        self._ck(self.L.ggd_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def debug_step(self, in_, targ, apply_update=True):
        in_, targ = _f32(in_), _f32(targ)
        self._ck(self.L.ggd_debug_step(self.h, in_.shape[0], _fp(in_), _fp(targ), int(apply_update)))

    def debug_read(self, what, layer=0):
        ls, M = self.layersizes, self.bunchsize
        shape = {0: (M, ls[-1]), 1: (M, ls[layer]), 2: (M, ls[layer]), 3: (ls[layer - 1] * ls[layer],), 4: (ls[layer],)}[what]
        out = np.zeros(shape, np.float32)
        self._ck(self.L.ggd_debug_read(self.h, what, layer, _fp(out)))
        return out


def nccl_unique_id():
    """128-byte ncclUniqueId to be broadcast to all ranks (rank 0 creates it)."""
    L = load_library()
    buf = C.create_string_buffer(128)
    rc = L.ggd_nccl_unique_id(buf)
    if rc != 0:
        raise GGDError("ggd_nccl_unique_id failed: %s" % L.ggd_last_error().decode())
    return buf.raw


def debug_gemm(a_mn, b_mn, I, J, R, bn, splits, A, B):
    """Runs the raw tcgen05 GEMM D[i][j] = sum_r A(i,r) B(j,r) with a plain fp32 store epilogue (tests only).
    A is [I][R] (K-major) or [R][I] (MN-major); B is [J][R] or [R][J]."""
    L = load_library()
    A, B = _f32(A), _f32(B)
    Cm = np.zeros((I, J), np.float32)
    rc = L.ggd_debug_gemm(a_mn, b_mn, I, J, R, bn, splits, _fp(A), _fp(B), _fp(Cm))
    if rc != 0:
        raise GGDError("ggd_debug_gemm error %d: %s" % (rc, L.ggd_last_error().decode()))
    return Cm
