"""ctypes drivers for the UNMODIFIED reference code built by oracle/build_ref.sh into oracle/_ref/:

  RefBPGPU      the reference's CUDA device path (BP_GPU.cu + DevFunc.cu + cuBLAS), needs a GPU
  RefInterface  the reference's host loader (Interface.cc), CPU only
  ref_wav2lps   the reference's Wav2LPS_be binary, CPU only

TEST INFRASTRUCTURE ONLY (tests/, bench.py baselines).  Everything here is optional: callers check
`available(...)` first because oracle/_ref/ only exists where /root/reference was present at build time.
"""
import ctypes as C
import os
import subprocess
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
PF = C.POINTER(C.c_float)


def available(name):
    return os.path.exists(os.path.join(REF, name))


def _fp(a):
    return a.ctypes.data_as(PF)


class RefBPGPU:
    def __init__(self, layersizes, bunchsize, lrate, momentum, weightcost, shapefactor, MLflag, W, b, gpu=0, seed=0):
        self.L = C.CDLL(os.path.join(REF, "libref_bpgpu.so"))
        self.L.refbp_create.restype = C.c_void_p
        self.L.refbp_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_float, C.c_float, C.c_float,
                                        C.POINTER(PF), C.POINTER(PF), C.c_float, C.c_int]
        self.L.refbp_train.argtypes = [C.c_void_p, C.c_int, PF, PF]
        self.L.refbp_cv.restype = C.c_float
        self.L.refbp_cv.argtypes = [C.c_void_p, C.c_int, C.c_int, PF, PF]
        self.L.refbp_weights.argtypes = [C.c_void_p, C.POINTER(PF), C.POINTER(PF)]
        self.L.refbp_destroy.argtypes = [C.c_void_p]
        self.ls = list(layersizes)
        n = len(self.ls)
        self._keep = [np.ascontiguousarray(w, np.float32) for w in W] + [np.ascontiguousarray(x, np.float32) for x in b]
        Wp, bp = (PF * 10)(), (PF * 10)()
        for l in range(1, n):
            Wp[l] = _fp(self._keep[l - 1]); bp[l] = _fp(self._keep[n - 1 + l - 1])
        ls = (C.c_int * n)(*self.ls)
        self.h = self.L.refbp_create(seed, gpu, n, ls, bunchsize, lrate, momentum, weightcost, Wp, bp, shapefactor, MLflag)

    def train(self, x, t):
        x = np.ascontiguousarray(x, np.float32); t = np.ascontiguousarray(t, np.float32)
        self.L.refbp_train(self.h, x.shape[0], _fp(x), _fp(t))

    def cv(self, which, x, t):
        x = np.ascontiguousarray(x, np.float32); t = np.ascontiguousarray(t, np.float32)
        return float(self.L.refbp_cv(self.h, which, x.shape[0], _fp(x), _fp(t)))

    def weights(self):
        n = len(self.ls)
        W = [np.zeros(self.ls[l] * self.ls[l - 1], np.float32) for l in range(1, n)]
        b = [np.zeros(self.ls[l], np.float32) for l in range(1, n)]
        Wp, bp = (PF * 10)(), (PF * 10)()
        for l in range(1, n):
            Wp[l] = _fp(W[l - 1]); bp[l] = _fp(b[l - 1])
        self.L.refbp_weights(self.h, Wp, bp)
        return W, b

    def close(self):
        if self.h:
            self.L.refbp_destroy(self.h)
            self.h = None


class RefInterface:
    """Drives the reference's Interface (loader) with finetune.pl-style key=value arguments."""

    def __init__(self, **kw):
        self.L = C.CDLL(os.path.join(REF, "libref_interface.so"))
        L = self.L
        L.refif_create.restype = C.c_void_p
        L.refif_create.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        for f in ("refif_in", "refif_targ"):
            getattr(L, f).restype = PF
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("refif_W", "refif_b"):
            getattr(L, f).restype = PF
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.refif_train_info.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.refif_cv_info.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.refif_shuffle_chunks.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
        L.refif_readchunk.argtypes = [C.c_void_p, C.c_int]
        L.refif_readchunk_cv.argtypes = [C.c_void_p, C.c_int]
        L.refif_writeweights.argtypes = [C.c_void_p]
        args = [b"BPtrain_Sigmoid"] + [("%s=%s" % (k, v)).encode() for k, v in kw.items()]
        self._bufs = [C.create_string_buffer(a) for a in args]     # Initial() writes into argv
        argv = (C.c_char_p * len(args))(*[C.cast(b, C.c_char_p) for b in self._bufs])
        self.h = L.refif_create(len(args), argv)
        self.kw = kw

    def train_info(self, rng):
        a, b = C.c_int(), C.c_int()
        self.L.refif_train_info(self.h, rng.encode(), C.byref(a), C.byref(b))
        return a.value, b.value

    def cv_info(self, rng):
        a, b = C.c_int(), C.c_int()
        self.L.refif_cv_info(self.h, rng.encode(), C.byref(a), C.byref(b))
        return a.value, b.value

    def shuffle_chunks(self, n):
        idx = (C.c_int * n)(*range(n))
        self.L.refif_shuffle_chunks(self.h, idx, n)
        return list(idx)

    def read_chunk(self, idx, in_dim, out_dim, cv=False):
        n = (self.L.refif_readchunk_cv if cv else self.L.refif_readchunk)(self.h, idx)
        x = np.ctypeslib.as_array(self.L.refif_in(self.h), shape=(n * in_dim,)).copy().reshape(n, in_dim)
        t = np.ctypeslib.as_array(self.L.refif_targ(self.h), shape=(n * out_dim,)).copy().reshape(n, out_dim)
        return x, t


def ref_wav2lps(pcm, binary="Wav2LPS_be_ref"):
    """Runs the reference binary on raw int16 PCM; returns (float32 [frames][257], seconds of wall time)."""
    import time
    from . import oracle as O
    with tempfile.TemporaryDirectory() as d:
        raw, out = os.path.join(d, "x.raw"), os.path.join(d, "x.lps")
        np.ascontiguousarray(pcm, np.int16).tofile(raw)
        t0 = time.time()
        subprocess.run([os.path.join(REF, binary), "-F", "RAW", "-fs", "16", raw, out], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.time() - t0
        _, feats = O.read_htk(out)
    return feats, dt
