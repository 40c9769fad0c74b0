"""Python mirror of the reference's Wav2LPS_be front end over the C ABI in include/lps_b200.h."""
import ctypes as C
import numpy as np
from .bp_gpu import load_library

FLAG_BIG_ENDIAN, FLAG_ZSCORE, FLAG_EXACT, FLAG_PFILE, FLAG_ACCUM_NORM = 1, 2, 4, 8, 16
PF = C.POINTER(C.c_float)
PS = C.POINTER(C.c_int16)
PL = C.POINTER(C.c_long)


class LPSError(RuntimeError):
    pass


def _lib():
    L = load_library()
    if not getattr(L, "_lps_ready", False):
        L.lps_last_error.restype = C.c_char_p
        L.lps_nframes.restype = C.c_long
        L.lps_nframes.argtypes = [C.c_long]
        L.lps_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.lps_destroy.argtypes = [C.c_void_p]
        L.lps_set_norm.argtypes = [C.c_void_p, PF, PF]
        L.lps_extract.argtypes = [C.c_void_p, PS, C.c_long, PF, C.c_int]
        L.lps_extract_batch.argtypes = [C.c_void_p, PS, PL, C.c_int, PF, C.c_int, PL]
        L.lps_extract_batch_device.argtypes = [C.c_void_p, C.c_void_p, PL, C.c_int, C.c_void_p, C.c_int, PL]
        L.lps_norm_reset.argtypes = [C.c_void_p]
        L.lps_norm_finalize.argtypes = [C.c_void_p, PF, PF, PL]
        L.lps_norm_accumulate_device.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int]
        L.lps_last_kernel_ms.restype = C.c_double
        L.lps_last_kernel_ms.argtypes = [C.c_void_p]
        L._lps_ready = True
    return L


def lps_nframes(n_samples):
    return int(_lib().lps_nframes(int(n_samples)))


class Wav2LPS:
    """16 kHz LPS extractor: 512-sample frames, 256-sample shift, 257 bins (Wav2LogSpec_be.c:41-59)."""

    def __init__(self, gpu=0):
        self.L = _lib()
        self.h = C.c_void_p()
        self._ck(self.L.lps_create(gpu, C.byref(self.h)))

    def _ck(self, rc):
        if rc != 0:
            raise LPSError("liblps error %d: %s" % (rc, self.L.lps_last_error().decode()))

    def close(self):
        if getattr(self, "h", None):
            self.L.lps_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_norm(self, mean, dvar):
        mean = np.ascontiguousarray(mean, np.float32); dvar = np.ascontiguousarray(dvar, np.float32)
        assert mean.size == 257 and dvar.size == 257
        self._ck(self.L.lps_set_norm(self.h, mean.ctypes.data_as(PF), dvar.ctypes.data_as(PF)))

    def extract(self, pcm, flags=0, out=None):
        pcm = np.ascontiguousarray(pcm, np.int16)
        nf = lps_nframes(pcm.size)
        if out is None:
            out = np.zeros((nf, 257), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size >= nf * 257
        self._ck(self.L.lps_extract(self.h, pcm.ctypes.data_as(PS), pcm.size, out.ctypes.data_as(PF), flags))
        return out

    def extract_batch(self, pcm, utt_off, flags=0):
        pcm = np.ascontiguousarray(pcm, np.int16)
        off = np.ascontiguousarray(utt_off, np.int64)
        total = sum(lps_nframes(int(off[i + 1] - off[i])) for i in range(len(off) - 1))
        out = np.zeros((total, 259 if flags & FLAG_PFILE else 257), np.float32)   # FLAG_PFILE: raw big-endian record words
        n = C.c_long()
        self._ck(self.L.lps_extract_batch(self.h, pcm.ctypes.data_as(PS), off.ctypes.data_as(PL), len(off) - 1,
                                          out.ctypes.data_as(PF), flags, C.byref(n)))
        assert n.value == total
        return out

    def extract_batch_device(self, d_pcm_ptr, utt_off, d_out_ptr, flags=0):
        off = np.ascontiguousarray(utt_off, np.int64)
        n = C.c_long()
        self._ck(self.L.lps_extract_batch_device(self.h, C.c_void_p(d_pcm_ptr), off.ctypes.data_as(PL), len(off) - 1,
                                                 C.c_void_p(d_out_ptr), flags, C.byref(n)))
        return n.value

    def norm_reset(self):
        self._ck(self.L.lps_norm_reset(self.h))

    def norm_of_device_features(self, d_feats_ptr, n_frames, pitch=257, skip=0, big_endian=False):
        """qnnorm over features resident on the device (e.g. the records of an existing pfile: pitch 259, skip 2, big_endian)"""
        self.norm_reset()
        self._ck(self.L.lps_norm_accumulate_device(self.h, C.c_void_p(d_feats_ptr), n_frames, pitch, skip, int(big_endian)))
        mean, dvar, _ = self.norm_finalize()
        return mean, dvar

    def norm_finalize(self):
        """(mean, reciprocal std, frames) over everything extracted with FLAG_ACCUM_NORM since norm_reset (qnnorm)"""
        mean, dvar, n = np.zeros(257, np.float32), np.zeros(257, np.float32), C.c_long()
        self._ck(self.L.lps_norm_finalize(self.h, mean.ctypes.data_as(PF), dvar.ctypes.data_as(PF), C.byref(n)))
        return mean, dvar, n.value

    def last_kernel_ms(self):
        return float(self.L.lps_last_kernel_ms(self.h))
