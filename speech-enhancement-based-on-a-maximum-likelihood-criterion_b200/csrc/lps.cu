// lps.cu -- B200 LPS front end (include/lps_b200.h).  Two kernels:
//
// lps_fast_kernel (default): register-resident FFT.  A 512-point real transform is a 256-point complex transform of
// z[n] = x[2n] + i x[2n+1] plus a split step; one warp = one frame, 8 complex points per lane, radix 8 x 8 x 4:
// radix-8 butterflies in registers, one shared-memory transpose, radix-8 again, the last radix-4 across lane quads
// with shuffles, a second transpose for the conjugate-pair split, hardware log2.  fp32 with exactly rounded twiddles;
// ~450 warp instructions per frame against ~6 000 for the schedule interpreter below.
//
// lps_kernel (LPS_FLAG_EXACT): the reference's butterfly network itself, bit-identical spectrum.
// One warp per frame.  The frame's 512 windowed samples live in shared memory; the warp executes the
// reference's in-place split-radix real FFT (FEfunc.c:146-293) as a precomputed butterfly SCHEDULE:
// the host walks the reference's loop nest once and records, stage by stage, every butterfly with its
// indices and twiddles.  Butterflies of one stage touch disjoint elements, so the 32 lanes execute
// them in parallel and only a __syncwarp() separates stages.  Every butterfly performs the reference's
// float operations in the reference's order with explicit round-to-nearest intrinsics (no FMA
// contraction), which makes the spectrum bit-identical to the reference; the floored natural log is
// taken in double like the reference ((float)log((double)P), Wav2LogSpec_be.c:475-479).
//
// Loads: frame n covers samples [256n, 256n+512) (coalesced 64-byte warp loads, each sample is read by
// the two frames that overlap it and hits L1/L2 the second time).  Stores: 257 consecutive floats per
// frame, optionally byte-swapped for the HTK writer or z-scored for the trainer.
#include "../../include/lps_b200.h"
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <vector>

namespace {

thread_local char g_lps_err[512] = "";
void lps_err(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_lps_err, sizeof g_lps_err, fmt, ap);
    va_end(ap);
}
#define LPS_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            lps_err("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
            return -2;                                                                      \
        }                                                                                   \
    } while (0)

constexpr int N = LPS_FRAME_LEN, LOGN = 9;
constexpr int WARPS = 8;
constexpr int MAX_OPS = 704, MAX_TW = 128, NSTAGE = 9;

// butterfly kinds
enum { OP_LEN2 = 0, OP_L0 = 1, OP_L8 = 2, OP_LTW = 3 };
// op word: [0,2) kind  [2,12) base index i  [12,20) j  [20,28) twiddle slot
static inline uint32_t mk_op(int kind, int i, int j, int tw) { return (uint32_t)kind | ((uint32_t)i << 2) | ((uint32_t)j << 12) | ((uint32_t)tw << 20); }

struct Schedule {
    uint32_t ops[MAX_OPS];
    float4 tw[MAX_TW];       // cc1, ss1, cc3, ss3
    float win[N / 2];
    int stage_beg[NSTAGE + 1];
    int stage_n4[NSTAGE];
    float floor_fb;          // (float)exp(-50.0)
};

// Walks the loop nest of the reference rfft(x, 512, 9) and records the butterflies (FEfunc.c:183-292).
void build_schedule(Schedule &S)
{
    int nops = 0, ntw = 0, st = 0;
    S.stage_beg[0] = 0; S.stage_n4[0] = 0;
    for (int is = 0, id = 4; is < N - 1; is = 2 * id - 2, id *= 4)          // length-two butterflies, :184-199
        for (int i0 = is; i0 < N; i0 += id) S.ops[nops++] = mk_op(OP_LEN2, i0, 0, 0);
    S.stage_beg[++st] = nops;
    int n2 = 2;
    for (int k = 1; k < LOGN; k++) {                                         // L-shaped butterflies, :202-292
        n2 <<= 1;
        const int n4 = n2 >> 2, n8 = n2 >> 3;
        const float e = (float)((3.14159265358979323846 * 2) / n2);
        S.stage_n4[st] = n4;
        // ops are grouped by kind inside a stage so that neighbouring lanes run the same code
        for (int is = 0, id = n2 << 1; is < N; is = 2 * id - n2, id *= 4)
            for (int i = is; i <= N - 1; i += id) S.ops[nops++] = mk_op(OP_L0, i, 0, 0);
        if (n4 != 1)
            for (int is = 0, id = n2 << 1; is < N; is = 2 * id - n2, id *= 4)
                for (int i = is; i <= N - 1; i += id) S.ops[nops++] = mk_op(OP_L8, i + n8, 0, 0);
        for (int j = 1; j < n8; j++) {
            const float a = j * e, a3 = 3 * a;
            // the reference is C: cos(float) promotes to double (in C++ the float overload would be picked)
            S.tw[ntw] = make_float4((float)cos((double)a), (float)sin((double)a), (float)cos((double)a3), (float)sin((double)a3));
            for (int is = 0, id = n2 << 1; is < N; is = 2 * id - n2, id *= 4)
                for (int i = is; i <= N - 1; i += id) S.ops[nops++] = mk_op(OP_LTW, i, j, ntw);
            ntw++;
        }
        S.stage_beg[++st] = nops;
    }
    for (int i = 0; i < N / 2; i++) S.win[i] = (float)(0.54 - 0.46 * cos(6.28318530717958647692 * i / (N - 1)));   // FEfunc.c:80-87
    S.floor_fb = (float)exp((double)-50.0);
    if (nops > MAX_OPS || ntw > MAX_TW || st != NSTAGE) { fprintf(stderr, "lps schedule overflow %d %d %d\n", nops, ntw, st); abort(); }
}

struct LpsArgs {
    const int16_t *pcm;
    const long long *utt_sample_off;   // [n_utts + 1]
    const long long *utt_frame_off;    // [n_utts + 1]
    int n_utts;
    long long frame_begin, total_frames;   // frames [frame_begin, total_frames) of the batch are processed
    float *out;
    const float *mean, *dvar;
    int flags;
    const Schedule *sched;
};

__global__ void __launch_bounds__(WARPS * 32) lps_kernel(const LpsArgs a)
{
    __shared__ Schedule S;
    __shared__ float xs[WARPS][N];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.sched);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&S);
        for (int i = threadIdx.x; i < (int)(sizeof(Schedule) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *x = xs[warp];
    const double sqrt2_d = 1.41421356237309504880;

    for (long long f = a.frame_begin + (long long)blockIdx.x * WARPS + warp; f < a.total_frames; f += (long long)gridDim.x * WARPS) {
        // utterance of this frame: largest u with frame_off[u] <= f
        int lo = 0, hi = a.n_utts;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (a.utt_frame_off[mid] <= f) lo = mid; else hi = mid;
        }
        const long long base = a.utt_sample_off[lo] + (f - a.utt_frame_off[lo]) * LPS_FRAME_SHIFT;
        const int16_t *src = a.pcm + base;
        // window and bit-reversed scatter (ReadWave fileio.c:268-282, Window FEfunc.c:106-118, rfft :157-181)
#pragma unroll
        for (int e = 0; e < N / 32; e++) {
            const int p = e * 32 + lane;
            const float w = S.win[p < N / 2 ? p : N - 1 - p];
            x[__brev((unsigned)p) >> (32 - LOGN)] = __fmul_rn((float)__ldg(src + p), w);
        }
        __syncwarp();
        // stage 0: length-two butterflies
        for (int o = S.stage_beg[0] + lane; o < S.stage_beg[1]; o += 32) {
            const int i0 = (S.ops[o] >> 2) & 1023;
            const float a0 = x[i0], a1 = x[i0 + 1];
            x[i0] = __fadd_rn(a0, a1);
            x[i0 + 1] = __fsub_rn(a0, a1);
        }
        __syncwarp();
        for (int st = 1; st < NSTAGE; st++) {
            const int n4 = S.stage_n4[st];
            for (int o = S.stage_beg[st] + lane; o < S.stage_beg[st + 1]; o += 32) {
                const uint32_t op = S.ops[o];
                const int kind = op & 3, i = (op >> 2) & 1023;
                if (kind == OP_L0) {                       // FEfunc.c:217-224
                    const int i1 = i, i3 = i + 2 * n4, i4 = i3 + n4;
                    const float x1 = x[i1], x3 = x[i3], x4 = x[i4];
                    const float t1 = __fadd_rn(x4, x3);
                    x[i4] = __fsub_rn(x4, x3);
                    x[i3] = __fsub_rn(x1, t1);
                    x[i1] = __fadd_rn(x1, t1);
                } else if (kind == OP_L8) {                // FEfunc.c:226-238 (division by sqrt(2) in double)
                    const int i1 = i, i2 = i1 + n4, i3 = i2 + n4, i4 = i3 + n4;
                    const float x1 = x[i1], x2 = x[i2], x3 = x[i3], x4 = x[i4];
                    const float t1 = __double2float_rn(__ddiv_rn((double)__fadd_rn(x3, x4), sqrt2_d));
                    const float t2 = __double2float_rn(__ddiv_rn((double)__fsub_rn(x3, x4), sqrt2_d));
                    x[i4] = __fsub_rn(x2, t1);
                    x[i3] = __fsub_rn(-x2, t1);
                    x[i2] = __fsub_rn(x1, t2);
                    x[i1] = __fadd_rn(x1, t2);
                } else {                                   // FEfunc.c:257-287
                    const int j = (op >> 12) & 255;
                    const float4 tw = S.tw[(op >> 20) & 255];
                    const float cc1 = tw.x, ss1 = tw.y, cc3 = tw.z, ss3 = tw.w;
                    const int i1 = i + j, i2 = i1 + n4, i3 = i2 + n4, i4 = i3 + n4;
                    const int i5 = i + n4 - j, i6 = i5 + n4, i7 = i6 + n4, i8 = i7 + n4;
                    const float x1 = x[i1], x2 = x[i2], x3 = x[i3], x4 = x[i4], x5 = x[i5], x6 = x[i6], x7 = x[i7], x8 = x[i8];
                    float t1 = __fadd_rn(__fmul_rn(x3, cc1), __fmul_rn(x7, ss1));
                    float t2 = __fsub_rn(__fmul_rn(x7, cc1), __fmul_rn(x3, ss1));
                    float t3 = __fadd_rn(__fmul_rn(x4, cc3), __fmul_rn(x8, ss3));
                    float t4 = __fsub_rn(__fmul_rn(x8, cc3), __fmul_rn(x4, ss3));
                    const float t5 = __fadd_rn(t1, t3), t6 = __fadd_rn(t2, t4);
                    t3 = __fsub_rn(t1, t3);
                    t4 = __fsub_rn(t2, t4);
                    x[i8] = __fadd_rn(x6, t6);
                    x[i3] = __fsub_rn(t6, x6);
                    x[i4] = __fsub_rn(x2, t3);
                    x[i7] = __fsub_rn(-x2, t3);
                    x[i1] = __fadd_rn(x1, t5);
                    x[i6] = __fsub_rn(x1, t5);
                    x[i2] = __fadd_rn(x5, t4);
                    x[i5] = __fsub_rn(x5, t4);
                }
            }
            __syncwarp();
        }
        // power spectrum + floored natural log (Wav2LogSpec_be.c:469-479); output order Re(0..256), Im(255..1)
        const bool pfile = (a.flags & LPS_FLAG_PFILE) != 0;
        float *dst = a.out + f * (pfile ? LPS_BINS + 2 : LPS_BINS) + (pfile ? 2 : 0);
        if (pfile && lane == 0) {     // pfile record = {sentence, frame in sentence, 257 floats}, big-endian words (Interface.cc:735-766)
            dst[-2] = __uint_as_float(__byte_perm((unsigned)lo, 0, 0x0123));
            dst[-1] = __uint_as_float(__byte_perm((unsigned)(f - a.utt_frame_off[lo]), 0, 0x0123));
        }
        for (int k = lane; k <= N / 2; k += 32) {
            const float re = x[k];
            float p = __fmul_rn(re, re);
            if (k != 0 && k != N / 2) {
                const float im = x[N - k];
                p = __fadd_rn(p, __fmul_rn(im, im));
            }
            float v = (p < S.floor_fb) ? -50.0f : __double2float_rn(log((double)p));
            if (a.flags & LPS_FLAG_ZSCORE) v = __fmul_rn(__fsub_rn(v, a.mean[k]), a.dvar[k]);   // Interface.cc:763-764
            if (a.flags & (LPS_FLAG_BIG_ENDIAN | LPS_FLAG_PFILE)) v = __uint_as_float(__byte_perm(__float_as_uint(v), 0, 0x0123));
            dst[k] = v;
        }
        __syncwarp();
    }
}


// =====================================================================================================================
// Fast path
// =====================================================================================================================
constexpr int XSTR = 36;                 // row pitch (float2) of the first transpose: conflict-free per half-warp
struct FastTab {
    float win[N];                        // Hamming window (FEfunc.c:80-87), in double then float like the reference
    float2 tw1[8][32];                   // w256^(b*k1)   [k1][b]
    float2 tw2[8][4];                    // w32^(d*p)     [p][d]
    float2 post[LPS_BINS];               // (cos, sin)(2 pi k / 512)
    float floor_fb;                      // (float)exp(-50.0)
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// a * (c - i s)  with tw = (c, s): forward-transform twiddle e^{-i theta}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 tw)
{
    return make_float2(fmaf(a.x, tw.x, a.y * tw.y), fmaf(a.y, tw.x, -a.x * tw.y));
}
// 8-point forward DFT in registers, natural order in and out
__device__ __forceinline__ void fft8(float2 *v)
{
    const float R = 0.70710678118654752440f;
    float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
    // b_j *= w8^j:  w8 = (1 - i)/sqrt2, w8^2 = -i, w8^3 = (-1 - i)/sqrt2
    b1 = make_float2((b1.x + b1.y) * R, (b1.y - b1.x) * R);
    b2 = make_float2(b2.y, -b2.x);
    b3 = make_float2((b3.y - b3.x) * R, -(b3.x + b3.y) * R);
    // two 4-point DFTs: even outputs from a, odd outputs from b
    {
        const float2 s0 = cadd(a0, a2), s1 = csub(a0, a2), s2 = cadd(a1, a3), t = csub(a1, a3);
        const float2 s3 = make_float2(t.y, -t.x);
        v[0] = cadd(s0, s2); v[4] = csub(s0, s2); v[2] = cadd(s1, s3); v[6] = csub(s1, s3);
    }
    {
        const float2 s0 = cadd(b0, b2), s1 = csub(b0, b2), s2 = cadd(b1, b3), t = csub(b1, b3);
        const float2 s3 = make_float2(t.y, -t.x);
        v[1] = cadd(s0, s2); v[5] = csub(s0, s2); v[3] = cadd(s1, s3); v[7] = csub(s1, s3);
    }
}

struct FastArgs {
    const int16_t *pcm;
    const long long *utt_sample_off, *utt_frame_off;
    int n_utts;
    long long frame_begin, total_frames;
    float *out;
    const float *mean, *dvar;
    int flags;
    const FastTab *tab;
};

constexpr int FWARPS = 8;
__global__ void __launch_bounds__(FWARPS * 32, 3) lps_fast_kernel(const FastArgs a)
{
    __shared__ float2 xch[FWARPS][8 * XSTR];          // transpose 1: [k1][b]; transpose 2: Z[k] at k + 4 (k >> 6)
    __shared__ float2 s_post[LPS_BINS];
    __shared__ float2 s_tw2[8][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < LPS_BINS; i += blockDim.x) s_post[i] = a.tab->post[i];
    if (threadIdx.x < 32) s_tw2[threadIdx.x >> 2][threadIdx.x & 3] = a.tab->tw2[threadIdx.x >> 2][threadIdx.x & 3];
    // per-lane constants: window at samples 64 a + 2 lane (+1), first-pass twiddles w256^(lane*k1)
    float2 win[8], tw1[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        win[j] = make_float2(a.tab->win[64 * j + 2 * lane], a.tab->win[64 * j + 2 * lane + 1]);
        tw1[j] = a.tab->tw1[j][lane];
    }
    const float floor_fb = a.tab->floor_fb;
    __syncthreads();
    float2 *xs = xch[warp];
    const int k1 = lane >> 2, d = lane & 3;              // second-pass role of this lane
    const int q = ((d & 1) << 1) | (d >> 1);             // output index of the cross-lane radix-4 (bit-reversed d)

    for (long long f = a.frame_begin + (long long)blockIdx.x * FWARPS + warp; f < a.total_frames; f += (long long)gridDim.x * FWARPS) {
        int lo = 0, hi = a.n_utts;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (a.utt_frame_off[mid] <= f) lo = mid; else hi = mid;
        }
        const long long base = a.utt_sample_off[lo] + (f - a.utt_frame_off[lo]) * LPS_FRAME_SHIFT;
        const int16_t *src = a.pcm + base + 2 * lane;
        // z[32 j + lane] = x[64 j + 2 lane] + i x[64 j + 2 lane + 1], windowed (ReadWave fileio.c:268-282, Window FEfunc.c:106-118)
        float2 v[8];
        if ((base & 1) == 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const unsigned int w = __ldg(reinterpret_cast<const unsigned int *>(src + 64 * j));
                v[j] = make_float2((float)(short)(w & 0xFFFFu) * win[j].x, (float)(short)(w >> 16) * win[j].y);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = make_float2((float)__ldg(src + 64 * j) * win[j].x, (float)__ldg(src + 64 * j + 1) * win[j].y);
        }
        // pass 1: DFT over j (stride-32 points), twiddle w256^(lane*k1), transpose through shared memory
        fft8(v);
#pragma unroll
        for (int j = 1; j < 8; j++) v[j] = cmul_conj(v[j], tw1[j]);
#pragma unroll
        for (int j = 0; j < 8; j++) xs[j * XSTR + lane] = v[j];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; c++) v[c] = xs[k1 * XSTR + 4 * c + d];
        __syncwarp();
        // pass 2: DFT over c, twiddle w32^(d*p), then the 4-point DFT over d across the lane quad
        fft8(v);
#pragma unroll
        for (int p = 1; p < 8; p++) v[p] = cmul_conj(v[p], s_tw2[p][d]);
#pragma unroll
        for (int p = 0; p < 8; p++) {
            float2 u = v[p];
            float2 t = make_float2(__shfl_xor_sync(0xffffffffu, u.x, 2), __shfl_xor_sync(0xffffffffu, u.y, 2));
            u = (d & 2) ? csub(t, u) : cadd(u, t);
            if (d == 3) u = make_float2(u.y, -u.x);      // * (-i)
            t = make_float2(__shfl_xor_sync(0xffffffffu, u.x, 1), __shfl_xor_sync(0xffffffffu, u.y, 1));
            u = (d & 1) ? csub(t, u) : cadd(u, t);
            // Z[k], k = k1 + 8 p + 64 q
            const int k = k1 + 8 * p + 64 * q;
            xs[k + 4 * (k >> 6)] = u;
        }
        __syncwarp();
        // split: X[k] = (Zk + conj Zm)/2 + e^{-2 pi i k/512} (Zk - conj Zm)/(2i), m = 256 - k; P = |X|^2; floored ln
        // (Wav2LogSpec_be.c:469-479)
        const bool pfile = (a.flags & LPS_FLAG_PFILE) != 0;
        float *dst = a.out + f * (pfile ? LPS_BINS + 2 : LPS_BINS) + (pfile ? 2 : 0);
        if (pfile && lane == 0) {     // pfile record = {sentence, frame in sentence, 257 floats}, big-endian words (Interface.cc:735-766)
            dst[-2] = __uint_as_float(__byte_perm((unsigned)lo, 0, 0x0123));
            dst[-1] = __uint_as_float(__byte_perm((unsigned)(f - a.utt_frame_off[lo]), 0, 0x0123));
        }
#pragma unroll
        for (int j = 0; j < 9; j++) {
            const int k = lane + 32 * j;
            if (k <= N / 2) {
                const int kk = k & 255, mm = (256 - k) & 255;
                const float2 zk = xs[kk + 4 * (kk >> 6)], zm = xs[mm + 4 * (mm >> 6)];
                const float2 A = make_float2(zk.x + zm.x, zk.y - zm.y);       // Zk + conj(Zm)
                const float2 B = make_float2(zk.x - zm.x, zk.y + zm.y);       // Zk - conj(Zm)
                const float2 w = s_post[k];                                    // (cos, sin)
                // -i B = (B.y, -B.x); times (c - i s)
                const float xr = A.x + fmaf(B.y, w.x, -B.x * w.y);
                const float xi = A.y - fmaf(B.x, w.x, B.y * w.y);
                const float pw = 0.25f * fmaf(xr, xr, xi * xi);
                float val = (pw < floor_fb) ? -50.0f : __logf(pw);
                if (a.flags & LPS_FLAG_ZSCORE) val = __fmul_rn(__fsub_rn(val, a.mean[k]), a.dvar[k]);   // Interface.cc:763-764
                if (a.flags & (LPS_FLAG_BIG_ENDIAN | LPS_FLAG_PFILE)) val = __uint_as_float(__byte_perm(__float_as_uint(val), 0, 0x0123));
                dst[k] = val;
            }
        }
        __syncwarp();
    }
}

// Per-bin sum and sum of squares over frames [f0, f1) in double (qnnorm: mean and reciprocal standard deviation, ddof 0;
// tools_pfile/get_norm.pl:4).  Rows are `pitch` words apart, the 257 values start `skip` words into a row and are big-endian
// when `swap` (pfile records).  One thread per bin, frames strided over the grid, one double atomic per thread at the end.
__global__ void __launch_bounds__(288) lps_norm_kernel(const float *feats, long long f0, long long f1, int pitch, int skip, int swap, double *acc)
{
    const int k = threadIdx.x;
    if (k >= LPS_BINS) return;
    double s = 0.0, ss = 0.0;
    for (long long f = f0 + blockIdx.x; f < f1; f += gridDim.x) {
        unsigned int w = __float_as_uint(feats[f * pitch + skip + k]);
        if (swap) w = __byte_perm(w, 0, 0x0123);
        const double v = (double)__uint_as_float(w);
        s += v; ss += v * v;
    }
    atomicAdd(acc + k, s);
    atomicAdd(acc + LPS_BINS + k, ss);
}

void build_fast_tab(FastTab &T)
{
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < N; i++) {
        const int p = i < N / 2 ? i : N - 1 - i;          // the reference applies win[i] to both ends (FEfunc.c:106-118)
        T.win[i] = (float)(0.54 - 0.46 * cos(6.28318530717958647692 * p / (N - 1)));
    }
    for (int k = 0; k < 8; k++)
        for (int b = 0; b < 32; b++) T.tw1[k][b] = make_float2((float)cos(2 * PI * b * k / 256), (float)sin(2 * PI * b * k / 256));
    for (int p = 0; p < 8; p++)
        for (int d = 0; d < 4; d++) T.tw2[p][d] = make_float2((float)cos(2 * PI * d * p / 32), (float)sin(2 * PI * d * p / 32));
    for (int k = 0; k < LPS_BINS; k++) T.post[k] = make_float2((float)cos(2 * PI * k / 512), (float)sin(2 * PI * k / 512));
    T.floor_fb = (float)exp((double)-50.0);
}

}  // namespace

struct lps_handle {
    int gpu, sm_count;
    Schedule *d_sched;
    FastTab *d_tab;
    double *d_norm_acc; long long norm_frames;   // running per-bin sum / sum of squares (LPS_FLAG_ACCUM_NORM)
    float *d_mean, *d_dvar;
    bool has_norm;
    cudaStream_t s, s2;                 // s2: second lane of the host-batch pipeline
    cudaEvent_t e0, e1;
    std::vector<cudaEvent_t> pe;        // per-piece kernel event pairs of the host-batch pipeline
    // staging (grown on demand)
    int16_t *d_pcm; size_t pcm_cap;
    float *d_out; size_t out_cap;
    long long *d_off; size_t off_cap;
    double last_ms;
};

extern "C" {

const char *lps_last_error(void) { return g_lps_err; }

long lps_nframes(long n_samples)
{
    if (n_samples < LPS_FRAME_LEN) return 0;
    return (n_samples - (LPS_FRAME_LEN - LPS_FRAME_SHIFT)) / LPS_FRAME_SHIFT;
}

int lps_create(int gpu, lps_handle **out)
{
    if (!out) { lps_err("lps_create: null argument"); return -1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); lps_err("no CUDA device: liblps has no CPU fallback"); return -2; }
    if (gpu < 0 || gpu >= ndev) { lps_err("gpu %d out of range 0-%d", gpu, ndev - 1); return -1; }
    LPS_CUDA(cudaSetDevice(gpu));
    cudaDeviceProp prop;
    LPS_CUDA(cudaGetDeviceProperties(&prop, gpu));
    if (prop.major != 10) { lps_err("device %d is sm_%d%d; built for sm_100a only", gpu, prop.major, prop.minor); return -2; }
    lps_handle *h = new lps_handle();
    h->gpu = gpu; h->sm_count = prop.multiProcessorCount;
    Schedule *S = new Schedule();
    memset(S, 0, sizeof *S);
    build_schedule(*S);
    LPS_CUDA(cudaMalloc(&h->d_sched, sizeof(Schedule)));
    LPS_CUDA(cudaMemcpy(h->d_sched, S, sizeof(Schedule), cudaMemcpyHostToDevice));
    delete S;
    FastTab *T = new FastTab();
    build_fast_tab(*T);
    LPS_CUDA(cudaMalloc(&h->d_tab, sizeof(FastTab)));
    LPS_CUDA(cudaMemcpy(h->d_tab, T, sizeof(FastTab), cudaMemcpyHostToDevice));
    delete T;
    LPS_CUDA(cudaMalloc(&h->d_norm_acc, 2 * LPS_BINS * sizeof(double)));
    LPS_CUDA(cudaMemset(h->d_norm_acc, 0, 2 * LPS_BINS * sizeof(double)));
    LPS_CUDA(cudaMalloc(&h->d_mean, LPS_BINS * sizeof(float)));
    LPS_CUDA(cudaMalloc(&h->d_dvar, LPS_BINS * sizeof(float)));
    LPS_CUDA(cudaStreamCreateWithFlags(&h->s, cudaStreamNonBlocking));
    LPS_CUDA(cudaStreamCreateWithFlags(&h->s2, cudaStreamNonBlocking));
    LPS_CUDA(cudaEventCreate(&h->e0));
    LPS_CUDA(cudaEventCreate(&h->e1));
    *out = h;
    return 0;
}

int lps_destroy(lps_handle *h)
{
    if (!h) return 0;
    cudaSetDevice(h->gpu);
    for (cudaEvent_t e : h->pe) cudaEventDestroy(e);
    if (h->s2) cudaStreamDestroy(h->s2);
    cudaFree(h->d_tab); cudaFree(h->d_norm_acc);
    cudaFree(h->d_sched); cudaFree(h->d_mean); cudaFree(h->d_dvar); cudaFree(h->d_pcm); cudaFree(h->d_out); cudaFree(h->d_off);
    cudaEventDestroy(h->e0); cudaEventDestroy(h->e1); cudaStreamDestroy(h->s);
    delete h;
    return 0;
}

int lps_set_norm(lps_handle *h, const float *mean, const float *dvar)
{
    if (!h || !mean || !dvar) { lps_err("lps_set_norm: null argument"); return -1; }
    LPS_CUDA(cudaSetDevice(h->gpu));
    LPS_CUDA(cudaMemcpy(h->d_mean, mean, LPS_BINS * sizeof(float), cudaMemcpyHostToDevice));
    LPS_CUDA(cudaMemcpy(h->d_dvar, dvar, LPS_BINS * sizeof(float), cudaMemcpyHostToDevice));
    h->has_norm = true;
    return 0;
}

static int offsets(lps_handle *h, const long *utt_off, int n_utts, std::vector<long long> &tab, long long *total)
{
    tab.resize(2 * (size_t)(n_utts + 1));
    long long fr = 0;
    for (int u = 0; u <= n_utts; u++) {
        tab[u] = utt_off[u];
        tab[n_utts + 1 + u] = fr;
        if (u < n_utts) {
            if (utt_off[u + 1] < utt_off[u]) { lps_err("utterance offsets must be non-decreasing"); return -1; }
            fr += lps_nframes(utt_off[u + 1] - utt_off[u]);
        }
    }
    *total = fr;
    if (tab.size() > h->off_cap) {
        cudaFree(h->d_off);
        LPS_CUDA(cudaMalloc(&h->d_off, tab.size() * sizeof(long long)));
        h->off_cap = tab.size();
    }
    LPS_CUDA(cudaMemcpyAsync(h->d_off, tab.data(), tab.size() * sizeof(long long), cudaMemcpyHostToDevice, h->s));
    return 0;
}

static int check_flags(lps_handle *h, int flags)
{
    if ((flags & LPS_FLAG_ZSCORE) && !h->has_norm) { lps_err("LPS_FLAG_ZSCORE needs lps_set_norm first"); return -1; }
    if ((flags & LPS_FLAG_ZSCORE) && (flags & (LPS_FLAG_BIG_ENDIAN | LPS_FLAG_PFILE))) { lps_err("ZSCORE cannot be combined with BIG_ENDIAN / PFILE"); return -1; }
    return 0;
}

// frames [f0, f1) of the batch on stream `st`
static int launch_frames(lps_handle *h, const int16_t *d_pcm, int n_utts, long long f0, long long f1, float *d_out, int flags, cudaStream_t st)
{
    if (f1 <= f0) return 0;
    const long long n = f1 - f0;
    if (flags & LPS_FLAG_EXACT) {
        LpsArgs a;
        a.pcm = d_pcm; a.utt_sample_off = h->d_off; a.utt_frame_off = h->d_off + n_utts + 1; a.n_utts = n_utts;
        a.frame_begin = f0; a.total_frames = f1; a.out = d_out; a.mean = h->d_mean; a.dvar = h->d_dvar; a.flags = flags; a.sched = h->d_sched;
        long long blocks = (n + WARPS - 1) / WARPS;
        const long long cap = (long long)h->sm_count * 6;   // persistent-style grid: 6 resident CTAs per SM
        if (blocks > cap) blocks = cap;
        lps_kernel<<<(int)blocks, WARPS * 32, 0, st>>>(a);
    } else {
        FastArgs a;
        a.pcm = d_pcm; a.utt_sample_off = h->d_off; a.utt_frame_off = h->d_off + n_utts + 1; a.n_utts = n_utts;
        a.frame_begin = f0; a.total_frames = f1; a.out = d_out; a.mean = h->d_mean; a.dvar = h->d_dvar; a.flags = flags; a.tab = h->d_tab;
        long long blocks = (n + FWARPS - 1) / FWARPS;
        const long long cap = (long long)h->sm_count * 12;  // 3 resident CTAs per SM (80 registers), 4 rounds for balance
        if (blocks > cap) blocks = cap;
        lps_fast_kernel<<<(int)blocks, FWARPS * 32, 0, st>>>(a);
    }
    if (flags & LPS_FLAG_ACCUM_NORM) {
        const bool pf = (flags & LPS_FLAG_PFILE) != 0;
        const int swap = (flags & (LPS_FLAG_BIG_ENDIAN | LPS_FLAG_PFILE)) ? 1 : 0;
        long long blocks = n < h->sm_count * 8 ? n : h->sm_count * 8;
        lps_norm_kernel<<<(int)blocks, 288, 0, st>>>(d_out, f0, f1, pf ? LPS_BINS + 2 : LPS_BINS, pf ? 2 : 0, swap, h->d_norm_acc);
        h->norm_frames += n;
    }
    LPS_CUDA(cudaGetLastError());
    return 0;
}

static int run_kernel(lps_handle *h, const int16_t *d_pcm, int n_utts, long long total, float *d_out, int flags)
{
    int rc = check_flags(h, flags);
    if (rc) return rc;
    LPS_CUDA(cudaEventRecord(h->e0, h->s));
    rc = launch_frames(h, d_pcm, n_utts, 0, total, d_out, flags, h->s);
    if (rc) return rc;
    LPS_CUDA(cudaEventRecord(h->e1, h->s));
    return 0;
}

int lps_extract_batch_device(lps_handle *h, const int16_t *d_pcm, const long *utt_off, int n_utts, float *d_out, int flags, long *total_frames)
{
    if (!h || !utt_off || n_utts < 0 || (n_utts > 0 && (!d_pcm || !d_out))) { lps_err("lps_extract_batch_device: bad argument"); return -1; }
    LPS_CUDA(cudaSetDevice(h->gpu));
    std::vector<long long> tab;
    long long total = 0;
    int rc = offsets(h, utt_off, n_utts, tab, &total);
    if (rc) return rc;
    rc = run_kernel(h, d_pcm, n_utts, total, d_out, flags);
    if (rc) return rc;
    LPS_CUDA(cudaStreamSynchronize(h->s));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->e0, h->e1);
    h->last_ms = ms;
    if (total_frames) *total_frames = (long)total;
    return 0;
}

int lps_extract_batch(lps_handle *h, const int16_t *pcm, const long *utt_off, int n_utts, float *out, int flags, long *total_frames)
{
    if (!h || !utt_off || n_utts < 0 || (n_utts > 0 && (!pcm || !out))) { lps_err("lps_extract_batch: bad argument"); return -1; }
    LPS_CUDA(cudaSetDevice(h->gpu));
    std::vector<long long> tab;
    long long total = 0;
    int rc = offsets(h, utt_off, n_utts, tab, &total);
    if (rc) return rc;
    const size_t ns = (size_t)(utt_off[n_utts] - utt_off[0]);
    // rebase offsets to the uploaded span
    if (utt_off[0] != 0) {
        for (int u = 0; u <= n_utts; u++) tab[u] -= utt_off[0];
        LPS_CUDA(cudaMemcpyAsync(h->d_off, tab.data(), tab.size() * sizeof(long long), cudaMemcpyHostToDevice, h->s));
    }
    if (ns + 16 > h->pcm_cap) { cudaFree(h->d_pcm); h->d_pcm = nullptr; LPS_CUDA(cudaMalloc(&h->d_pcm, (ns + 16) * sizeof(int16_t))); h->pcm_cap = ns + 16; }
    const size_t pitch = (flags & LPS_FLAG_PFILE) ? LPS_BINS + 2 : LPS_BINS;
    const size_t no = (size_t)total * pitch;
    if (no > h->out_cap) { cudaFree(h->d_out); h->d_out = nullptr; LPS_CUDA(cudaMalloc(&h->d_out, (no + 1) * sizeof(float))); h->out_cap = no + 1; }
    rc = check_flags(h, flags);
    if (rc) return rc;
    // Pipeline in pieces of PIECE frames on two streams: the upload of piece i+1 and the download of piece i-1 run
    // under the kernel of piece i (PCIe is full duplex; the whole call is bound by the 1 028 B/frame going back).
    const long long PIECE = 32768;
    const int npieces = (int)((total + PIECE - 1) / PIECE);
    while ((int)h->pe.size() < 2 * npieces + 2) { cudaEvent_t e; LPS_CUDA(cudaEventCreate(&e)); h->pe.push_back(e); }
    LPS_CUDA(cudaStreamSynchronize(h->s));     // the offset table (uploaded on h->s) is visible to both streams
    // sample position (relative to the uploaded span) of the first sample of global frame f
    auto frame_sample = [&](long long f) {
        int lo = 0, hi = n_utts;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tab[n_utts + 1 + mid] <= f) lo = mid; else hi = mid; }
        return tab[lo] + (f - tab[n_utts + 1 + lo]) * LPS_FRAME_SHIFT;
    };
    for (int p = 0; p < npieces; p++) {
        cudaStream_t st = (p & 1) ? h->s2 : h->s;
        const long long f0 = (long long)p * PIECE, f1 = (f0 + PIECE < total) ? f0 + PIECE : total;
        const long long s0 = frame_sample(f0), s1 = frame_sample(f1 - 1) + LPS_FRAME_LEN;
        LPS_CUDA(cudaMemcpyAsync(h->d_pcm + s0, pcm + utt_off[0] + s0, (size_t)(s1 - s0) * sizeof(int16_t), cudaMemcpyHostToDevice, st));
        LPS_CUDA(cudaEventRecord(h->pe[2 * p], st));
        rc = launch_frames(h, h->d_pcm, n_utts, f0, f1, h->d_out, flags, st);
        if (rc) return rc;
        LPS_CUDA(cudaEventRecord(h->pe[2 * p + 1], st));
        LPS_CUDA(cudaMemcpyAsync(out + f0 * pitch, h->d_out + f0 * pitch, (size_t)(f1 - f0) * pitch * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    LPS_CUDA(cudaStreamSynchronize(h->s));
    LPS_CUDA(cudaStreamSynchronize(h->s2));
    double ms_sum = 0;
    for (int p = 0; p < npieces; p++) { float ms = 0; cudaEventElapsedTime(&ms, h->pe[2 * p], h->pe[2 * p + 1]); ms_sum += ms; }
    h->last_ms = ms_sum;
    if (total_frames) *total_frames = (long)total;
    return 0;
}

int lps_extract(lps_handle *h, const int16_t *pcm, long n_samples, float *out, int flags)
{
    const long off[2] = {0, n_samples};
    return lps_extract_batch(h, pcm, off, 1, out, flags, nullptr);
}

double lps_last_kernel_ms(lps_handle *h) { return h ? h->last_ms : 0.0; }

int lps_norm_reset(lps_handle *h)
{
    if (!h) { lps_err("lps_norm_reset: null handle"); return -1; }
    LPS_CUDA(cudaSetDevice(h->gpu));
    LPS_CUDA(cudaMemset(h->d_norm_acc, 0, 2 * LPS_BINS * sizeof(double)));
    h->norm_frames = 0;
    return 0;
}

int lps_norm_accumulate_device(lps_handle *h, const float *d_feats, long n_frames, int pitch, int skip, int big_endian)
{
    if (!h || !d_feats || n_frames < 0 || pitch < LPS_BINS || skip < 0 || skip + LPS_BINS > pitch) { lps_err("lps_norm_accumulate_device: bad argument"); return -1; }
    if (n_frames == 0) return 0;
    LPS_CUDA(cudaSetDevice(h->gpu));
    long long blocks = n_frames < h->sm_count * 8 ? n_frames : h->sm_count * 8;
    lps_norm_kernel<<<(int)blocks, 288, 0, h->s>>>(d_feats, 0, n_frames, pitch, skip, big_endian ? 1 : 0, h->d_norm_acc);
    LPS_CUDA(cudaGetLastError());
    LPS_CUDA(cudaStreamSynchronize(h->s));
    h->norm_frames += n_frames;
    return 0;
}

int lps_norm_finalize(lps_handle *h, float *mean, float *dvar, long *n_frames)
{
    if (!h || !mean || !dvar) { lps_err("lps_norm_finalize: null argument"); return -1; }
    LPS_CUDA(cudaSetDevice(h->gpu));
    LPS_CUDA(cudaStreamSynchronize(h->s));
    LPS_CUDA(cudaStreamSynchronize(h->s2));
    double acc[2 * LPS_BINS];
    LPS_CUDA(cudaMemcpy(acc, h->d_norm_acc, sizeof acc, cudaMemcpyDeviceToHost));
    if (n_frames) *n_frames = (long)h->norm_frames;
    if (h->norm_frames <= 0) { lps_err("lps_norm_finalize: no frames accumulated (LPS_FLAG_ACCUM_NORM)"); return -1; }
    const double n = (double)h->norm_frames;
    for (int k = 0; k < LPS_BINS; k++) {
        const double m = acc[k] / n, var = acc[LPS_BINS + k] / n - m * m;       // ddof 0, like qnnorm
        mean[k] = (float)m;
        dvar[k] = (float)(1.0 / sqrt(var > 0 ? var : 1e-300));                   // the .norm file holds the RECIPROCAL std (Interface.cc:385-396)
    }
    return 0;
}

}  // extern "C"
