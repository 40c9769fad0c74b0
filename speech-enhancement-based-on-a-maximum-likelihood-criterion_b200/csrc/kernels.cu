// kernels.cu -- the HBM-bound kernels of the training step and the fp32 validation GEMM.
//
//   split_rows_kernel   chunk input fp32 -> bf16 hi/lo rows (operand format of the tcgen05 GEMMs)
//   loss_kernel         ONE kernel for the reference's 11-launch loss-gradient chain
//                       (BP_GPU.cu:408-424: DevSubClean2, DevVecMulNum, Deverror, Devabsolutevalus,
//                        Devindex2, DevSumcol, DevDivide, DevVecMulNum, Devindex2, Devfunc2, DevVecMulNum)
//   bias_grad_kernel    kernAccSumrow (DevFunc.cu:267-285)
//   update_kernel       kernUpdatedelta + kernAccSum for weights AND biases in one launch
//                       (DevFunc.cu:490-507, 427-443; BP_GPU.cu:433-437)
#include "kernels.cuh"
#include "pipe.cuh"
#include <math.h>

namespace ggd {

// ------------------------------------------------------------------------------------------------
__global__ void split_rows_kernel(const float *__restrict__ src, int rows, int cols, bf16 *__restrict__ hi,
                                  bf16 *__restrict__ lo, int ld)
{
    const int half = ld >> 1;  // pairs per row
    const long long total = (long long)rows * half;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / half), c = (int)(i % half) * 2;
        const float *p = src + (long long)r * cols;
        const float a = (c < cols) ? __ldg(p + c) : 0.0f;
        const float b = (c + 1 < cols) ? __ldg(p + c + 1) : 0.0f;
        bf16 ah, al, bh, bl;
        split_bf16(a, ah, al);
        split_bf16(b, bh, bl);
        reinterpret_cast<uint32_t *>(hi + (long long)r * ld)[c >> 1] = pack_bf16x2(ah, bh);
        reinterpret_cast<uint32_t *>(lo + (long long)r * ld)[c >> 1] = pack_bf16x2(al, bl);
    }
}

void launch_split_rows(const float *src, int rows, int cols, bf16 *hi, bf16 *lo, int ld, cudaStream_t s)
{
    if (rows <= 0) return;
    long long total = (long long)rows * (ld / 2);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_rows_kernel<<<blocks, 256, 0, s>>>(src, rows, cols, hi, lo, ld);
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float be_word_zscore(unsigned int w, float mean, float dvar)
{
    const float v = __uint_as_float(__byte_perm(w, 0, 0x0123));      // big-endian float32 (Interface.cc:744, SwapBytes)
    return __fmul_rn(__fsub_rn(v, mean), dvar);                      // two roundings, exactly as the host loader (no FMA)
}

__global__ void __launch_bounds__(256) expand_chunk_kernel(const ExpandArgs a)
{
    const int r = blockIdx.x;
    const int f0 = a.first[r];
    const int in_dim = a.fea_dim * a.ctx;
    const unsigned int *src = a.fea_rec + (size_t)f0 * (2 + a.fea_dim);
    for (int e = threadIdx.x; e < a.ld; e += blockDim.x) {
        float v = 0.0f;
        if (e < in_dim) {
            const int c = e / a.fea_dim, j = e - c * a.fea_dim;
            v = be_word_zscore(__ldg(src + (size_t)c * (2 + a.fea_dim) + 2 + j), __ldg(a.mean + j), __ldg(a.dvar + j));
            if (a.in32) a.in32[(size_t)r * in_dim + e] = v;
        }
        if (a.in_hi) {
            bf16 h, l;
            split_bf16(v, h, l);
            a.in_hi[(size_t)r * a.ld + e] = h;
            a.in_lo[(size_t)r * a.ld + e] = l;
        }
    }
    const unsigned int *ts = a.targ_rec + (size_t)(f0 + a.targ_offset) * (2 + a.D) + 2;
    for (int j = threadIdx.x; j < a.D; j += blockDim.x) {
        const int k = j % a.fea_dim;                                  // targets use the NOISY mean / dVar too (:807-808)
        a.targ[(size_t)r * a.D + j] = be_word_zscore(__ldg(ts + j), __ldg(a.mean + k), __ldg(a.dvar + k));
    }
}

__global__ void __launch_bounds__(256) expand_edges_kernel(const EdgeExpandArgs a)
{
    const int t = blockIdx.x, half = (a.ctx - 1) / 2;
    const int in_dim = a.fea_dim * a.ctx;
    for (int e = threadIdx.x; e < a.ld; e += blockDim.x) {
        float v = 0.0f;
        if (e < in_dim) {
            const int c = e / a.fea_dim, j = e - c * a.fea_dim;
            int f = t + c - half;                                   // frame_expand.m:8-23: left context, centre, right context
            f = f < 0 ? 0 : (f > a.frames - 1 ? a.frames - 1 : f);  // edges replicate the first / last frame
            v = __fmul_rn(__fsub_rn(__ldg(a.lps + (size_t)f * a.fea_dim + j), __ldg(a.mean + j)), __ldg(a.dvar + j));   // decode.m:31-33
            if (a.in32) a.in32[(size_t)t * in_dim + e] = v;
        }
        if (a.in_hi) {
            bf16 h, l;
            split_bf16(v, h, l);
            a.in_hi[(size_t)t * a.ld + e] = h;
            a.in_lo[(size_t)t * a.ld + e] = l;
        }
    }
}

void launch_expand_edges(const EdgeExpandArgs &a, cudaStream_t s)
{
    if (a.frames <= 0) return;
    expand_edges_kernel<<<a.frames, 256, 0, s>>>(a);
}

__global__ void denorm_kernel(float *out, long long n, int D, const float *mean, const float *dvar, int fea_dim)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % D) % fea_dim;
        out[i] = __fadd_rn(__fdiv_rn(out[i], __ldg(dvar + k)), __ldg(mean + k));
    }
}

void launch_denorm(float *out, long long frames, int D, const float *mean, const float *dvar, int fea_dim, cudaStream_t s)
{
    const long long n = frames * D;
    if (n <= 0) return;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    denorm_kernel<<<(int)blocks, 256, 0, s>>>(out, n, D, mean, dvar, fea_dim);
}

void launch_expand_chunk(const ExpandArgs &a, cudaStream_t s)
{
    if (a.samples <= 0) return;
    expand_chunk_kernel<<<a.samples, 256, 0, s>>>(a);
}

// ------------------------------------------------------------------------------------------------
// |e|^p: exact for the two named configurations (beta = 2, beta = 1); otherwise exp2(p*log2 a) with the hardware
// approximations (relative error ~1e-6 * p*|log2 a|, far inside the 1e-3 tolerance; the reference uses powf).
__device__ __forceinline__ float pow_abs(float a, float p)
{
    if (p == 2.0f) return a * a;
    if (p == 1.0f) return a;
    if (p == 0.0f) return 1.0f;
    return (a > 0.0f) ? exp2f(p * __log2f(a)) : 0.0f;
}

constexpr int LOSS_COLS = 16;   // columns per block (64-byte row segments): 17 blocks for 257 outputs instead of 9 -- the kernel is
                                // latency bound (9 blocks of 32 columns took 39 us at 1024 frames)
constexpr int LOSS_ROWL = 64;   // row lanes per block (1024 threads)
constexpr int LOSS_RMAX = 16;   // errors kept in registers per thread (bunches up to 1024 frames make one pass over memory)
static_assert(LOSS_COLS == 16, "the alpha exchange uses one flag per 16-column block (FX_EV_LOSS + block), like the fused epilogue");

__global__ void __launch_bounds__(LOSS_COLS *LOSS_ROWL) loss_kernel(LossArgs a)
{
    __shared__ float red[LOSS_ROWL][LOSS_COLS + 1];
    __shared__ float s_col[LOSS_COLS], s_pa[LOSS_COLS];
    const int tx = threadIdx.x % LOSS_COLS, ty = threadIdx.x / LOSS_COLS;
    const int d = blockIdx.x * LOSS_COLS + tx;
    const int bunch = a.ctl->bunch_idx;
    const float *targ = a.ctl->targ + (size_t)bunch * a.M * a.D;
    const bool live = d < a.D;
    const float beta = a.beta;
    const bool in_regs = a.M <= LOSS_ROWL * LOSS_RMAX;
    float er[LOSS_RMAX];

    // pass 1: e = out - targ;  s_d = sum_m |e_md|^beta  (Deverror, Devabsolutevalus, Devindex2, DevSumcol)
    {
        float s = 0.0f;
        if (live) {
            if (in_regs) {
#pragma unroll
                for (int i = 0; i < LOSS_RMAX; i++) {
                    const int m = ty + i * LOSS_ROWL;
                    er[i] = (m < a.M) ? a.out[(size_t)m * a.ldo + d] - __ldg(targ + (size_t)m * a.D + d) : 0.0f;
                }
                if (a.mode != 2) {
#pragma unroll
                    for (int i = 0; i < LOSS_RMAX; i++) s += pow_abs(fabsf(er[i]), beta);
                }
            } else if (a.mode != 2) {
                for (int m = ty; m < a.M; m += LOSS_ROWL) {
                    const float e = a.out[(size_t)m * a.ldo + d] - __ldg(targ + (size_t)m * a.D + d);
                    s += pow_abs(fabsf(e), beta);
                }
            }
        }
        if (a.mode != 2) {
            red[ty][tx] = s;
            __syncthreads();
            if (ty == 0) {
                float t = red[0][tx];
#pragma unroll
                for (int r = 1; r < LOSS_ROWL; r++) t += red[r][tx];
                s_col[tx] = t;
                if (live && a.mode == 1) a.colsum[d] = t;
            }
            __syncthreads();
            if (a.mode == 1) return;
            if (a.mode == 3) {
                // exchange the partial sums with every rank (warp 0 of this block owns the block's 32 columns)
                if (ty == 0) {
                    const unsigned int step = *a.step_counter + 1u;
                    if (live)
                        for (int p = 0; p < a.world; p++) a.asum_slot[p][(size_t)a.rank * a.D + d] = s_col[tx];
                    __threadfence_system();
                    __syncwarp(0x0000ffffu);
                    if (tx == 0)
                        for (int p = 0; p < a.world; p++)
                            if (p != a.rank) st_relaxed_sys_u32(a.lflags[p] + a.rank * FX_STRIDE + FX_EV_LOSS + blockIdx.x, step);   // (fenced above)
                    if (tx < a.world && tx != a.rank) {
                        const unsigned int *f = a.lflags[a.rank] + tx * FX_STRIDE + FX_EV_LOSS + blockIdx.x;
                        const long long t0 = clock64();
                        while ((int)(ld_acquire_sys_u32(f) - step) < 0) {
                            if (clock64() - t0 > (1ll << 32)) { *a.error_flag = 1u + tx; break; }
                            __nanosleep(32);
                        }
                    }
                    __syncwarp(0x0000ffffu);
                    if (live) {
                        float t = 0.0f;
                        for (int p = 0; p < a.world; p++) t += ld_relaxed_sys_f32(a.asum_slot[a.rank] + (size_t)p * a.D + d);   // rank order
                        s_col[tx] = t;
                    }
                }
                __syncthreads();
            }
        } else {
            if (ty == 0) s_col[tx] = live ? a.colsum[d] : 0.0f;
            __syncthreads();
        }
    }

    // alpha_d = (beta * s_d / Mg)^(1/beta)   (DevDivide, DevVecMulNum, Devindex2; BP_GPU.cu:417-420)
    if (ty == 0) {
        float pa = 1.0f, contrib = 0.0f;
        if (live) {
            const float s = s_col[tx];
            if (a.ml) {
                const float v1 = s / (float)a.Mg;
                const float v2 = v1 * beta;
                const float al = powf(v2, 1.0f / beta);
                a.alpha[d] = al;
                pa = (beta == 2.0f) ? al * al : ((beta == 1.0f) ? al : powf(al, beta));
                contrib = logf(al) + s / ((float)a.Mg * pa);   // ln alpha_d + sum_m (|e|/alpha_d)^beta / M
            } else {
                contrib = s / (float)a.Mg;                      // E_beta = sum |e|^beta / M
            }
        }
        s_pa[tx] = pa;
        if (a.trace) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) contrib += __shfl_xor_sync(0x0000ffffu, contrib, o);     // ty == 0: lanes 0..15
            if (tx == 0) atomicAdd(a.trace + bunch, (double)contrib);
        }
    }
    __syncthreads();

    // pass 2: dE/dx = (1/Mg) * sgn(e) |e|^(beta-1) * beta [/ alpha_d^beta]
    //         (DevSubClean2 + DevVecMulNum, or Devfunc2 + DevVecMulNum when MLflag == 1); exactly 0 at e == 0
    if (!live) return;
    const float invM = 1.0f / a.Mg;
    const float scale = a.ml ? beta / s_pa[tx] : beta;
    auto emit = [&](int m, float e) {
        float r = 0.0f;
        if (e != 0.0f) {
            r = pow_abs(fabsf(e), beta - 1.0f) * scale;
            r = (e > 0.0f) ? r : -r;
        }
        r *= invM;
        const size_t o = (size_t)m * a.ldx + d;
        if (a.dx32) a.dx32[o] = r;
        if (a.dx_hi) {
            bf16 h, l;
            split_bf16(r, h, l);
            a.dx_hi[o] = h;
            a.dx_lo[o] = l;
        }
    };
    if (in_regs) {
#pragma unroll
        for (int i = 0; i < LOSS_RMAX; i++) {
            const int m = ty + i * LOSS_ROWL;
            if (m < a.M) emit(m, er[i]);
        }
    } else {
        for (int m = ty; m < a.M; m += LOSS_ROWL) emit(m, a.out[(size_t)m * a.ldo + d] - __ldg(targ + (size_t)m * a.D + d));
    }
}

void launch_loss(const LossArgs &a, cudaStream_t s)
{
    loss_kernel<<<ceil_div(a.D, LOSS_COLS), LOSS_COLS * LOSS_ROWL, 0, s>>>(a);
}

// ------------------------------------------------------------------------------------------------
constexpr int BG_COLS = 32, BG_ROWL = 8;
__global__ void __launch_bounds__(BG_COLS *BG_ROWL) bias_grad_kernel(const BiasGradArgs a)
{
    __shared__ float red[BG_ROWL][BG_COLS + 1];
    const BiasGradLayer L = a.layer[blockIdx.y];
    const int tx = threadIdx.x % BG_COLS, ty = threadIdx.x / BG_COLS;
    const int n = blockIdx.x * BG_COLS + tx;
    if (a.ctl && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) a.ctl->bunch_idx += 1;
    if (blockIdx.x * BG_COLS >= L.N) return;
    float s = 0.0f;
    if (n < L.N) {
        if (L.dx32) {
            for (int m = ty; m < a.M; m += BG_ROWL) s += L.dx32[(size_t)m * L.ld + n];
        } else {
            for (int m = ty; m < a.M; m += BG_ROWL) s += join_bf16(L.hi[(size_t)m * L.ld + n], L.lo[(size_t)m * L.ld + n]);
        }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < L.N) {
        float t = red[0][tx];
#pragma unroll
        for (int r = 1; r < BG_ROWL; r++) t += red[r][tx];
        if (a.apply) {
            const float d = a.mom * L.db[n] - a.lr * (t / a.Mg);
            L.db[n] = d;
            L.b[n] = d + L.b[n];
        } else {
            L.dst[n] = t;
        }
    }
}

void launch_bias_grad(const BiasGradArgs &a, cudaStream_t s)
{
    int maxn = 0;
    for (int l = 0; l < a.nlayers; l++) maxn = a.layer[l].N > maxn ? a.layer[l].N : maxn;
    dim3 grid(ceil_div(maxn, BG_COLS), a.nlayers);
    bias_grad_kernel<<<grid, BG_COLS * BG_ROWL, 0, s>>>(a);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) update_kernel(const UpdArgs a)
{
    const UpdSeg sg = a.seg[blockIdx.y];
    const long long n4 = sg.n >> 2;
    float4 *P = reinterpret_cast<float4 *>(a.P + sg.off);
    float4 *D = reinterpret_cast<float4 *>(a.Dl + sg.off);
    const float4 *G = reinterpret_cast<const float4 *>(a.G + sg.goff);
    uint2 *Hi = reinterpret_cast<uint2 *>(a.Phi + sg.off);
    uint2 *Lo = reinterpret_cast<uint2 *>(a.Plo + sg.off);
    const float mom = a.mom, lr = a.lr, Mg = a.Mg, wc = sg.wc;
    // last kernel of the step: move the device-side bunch counter on (no kernel of this step reads it any more,
    // and the next step's kernels are stream-ordered behind this one)
    if (a.ctl && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) a.ctl->bunch_idx += 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 w = P[i], d = D[i];
        const float4 g = __ldcs(G + i);
        d.x = mom * d.x - lr * (g.x / Mg + wc * w.x);
        d.y = mom * d.y - lr * (g.y / Mg + wc * w.y);
        d.z = mom * d.z - lr * (g.z / Mg + wc * w.z);
        d.w = mom * d.w - lr * (g.w / Mg + wc * w.w);
        w.x = d.x + w.x; w.y = d.y + w.y; w.z = d.z + w.z; w.w = d.w + w.w;
        D[i] = d;
        P[i] = w;
        if (sg.shadow) {
            bf16 h0, l0, h1, l1, h2, l2, h3, l3;
            split_bf16(w.x, h0, l0); split_bf16(w.y, h1, l1); split_bf16(w.z, h2, l2); split_bf16(w.w, h3, l3);
            Hi[i] = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
            Lo[i] = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
        }
    }
}

void launch_update(const UpdArgs &a, int sm_count, cudaStream_t s)
{
    dim3 grid(sm_count * 4, a.nseg);
    update_kernel<<<grid, 256, 0, s>>>(a);
}

__global__ void advance_kernel(StepCtl *ctl) { ctl->bunch_idx += 1; }
void launch_advance(StepCtl *ctl, cudaStream_t s) { advance_kernel<<<1, 1, 0, s>>>(ctl); }

// ------------------------------------------------------------------------------------------------
// fp32 validation path (CUDA cores). 32x32 output tile, 16x16 threads, 2x2 outputs each.
__global__ void __launch_bounds__(256) simt_gemm_kernel(const float *__restrict__ A, long sAi, long sAr, const float *__restrict__ B,
                                                        long sBj, long sBr, float *__restrict__ C, long ldc, int I, int J, int R)
{
    __shared__ float As[32][33], Bs[32][33];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    float acc[2][2] = {{0, 0}, {0, 0}};
    for (int r0 = 0; r0 < R; r0 += 32) {
        for (int t = threadIdx.x; t < 32 * 32; t += 256) {
            // pick the fast-varying index per operand so that global reads are as coalesced as the strides allow
            int ai, ar, bj, br;
            if (sAr == 1) { ar = t % 32; ai = t / 32; } else { ai = t % 32; ar = t / 32; }
            if (sBr == 1) { br = t % 32; bj = t / 32; } else { bj = t % 32; br = t / 32; }
            As[ai][ar] = (i0 + ai < I && r0 + ar < R) ? A[(long)(i0 + ai) * sAi + (long)(r0 + ar) * sAr] : 0.0f;
            Bs[bj][br] = (j0 + bj < J && r0 + br < R) ? B[(long)(j0 + bj) * sBj + (long)(r0 + br) * sBr] : 0.0f;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 32; r++) {
            const float a0 = As[ty][r], a1 = As[ty + 16][r], b0 = Bs[tx][r], b1 = Bs[tx + 16][r];
            acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int i = i0 + ty + 16 * p, j = j0 + tx + 16 * q;
            if (i < I && j < J) C[(long)i * ldc + j] = acc[p][q];
        }
}

void launch_simt_gemm(const float *A, long sAi, long sAr, const float *B, long sBj, long sBr, float *C, long ldc, int I, int J,
                      int R, cudaStream_t s)
{
    dim3 grid(ceil_div(J, 32), ceil_div(I, 32));
    simt_gemm_kernel<<<grid, 256, 0, s>>>(A, sAi, sAr, B, sBj, sBr, C, ldc, I, J, R);
}

__global__ void simt_bias_act_kernel(const float *__restrict__ x, int ld, const float *__restrict__ bias, float *__restrict__ y, int M,
                                     int N, int linear)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * N) return;
    const int m = (int)(i / N), n = (int)(i % N);
    const float v = x[(size_t)m * ld + n] + bias[n];
    y[(size_t)m * ld + n] = linear ? v : 1.0f / (1.0f + expf(-v));   // kernSigmoid, DevFunc.cu:36-51
}
void launch_simt_bias_act(const float *x, int ld, const float *bias, float *y, int M, int N, int linear, cudaStream_t s)
{
    long long n = (long long)M * N;
    simt_bias_act_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(x, ld, bias, y, M, N, linear);
}

__global__ void simt_dsigmoid_kernel(const float *__restrict__ y, const float *__restrict__ dedy, float *__restrict__ dedx, int ld,
                                     int M, int N)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * N) return;
    const int m = (int)(i / N), n = (int)(i % N);
    const size_t o = (size_t)m * ld + n;
    const float v = y[o];
    dedx[o] = (1.0f - v) * v * dedy[o];   // kernDsigmoid, DevFunc.cu:53-71
}
void launch_simt_dsigmoid(const float *y, const float *dedy, float *dedx, int ld, int M, int N, cudaStream_t s)
{
    long long n = (long long)M * N;
    simt_dsigmoid_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(y, dedy, dedx, ld, M, N);
}

__global__ void simt_gather_in_kernel(const StepCtl *ctl, int M, int K, float *__restrict__ dst, int ld)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * K) return;
    const int m = (int)(i / K), k = (int)(i % K);
    dst[(size_t)m * ld + k] = ctl->in32[((size_t)ctl->bunch_idx * M + m) * K + k];
}
void launch_simt_gather_in(const StepCtl *ctl, int M, int K, float *dst, int ld, cudaStream_t s)
{
    long long n = (long long)M * K;
    simt_gather_in_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(ctl, M, K, dst, ld);
}

}  // namespace ggd
