// dp_update.cu -- see dp_update.cuh
#include "dp_update.cuh"

namespace ggd {

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// wait until every rank's counter in flags[phase][*] has reached `step`; bounded so that a lost peer cannot hang the GPU
__device__ bool wait_all(const DpArgs &a, int phase, unsigned int step)
{
    const unsigned int *f = a.flags[a.rank] + phase * DP_MAX_RANKS;
    const long long t0 = clock64();
    for (int p = 0; p < a.world; p++) {
        while ((int)(ld_acquire_sys(f + p) - step) < 0) {
            if (clock64() - t0 > (1ll << 32)) { *a.error_flag = 1u + p; return false; }   // ~2 s at 2 GHz
            __nanosleep(100);
        }
    }
    return true;
}

__global__ void __launch_bounds__(256) dp_update_kernel(const DpArgs a)
{
    __shared__ int s_last;
    const unsigned int step = *a.step_counter + 1;
    // ---- phase A: my gradients are complete (stream order); publish that and wait for everyone's
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int p = 0; p < a.world; p++) st_release_sys(a.flags[p] + 0 * DP_MAX_RANKS + a.rank, step);
    }
    if (threadIdx.x == 0) wait_all(a, 0, step);
    __syncthreads();

    const float mom = a.mom, lr = a.lr, inv_mg = 1.0f / a.Mg;
    for (int pc = 0; pc < a.npieces; pc++) {
        const DpPiece pi = a.piece[pc];
        const long long n4 = pi.n >> 2;
        float4 *P = reinterpret_cast<float4 *>(a.P[a.rank] + pi.off);
        float4 *D = reinterpret_cast<float4 *>(a.Dl + pi.off);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
            // gradient of the GLOBAL minibatch: peers are read over NVLink, always in rank order
            float4 g = __ldcs(reinterpret_cast<const float4 *>(a.G[0] + pi.goff) + i);
            for (int p = 1; p < a.world; p++) {
                const float4 t = __ldcs(reinterpret_cast<const float4 *>(a.G[p] + pi.goff) + i);
                g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
            }
            float4 w = P[i], d = D[i];
            d.x = mom * d.x - lr * (g.x * inv_mg + pi.wc * w.x);
            d.y = mom * d.y - lr * (g.y * inv_mg + pi.wc * w.y);
            d.z = mom * d.z - lr * (g.z * inv_mg + pi.wc * w.z);
            d.w = mom * d.w - lr * (g.w * inv_mg + pi.wc * w.w);
            w.x += d.x; w.y += d.y; w.z += d.z; w.w += d.w;
            D[i] = d;
            P[i] = w;
            if (pi.shadow) {
                const __nv_bfloat162 h01 = __floats2bfloat162_rn(w.x, w.y), h23 = __floats2bfloat162_rn(w.z, w.w);
                const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                const __nv_bfloat162 l01 = __floats2bfloat162_rn(w.x - f01.x, w.y - f01.y), l23 = __floats2bfloat162_rn(w.z - f23.x, w.w - f23.y);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<const uint32_t *>(&h01); hv.y = *reinterpret_cast<const uint32_t *>(&h23);
                lv.x = *reinterpret_cast<const uint32_t *>(&l01); lv.y = *reinterpret_cast<const uint32_t *>(&l23);
                for (int p = 0; p < a.world; p++) {     // all-gather of the operand shadows: store into every rank's copy
                    reinterpret_cast<uint2 *>(a.hi[p] + pi.off)[i] = hv;
                    reinterpret_cast<uint2 *>(a.lo[p] + pi.off)[i] = lv;
                }
            } else {
                for (int p = 0; p < a.world; p++)
                    if (p != a.rank) reinterpret_cast<float4 *>(a.P[p] + pi.off)[i] = w;   // biases are consumed in fp32
            }
        }
    }
    // ---- phase B: publish "my slice is written everywhere" once ALL blocks of this launch have fenced their stores
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(a.block_counter, 1u) + 1u;
        s_last = (done == gridDim.x);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence_system();
        for (int p = 0; p < a.world; p++) st_release_sys(a.flags[p] + 1 * DP_MAX_RANKS + a.rank, step);
        wait_all(a, 1, step);        // nobody may start the next forward before every slice has landed
        *a.block_counter = 0;
        *a.step_counter = step;
        if (a.ctl) a.ctl->bunch_idx += 1;
        __threadfence();
    }
}

void launch_dp_update(const DpArgs &a, int blocks, cudaStream_t s)
{
    dp_update_kernel<<<blocks, 256, 0, s>>>(a);
}

}  // namespace ggd
