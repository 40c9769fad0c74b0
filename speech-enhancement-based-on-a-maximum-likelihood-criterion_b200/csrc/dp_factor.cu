// dp_factor.cu -- see dp_factor.cuh: factor push over NVLink peer memory and the whole-minibatch bias update.
#include "dp_factor.cuh"
#include "pipe.cuh"
#include "../../include/ggd_train.h"

namespace ggd {

__device__ __forceinline__ void st_na_v4(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256) factor_push_kernel(const FxPushArgs a)
{
    const unsigned int step = *a.step_counter + 1u;
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[a.event * 4 + 0] = gtime();
    if (a.wait_done) {
        // every peer must have finished step-1 (it no longer reads the arena slices this step overwrites)
        if (threadIdx.x < a.world && (int)threadIdx.x != a.rank) {
            const unsigned int *f = a.my_flags + threadIdx.x * FX_STRIDE + FX_EV_DONE;
            const long long t0 = clock64();
            while ((int)(ld_acquire_sys_u32(f) - (step - 1u)) < 0) {
                if (clock64() - t0 > (1ll << 33)) {
                    *a.error_flag = 1u + threadIdx.x;
                    __threadfence_system();
                    hang_report(a.hang, 200, (int)step, threadIdx.x);
                }
                __nanosleep(64);
            }
        }
        __syncthreads();
    }
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[a.event * 4 + 1] = gtime();
    const long long bunch = a.ctl->bunch_idx;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int sg = 0; sg < a.nseg; sg++) {
        const FxSeg S = a.seg[sg];
        const uint4 *src = reinterpret_cast<const uint4 *>(S.src + S.src_bunch_stride * bunch);
        const long long n16 = S.bytes >> 4;
        for (long long i = tid; i < n16; i += 4ll * nth) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const long long j = i + (long long)u * nth;
                if (j < n16) v[u] = __ldg(src + j);
            }
            for (int p = 0; p < a.world; p++) {
                if (p == a.rank && !a.include_self) continue;
                uint4 *dst = reinterpret_cast<uint4 *>(a.peer_arena[p] + S.dst_off);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const long long j = i + (long long)u * nth;
                    if (j < n16) st_na_v4(dst + j, v[u]);
                }
            }
        }
    }
    // flag: once EVERY block's stores are performed system-wide, tell all peers.  One fence per block: the barrier orders
    // the block's stores before thread 0's system-scope fence (cumulativity), 24 000 concurrent membar.sys cost ~15 us
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(a.block_counter, 1u);
        if (prev == gridDim.x - 1) {
            *a.block_counter = 0;
            __threadfence_system();
            if (a.trace) a.trace[a.event * 4 + 2] = gtime();
            for (int p = 0; p < a.world; p++)
                if (p != a.rank) st_relaxed_sys_u32(a.peer_flags[p] + a.rank * FX_STRIDE + a.event, step);
            if (a.trace) a.trace[a.event * 4 + 3] = gtime();
        }
    }
}

void launch_factor_push(const FxPushArgs &a, int grid, cudaStream_t s)
{
    factor_push_kernel<<<grid, 256, 0, s>>>(a);
}

}  // namespace ggd
