// pipe.cuh -- device helpers shared by the persistent TMA-pipelined kernels (dw_persist.cu, dw_wide.cu, dp_factor.cu):
// bounded mbarrier waits with a host-visible watchdog record, TMEM loads, TMA stores and L2 cache policies.
#pragma once
#include "common.cuh"

namespace ggd {

// Bounded mbarrier wait: a pipeline bug must surface as an error, never as a hung GPU.  After ~8.7 s of failed probes
// the waiter records {code, block, iteration, parity} in host-mapped memory and traps.
static __device__ __noinline__ void hang_report(unsigned int *rec, int code, int it, uint32_t parity)
{
    if (rec) {
        rec[1] = (unsigned int)code; rec[2] = blockIdx.x; rec[3] = (unsigned int)it; rec[4] = parity; rec[5] = threadIdx.x;
        __threadfence_system();
        rec[0] = 0xDEADu;
        __threadfence_system();
    }
    __trap();
}
// The watchdog lives in an out-of-line slow path: inlined into the hot loops (a call site inside every spin loop) it cost
// ~0.7 us per GEMM launch through register allocation around the call; the first probe is inline, everything else is not.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar_saddr, uint32_t parity, unsigned int *rec, int code, int it)
{
    unsigned int spins = 0;
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_saddr), "r"(parity)
            : "memory");
        if (ok) return;
        ++spins;
        if (spins == (1u << 16) && rec && (threadIdx.x & 31) == 0) {   // note who is waiting on what, per warp
            unsigned int *w = rec + 8 + (blockIdx.x * 12 + (threadIdx.x >> 5)) * 4;
            w[0] = (unsigned int)code; w[1] = (unsigned int)it; w[2] = parity; w[3] = 1;
            __threadfence_system();
        }
        // time based (~8.7 s at 1.97 GHz): LONGER than the 4.4 s a data-parallel kernel may legitimately wait for a peer's
        // flags (that wait reports the missing rank itself), so a slow peer is never mistaken for a pipeline bug
        if ((spins & 1023u) == 0 && clock64() - t0 > (1ll << 34)) hang_report(rec, code, it, parity);
    }
}
static __device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity, unsigned int *rec, int code, int it)
{
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(smem_u32(bar), parity, rec, code, it);
}

// 32 lanes x 8 consecutive fp32 columns -> 8 registers per thread
static __device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v)
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}

// TMA 2-D tiled store shared -> global (bulk async-group completion)
static __device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, const void *smem_src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
// L2 cache policies: the fp32 master weights / momentum are touched once per step (evict first); the bf16 shadows are
// re-read by the forward and backward GEMMs of the next step and fit the 126 MB L2 (evict last).
static __device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
static __device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
static __device__ __forceinline__ void tma_load_2d_hint(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
        : "memory");
}
static __device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap *m, const void *smem_src, int c0, int c1, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(pol)
                 : "memory");
}
static __device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
static __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
static __device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// system-scope flags in (peer) global memory
static __device__ __forceinline__ void st_release_sys_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// relaxed flag store: the caller has ALREADY issued one system-scope fence after the data stores (fence + relaxed store is a
// release pattern); a st.release.sys per peer pays one NVLink round trip EACH (~5 us: 35-40 us to flag 7 peers)
static __device__ __forceinline__ void st_relaxed_sys_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
static __device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ float ld_relaxed_sys_f32(const float *p)
{
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

}  // namespace ggd
