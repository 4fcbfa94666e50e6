// umma_common.cuh -- PTX wrappers (mbarrier, TMA, tcgen05) and tensor-map helpers shared by the
// tensor-core kernels (conv_umma.cu, conv_halo.cu).  sm_100a only.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace hoig {
namespace {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (-> launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if ((spin & 0x3ff) == 0x3ff) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 6000000000ll) __trap();
        }
    }
}
// Same wait with exponential back-off (nanosleep) between polls: for warps whose wait is long and not latency-critical (the epilogue
// warps waiting for a whole tile's MMAs).  Twenty spinning warps per SM cost issue slots and, on a power-capped part, clock.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, uint32_t max_ns)
{
    uint32_t done = 0, ns = 32;
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (max_ns) {
            __nanosleep(ns);
            if (ns < max_ns) ns <<= 1;
        }
        if ((spin & 0x3ff) == 0x3ff) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 6000000000ll) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// One lane of a CONVERGED warp.  Single-thread roles (TMA producer, MMA issuer) run their loops on the whole warp and
// guard only the issue with this: under `if (lane == 0)` the compiler wraps every uniform-datapath instruction
// (UTCHMMA, UTMALDG, UTCBAR) in an ELECT / BRA.U.ANY loop plus R2UR moves, which costs more than a narrow MMA itself.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version 1):
// start address >> 4 | LBO (unused for swizzled K-major, canonical 1) | SBO = 1024 B (8 rows x 128 B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;   // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- thread-block cluster / CTA-pair (cta_group::2) helpers
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of the same smem location in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Relaxed arrivals: they only SIGNAL (no memory ordering), so the arriving warp does not wait for its outstanding global stores.
// Used to hand a TMEM accumulator stage back once tcgen05.wait::ld has put its contents into registers.
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar)
{
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in this CTA's smem, completion bytes are signalled on `bar`, a shared::cluster
// address that may live in the peer (leader) CTA
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// 256 x N x 16 MMA over the pair: each CTA holds its 128 rows of A and its N/2 rows of B at the same smem offsets
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the barrier at this smem offset in BOTH CTAs of the pair when all prior MMAs have retired
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// 32 consecutive columns of this warp's 32 lanes (two x16 loads, one wait: tmem_ld_wait32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r)
{
    tmem_ld16(taddr, r);
    tmem_ld16(taddr + 16u, r + 16);
}
__device__ __forceinline__ void tmem_ld_wait32(uint32_t *r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
// wait::ld with the destination registers as read-write operands, so the compiler cannot hoist their uses
__device__ __forceinline__ void tmem_ld_wait(uint32_t *r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
// 32 bytes per lane in one instruction (sm_100: STG.256): a lane then writes one whole 32-byte sector, so row-per-lane epilogue stores need no
// lane-pair exchange to fill their sectors.  ``p`` must be 32-byte aligned.
__device__ __forceinline__ void st_global_v8(void *p, const uint32_t (&v)[8])
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
                 "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// 16 values per lane x 32 lanes -> column totals: after the call lane L holds the total of
// column  8*b4 + 4*b3 + 2*b2 + b1  (bits of L), duplicated on lanes L and L^1.
__device__ __forceinline__ float transpose_reduce16(const float v[16], int lane)
{
    float a[8], b[4], c[2];
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = h4 ? v[i] : v[i + 8];
        const float keep = h4 ? v[i + 8] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h3 ? a[i] : a[i + 4];
        const float keep = h3 ? a[i + 4] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h2 ? b[i] : b[i + 2];
        const float keep = h2 ? b[i + 2] : b[i];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float send = h1 ? c[0] : c[1];
    const float keep = h1 ? c[1] : c[0];
    float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}


// --------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int make_map(CUtensorMap *map, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes,
             const cuuint32_t *box, const char *what, int dtype)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return HOIG_ERR_CUDA; }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = fn(map, dtype == HOIG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), dims, strides_bytes, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r); return HOIG_ERR_CUDA; }
    return HOIG_OK;
}


// ---- column sums of a 32-row x 16-column block held one row per lane (the epilogue's register layout) on the legacy
// warp-level tensor-core path.  Each lane passes its own row's packed 16-bit pairs as the B fragment of m16n8k16
// (k slots (2t, 2t+1, 2t+8, 2t+9) of column n = lane / 4 are the row's four values), and A is a 0/1 selection matrix that
// routes slot s of MMA j to output row 4*j + s, so D[m][n] = sum over the four rows 4n..4n+3 of column m.  Adding the two D
// values of a lane and reducing over the four lanes of a group leaves, on every lane, the totals of columns lane/4 and
// lane/4 + 8: 4 MMAs + 4 shuffles instead of a 31-shuffle transpose-reduce.
template <bool F16>
__device__ __forceinline__ void hmma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    if (F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// e[i] = (lane/4 == 2i ? 1 : 0, lane/4 == 2i+1 ? 1 : 0) as a packed 16-bit pair of the operand type
template <bool F16>
__device__ __forceinline__ void colsum_select(int lane, uint32_t (&e)[4])
{
    const uint32_t one = F16 ? 0x3C00u : 0x3F80u;
    const int g = lane >> 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) e[i] = (g == 2 * i ? one : 0u) | (g == 2 * i + 1 ? one << 16 : 0u);
}
// the four MMAs of colsum16 accumulating into a caller-held fragment (reduce later: lo = d0 + d1, hi = d2 + d3, then the two shuffles)
template <bool F16>
__device__ __forceinline__ void colsum16_acc(const uint32_t (&pk)[8], const uint32_t (&e)[4], float (&d)[4])
{
    hmma_16816<F16>(d, e[0], 0u, e[1], 0u, pk[0], pk[1]);
    hmma_16816<F16>(d, e[2], 0u, e[3], 0u, pk[2], pk[3]);
    hmma_16816<F16>(d, 0u, e[0], 0u, e[1], pk[4], pk[5]);
    hmma_16816<F16>(d, 0u, e[2], 0u, e[3], pk[6], pk[7]);
}
template <bool F16>
__device__ __forceinline__ void colsum16(const uint32_t (&pk)[8], const uint32_t (&e)[4], float &lo, float &hi)
{
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    hmma_16816<F16>(d, e[0], 0u, e[1], 0u, pk[0], pk[1]);   // columns 0-3   -> D rows 0-3
    hmma_16816<F16>(d, e[2], 0u, e[3], 0u, pk[2], pk[3]);   // columns 4-7   -> D rows 4-7
    hmma_16816<F16>(d, 0u, e[0], 0u, e[1], pk[4], pk[5]);   // columns 8-11  -> D rows 8-11
    hmma_16816<F16>(d, 0u, e[2], 0u, e[3], pk[6], pk[7]);   // columns 12-15 -> D rows 12-15
    lo = d[0] + d[1];
    hi = d[2] + d[3];
    lo += __shfl_xor_sync(0xffffffffu, lo, 1); hi += __shfl_xor_sync(0xffffffffu, hi, 1);
    lo += __shfl_xor_sync(0xffffffffu, lo, 2); hi += __shfl_xor_sync(0xffffffffu, hi, 2);
}

}  // namespace
}  // namespace hoig
