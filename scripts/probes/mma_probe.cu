// mma_probe.cu -- how long does one tcgen05.mma (128 x N x 16, kind::f16, operands in smem) take as a function of N,
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_probe mma_probe.cu -lcuda
// the number of accumulators the issue stream rotates over, and the commit cadence?  One CTA per SM, no TMA traffic.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int N, int naccs, int per_commit, int iters, int kstep_mode, int extra, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t bar, bar2, bar3, bar4[4];
    __shared__ uint32_t tmem_base_s;
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)));
        for (int w = 0; w < 4; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar4[w])));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar3)) : "memory");   // phase 0 of bar3 is complete
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const int nissue = extra < 1 ? 1 : extra;
    if (threadIdx.x < 32 * nissue) {
        const int wid = threadIdx.x / 32;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a = smem_u32(smem), b = a + 16384;
        uint32_t ncommit = 0;
        const long long t0 = clock64();
        const uint64_t da = umma_desc(a), db = umma_desc(b);
        const uint32_t amask = (uint32_t)naccs - 1, bar_a = smem_u32(&bar);
        const int shift = (kstep_mode & 2) ? 0 : 2;
#pragma unroll 1
        for (int i = 0; i < iters; i += 4) {
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t acc_i = ((uint32_t)(i + j) >> shift) & amask;
                    umma(tmem + (acc_i + (uint32_t)wid * (uint32_t)naccs) * (uint32_t)N, da + (uint64_t)(kstep_mode ? j * 2 : 0), db + (uint64_t)(kstep_mode ? j * 2 : 0), idesc, 1u);
                }
                if (per_commit) { commit(bar_a); ++ncommit; }
            }
            __syncwarp();

        }
        const long long t1 = clock64();
        if (elect_one()) commit(smem_u32(&bar4[wid]));      // fires when every MMA above has retired
        __syncwarp();
        mbar_wait(smem_u32(&bar4[wid]), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main()
{
    long long *out;
    cudaMalloc(&out, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 4096;
    printf("grid N naccs per_commit issuing warps (each its own accumulator; cycles are per MMA of ONE warp) | issue cyc/MMA | total cyc/MMA | floor(N/2)\n");
    for (int grid : {1, 148})
        for (int N : {64, 128, 256})
            for (int naccs : {1})
                for (int pc : {4})
                    for (int ks : {1, 2, 4}) {     // ks = number of issuing warps here
                        if (naccs * ks * N > 512 || grid == 1) continue;
                        probe<<<grid, 128, 50 * 1024 + 1024>>>(N, naccs, pc, iters, 1, ks, out);
                        long long h[2];
                        cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                        printf("%4d %4d %2d %2d %d | %8.1f | %8.1f | %d\n", grid, N, naccs, pc, ks, (double)h[0] / iters, (double)h[1] / iters, N / 2);
                    }
    return 0;
}
