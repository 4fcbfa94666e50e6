// common.cuh -- shared helpers for the hoig_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hoig_b200.h"

namespace hoig {

// thread-local last-error text, surfaced through hoig_last_error()
void set_error(const char *fmt, ...);
int check_launch(const char *what);  // cudaGetLastError -> status

#define HOIG_REQUIRE(cond, ...)                \
    do {                                       \
        if (!(cond)) {                         \
            hoig::set_error(__VA_ARGS__);      \
            return HOIG_ERR_INVALID;           \
        }                                      \
    } while (0)

// One process may drive several GPUs (the reference nn.Module works on whatever device its tensors are on): everything cached per
// process is keyed by the current device.
constexpr int kMaxDevices = 64;
enum DeviceSlot { SLOT_CONV_UMMA_BF16_1, SLOT_CONV_UMMA_BF16_1P, SLOT_CONV_UMMA_BF16_2, SLOT_CONV_UMMA_BF16_2P, SLOT_CONV_UMMA_F16_1,
                  SLOT_CONV_UMMA_F16_1P, SLOT_CONV_UMMA_F16_2, SLOT_CONV_UMMA_F16_2P,
                  SLOT_CONV_UMMA_FAST_FIRST, SLOT_CONV_UMMA_FAST_LAST = SLOT_CONV_UMMA_FAST_FIRST + 7,   // the same eight with the streamlined epilogue
                  SLOT_CONV_HALO, SLOT_RASTERIZE, SLOT_ATTN_TC_F16, SLOT_ATTN_TC_BF16,
                  SLOT_COUNT };
int device_sm_count();               // SM count of the CURRENT device
bool first_use_on_device(int slot);  // true exactly once per (current device, slot): set the kernel's function attributes then

static inline cudaStream_t as_stream(hoigStream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

template <typename T> struct DT;
template <> struct DT<float> {
    static __device__ __forceinline__ float ld(const float *p) { return *p; }
    static __device__ __forceinline__ void st(float *p, float v) { *p = v; }
};
template <> struct DT<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
};

template <> struct DT<__half> {
    static __device__ __forceinline__ float ld(const __half *p) { return __half2float(*p); }
    static __device__ __forceinline__ void st(__half *p, float v)
    {
        unsigned short r;
        asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
        *reinterpret_cast<unsigned short *>(p) = r;
    }
};

// two 16-bit storage values <-> two floats
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi)
{
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&t);
}
// fp16 storage saturates at +-65504 instead of overflowing to inf (one F2FP.SATFINITE either way): an out-of-range activation
// then stays finite through the following normalisation instead of turning a whole plane into NaN
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <typename T> __device__ __forceinline__ void unpack2(uint32_t w, float &lo, float &hi);
template <> __device__ __forceinline__ void unpack2<__nv_bfloat16>(uint32_t w, float &lo, float &hi)
{
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack2<__half>(uint32_t w, float &lo, float &hi)
{
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w));
    lo = f.x; hi = f.y;
}

// 8 consecutive channels (one 16-byte chunk for bf16, two for f32) <-> float[8]
__device__ __forceinline__ void load8(const float *p, float v[8])
{
    const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16 *p, float v[8])
{
    const uint4 r = *reinterpret_cast<const uint4 *>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void load8(const __half *p, float v[8])
{
    const uint4 r = *reinterpret_cast<const uint4 *>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack2<__half>(w[i], v[2 * i], v[2 * i + 1]);
}
__device__ __forceinline__ void store8(__half *p, const float v[8])
{
    uint4 r;
    r.x = pack2<__half>(v[0], v[1]); r.y = pack2<__half>(v[2], v[3]);
    r.z = pack2<__half>(v[4], v[5]); r.w = pack2<__half>(v[6], v[7]);
    *reinterpret_cast<uint4 *>(p) = r;
}
__device__ __forceinline__ void store8(float *p, const float v[8])
{
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4 *>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ void store8(__nv_bfloat16 *p, const float v[8])
{
    uint4 r;
    r.x = pack_bf16x2(v[0], v[1]); r.y = pack_bf16x2(v[2], v[3]);
    r.z = pack_bf16x2(v[4], v[5]); r.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4 *>(p) = r;
}
// value as it will be stored (bf16 rounding for the bf16 path, identity for f32)
template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_to<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
template <> __device__ __forceinline__ float round_to<__half>(float v) { return __half2float(__float2half_rn(v)); }

__device__ __forceinline__ float apply_act(float v, int act)
{
    switch (act) {
    case HOIG_ACT_RELU: return fmaxf(v, 0.f);
    case HOIG_ACT_LEAKY: return v > 0.f ? v : 0.01f * v;
    case HOIG_ACT_TANH: return tanhf(v);
    case HOIG_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
    }
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace hoig
