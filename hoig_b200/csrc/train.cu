// train.cu -- backward kernels of the fp32 training path: weight gradient of the implicit-GEMM
// convolution and the backward of InstanceNorm.  The data gradient of a convolution is itself a convolution and runs on the
// forward kernels (hoig_b200/autograd.py); BlockExtractor / LocalAttnReshape backward live in ops.cu.
//
// Reference behaviour being reproduced: torch.autograd of nn.Conv2d / nn.ConvTranspose2d / nn.InstanceNorm2d as used by
// models/networks/generator.py and models/networks/discriminator.py (the reference has no hand-written backward for them).
#include "common.cuh"

namespace hoig {
namespace {

constexpr int WG_TILE = 64;     // output channels x input channels per CTA
constexpr int WG_PIX = 32;      // pixels staged per iteration

// dW[co][r][s][ci] += sum over a pixel range of g[n][oy][ox][co] * x[n][oy*stride + r - pad_h][ox*stride + s - pad_w][ci]
// grid (co tiles * ci tiles, KH*KW, splits); 256 threads, each a 4 x 4 block of the 64 x 64 tile.
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ g, int64_t ldg,
                                                         float *__restrict__ dw, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                                                         int KH, int KW, int stride, int pad_h, int pad_w, int ci_tiles, int64_t pix_per_split)
{
    __shared__ __align__(16) float gs[WG_PIX][WG_TILE];
    __shared__ __align__(16) float xs[WG_PIX][WG_TILE];
    const int co0 = (blockIdx.x / ci_tiles) * WG_TILE, ci0 = (blockIdx.x % ci_tiles) * WG_TILE;
    const int r = blockIdx.y / KW, s = blockIdx.y % KW;
    const int64_t P = (int64_t)N * OH * OW;
    const int64_t p_begin = (int64_t)blockIdx.z * pix_per_split;
    const int64_t p_end = p_begin + pix_per_split < P ? p_begin + pix_per_split : P;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;     // tx: 4 input channels, ty: 4 output channels
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // loader mapping: 256 threads move 32 pixels x 64 channels as float4: thread -> (pixel = tid / 8, 4-channel groups tid % 8 and + 8)
    const int lp = threadIdx.x / 8, lc = (threadIdx.x % 8) * 4;
    for (int64_t p0 = p_begin; p0 < p_end; p0 += WG_PIX) {
        const int64_t p = p0 + lp;
        float4 gv[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        float4 xv[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        if (p < p_end) {
            const int n = (int)(p / ((int64_t)OH * OW));
            const int rem = (int)(p - (int64_t)n * OH * OW);
            const int oy = rem / OW, ox = rem - oy * OW;
            const float *gp = g + p * ldg;
            const int iy = oy * stride + r - pad_h, ix = ox * stride + s - pad_w;
            const bool inb = iy >= 0 && iy < H && ix >= 0 && ix < W;
            const float *xp = x + (((int64_t)n * H + iy) * W + ix) * ldx;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = lc + h * 32;
                if (co0 + c + 3 < Cout) gv[h] = *reinterpret_cast<const float4 *>(gp + co0 + c);
                else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int e = 0; e < 4; ++e) if (co0 + c + e < Cout) t[e] = gp[co0 + c + e];
                    gv[h] = make_float4(t[0], t[1], t[2], t[3]);
                }
                if (inb) {
                    if (ci0 + c + 3 < Cin) xv[h] = *reinterpret_cast<const float4 *>(xp + ci0 + c);
                    else {
                        float t[4] = {0.f, 0.f, 0.f, 0.f};
                        for (int e = 0; e < 4; ++e) if (ci0 + c + e < Cin) t[e] = xp[ci0 + c + e];
                        xv[h] = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
            }
        }
        __syncthreads();     // previous iteration's reads are done
        *reinterpret_cast<float4 *>(&gs[lp][lc]) = gv[0]; *reinterpret_cast<float4 *>(&gs[lp][lc + 32]) = gv[1];
        *reinterpret_cast<float4 *>(&xs[lp][lc]) = xv[0]; *reinterpret_cast<float4 *>(&xs[lp][lc + 32]) = xv[1];
        __syncthreads();
#pragma unroll 8
        for (int q = 0; q < WG_PIX; ++q) {
            const float4 a = *reinterpret_cast<const float4 *>(&gs[q][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&xs[q][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = ci0 + tx * 4 + j;
            if (ci < Cin) atomicAdd(&dw[(((int64_t)co * KH + r) * KW + s) * Cin + ci], acc[i][j]);
        }
    }
}

// pass 1 of the InstanceNorm backward: a[n][c] = sum_p gy, b[n][c] = sum_p gy * xhat   (grid (slabs, N), thread = (pixel lane, 8 channels))
__global__ void instnorm_bwd_reduce_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ gy, int64_t ldg,
                                           const double *__restrict__ stats, int HW, int C, int pix_per_block, float eps, double *__restrict__ ab)
{
    extern __shared__ float sm[];  // [lanes][C][2]
    const int chunks = C / 8;
    const int lanes = blockDim.x / chunks;
    const int cc = threadIdx.x % chunks, pl = threadIdx.x / chunks;
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
    float a[8], b[8], mean[8], rstd[8];
    const double inv_hw = 1.0 / (double)HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = b[j] = 0.f;
        const double s = stats[((int64_t)n * C + cc * 8 + j) * 2], q = stats[((int64_t)n * C + cc * 8 + j) * 2 + 1];
        const double m = s * inv_hw;
        double var = q * inv_hw - m * m;
        if (var < 0) var = 0;
        mean[j] = (float)m;
        rstd[j] = 1.0f / sqrtf((float)var + eps);
    }
    if (pl < lanes) {
        for (int p = p0 + pl; p < p1; p += lanes) {
            float xv[8], gv[8];
            load8(x + ((int64_t)n * HW + p) * ldx + cc * 8, xv);
            load8(gy + ((int64_t)n * HW + p) * ldg + cc * 8, gv);
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] += gv[j]; b[j] = fmaf(gv[j], (xv[j] - mean[j]) * rstd[j], b[j]); }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sm[((size_t)pl * C + cc * 8 + j) * 2] = a[j];
            sm[((size_t)pl * C + cc * 8 + j) * 2 + 1] = b[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += sm[(size_t)l * C * 2 + i];
        atomicAdd(&ab[(int64_t)n * C * 2 + i], (double)t);
    }
}

// pass 2: dx = gamma * rstd * (gy - a / HW - xhat * b / HW);  dgamma[c] += sum_n b, dbeta[c] += sum_n a (first slab of each image)
__global__ void instnorm_bwd_apply_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ gy, int64_t ldg,
                                          const double *__restrict__ stats, const double *__restrict__ ab, const float *__restrict__ gamma,
                                          float *__restrict__ dx, int64_t lddx, float *__restrict__ dgamma, float *__restrict__ dbeta,
                                          int HW, int C, int pix_per_block, float eps)
{
    extern __shared__ float sm[];  // mean[C] | rstd[C] | k1[C] (= a / HW) | k2[C] (= b / HW) | scale[C] (= gamma * rstd)
    float *meanv = sm, *rstdv = sm + C, *k1 = sm + 2 * C, *k2 = sm + 3 * C, *scale = sm + 4 * C;
    const int n = blockIdx.y;
    const double inv_hw = 1.0 / (double)HW;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double s = stats[((int64_t)n * C + c) * 2], q = stats[((int64_t)n * C + c) * 2 + 1];
        const double m = s * inv_hw;
        double var = q * inv_hw - m * m;
        if (var < 0) var = 0;
        const float r = 1.0f / sqrtf((float)var + eps);
        const double a = ab[((int64_t)n * C + c) * 2], b = ab[((int64_t)n * C + c) * 2 + 1];
        meanv[c] = (float)m; rstdv[c] = r;
        k1[c] = (float)(a * inv_hw); k2[c] = (float)(b * inv_hw);
        scale[c] = (gamma ? gamma[c] : 1.f) * r;
        if (blockIdx.x == 0 && dgamma) { atomicAdd(&dgamma[c], (float)b); atomicAdd(&dbeta[c], (float)a); }
    }
    __syncthreads();
    const int chunks = C / 8;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
    for (int64_t i = threadIdx.x; i < (int64_t)(p1 - p0) * chunks; i += blockDim.x) {
        const int p = p0 + (int)(i / chunks), c0 = (int)(i % chunks) * 8;
        float xv[8], gv[8], o[8];
        load8(x + ((int64_t)n * HW + p) * ldx + c0, xv);
        load8(gy + ((int64_t)n * HW + p) * ldg + c0, gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = (xv[j] - meanv[c0 + j]) * rstdv[c0 + j];
            o[j] = scale[c0 + j] * (gv[j] - k1[c0 + j] - xh * k2[c0 + j]);
        }
        store8(dx + ((int64_t)n * HW + p) * lddx + c0, o);
    }
}

}  // namespace
}  // namespace hoig

using namespace hoig;

extern "C" int hoig_conv2d_wgrad_f32(const float *x, int64_t ldx, const float *g, int64_t ldg, float *dw, int N, int H, int W, int Cin,
                                     int OH, int OW, int Cout, int KH, int KW, int stride, int pad_h, int pad_w, hoigStream_t stream)
{
    HOIG_REQUIRE(x && g && dw, "conv2d_wgrad: null pointer");
    HOIG_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && OH > 0 && OW > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && KH * KW <= 65535,
                 "conv2d_wgrad: bad shape");
    HOIG_REQUIRE(ldx >= Cin && ldg >= Cout && ldx % 4 == 0 && ldg % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)g % 16) == 0,
                 "conv2d_wgrad: pixel strides must be multiples of 4 floats and the tensors 16-byte aligned");
    const int co_tiles = ceil_div(Cout, WG_TILE), ci_tiles = ceil_div(Cin, WG_TILE);
    const int64_t P = (int64_t)N * OH * OW;
    const int64_t tiles = (int64_t)co_tiles * ci_tiles * KH * KW;
    int64_t splits = ceil_div(4 * (int64_t)device_sm_count(), tiles);
    const int64_t max_splits = P / 1024 > 0 ? P / 1024 : 1;
    if (splits > max_splits) splits = max_splits;
    if (splits > 65535) splits = 65535;
    if (splits < 1) splits = 1;
    int64_t pps = (P + splits - 1) / splits;
    pps = (pps + WG_PIX - 1) / WG_PIX * WG_PIX;
    splits = (P + pps - 1) / pps;
    dim3 grid((unsigned)(co_tiles * ci_tiles), (unsigned)(KH * KW), (unsigned)splits);
    conv_wgrad_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, ldx, g, ldg, dw, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad_h, pad_w,
                                                           ci_tiles, pps);
    return check_launch("conv_wgrad_kernel");
}

extern "C" int hoig_instnorm_backward_f32(const float *x, int64_t ldx, const float *gy, int64_t ldg, const double *stats, const float *gamma,
                                          float *dx, int64_t lddx, double *scratch, float *dgamma, float *dbeta, int N, int HW, int C,
                                          float eps, hoigStream_t stream)
{
    HOIG_REQUIRE(x && gy && stats && dx && scratch, "instnorm_backward: null pointer");
    HOIG_REQUIRE(C % 8 == 0 && C / 8 <= 256 && ldx % 8 == 0 && ldg % 8 == 0 && lddx % 8 == 0 && ldx >= C && ldg >= C && lddx >= C,
                 "instnorm_backward: channels / strides must be multiples of 8 (C=%d)", C);
    HOIG_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "instnorm_backward: dgamma and dbeta go together");
    if (N == 0 || HW == 0) return HOIG_OK;
    const int tpb = 256;
    int pp = 1024;
    while (pp > 64 && (int64_t)ceil_div(HW, pp) * N < 2 * device_sm_count()) pp /= 2;
    const int lanes = tpb / (C / 8);
    dim3 grid(ceil_div(HW, pp), N);
    instnorm_bwd_reduce_kernel<<<grid, tpb, (size_t)lanes * C * 2 * sizeof(float), as_stream(stream)>>>(x, ldx, gy, ldg, stats, HW, C, pp, eps, scratch);
    int st = check_launch("instnorm_bwd_reduce_kernel");
    if (st != HOIG_OK) return st;
    instnorm_bwd_apply_kernel<<<grid, tpb, 5 * (size_t)C * sizeof(float), as_stream(stream)>>>(x, ldx, gy, ldg, stats, scratch, gamma, dx, lddx, dgamma,
                                                                                              dbeta, HW, C, pp, eps);
    return check_launch("instnorm_bwd_apply_kernel");
}
