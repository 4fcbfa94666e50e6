// ops.cu -- bandwidth-bound stages of the generator forward and the two
// reference op boundaries (BlockExtractor / LocalAttnReshape forward).
// All kernels move 8-channel (16/32-byte) chunks of NHWC rows so that warps
// issue fully coalesced 128-bit accesses.
#include <stdlib.h>

#include <type_traits>

#include "conv_common.cuh"

namespace hoig {
namespace {

constexpr int TPB = 256;

// ---------------------------------------------------------------- layout glue
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float *__restrict__ src, int B, int C, int HW, T *__restrict__ dst, int64_t ldd, int Cpad)
{
    const int chunks = Cpad / 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HW * chunks) return;
    const int ch = (int)(i % chunks);
    const int64_t bp = i / chunks;
    const int b = (int)(bp / HW), p = (int)(bp % HW);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = ch * 8 + j;
        v[j] = c < C ? src[((int64_t)b * C + c) * HW + p] : 0.f;
    }
    store8(dst + bp * ldd + ch * 8, v);
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T *__restrict__ src, int64_t lds, int B, int C, int HW, float *__restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * C * HW) return;
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const int b = (int)(i / ((int64_t)HW * C));
    dst[i] = DT<T>::ld(src + ((int64_t)b * HW + p) * lds + c);
}

// spade.py:30, legacy 'nearest': src = min(floor(dst * in/out), in-1)
template <typename T>
__global__ void seg_resize_kernel(const float *__restrict__ seg, int B, int C, int Hi, int Wi, T *__restrict__ dst,
                                  int64_t ldd, int Cpad, int Ho, int Wo)
{
    const int chunks = Cpad / 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Ho * Wo * chunks) return;
    const int ch = (int)(i % chunks);
    const int64_t bp = i / chunks;
    const int x = (int)(bp % Wo), y = (int)((bp / Wo) % Ho), b = (int)(bp / ((int64_t)Wo * Ho));
    const float sy = (float)Hi / (float)Ho, sx = (float)Wi / (float)Wo;
    const int iy = min((int)floorf(y * sy), Hi - 1), ix = min((int)floorf(x * sx), Wi - 1);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = ch * 8 + j;
        v[j] = c < C ? seg[(((int64_t)b * C + c) * Hi + iy) * Wi + ix] : 0.f;
    }
    store8(dst + bp * ldd + ch * 8, v);
}

// spade.py:30-31 resize + the 3x3 im2col of mlp_shared's input in one pass: dst[b,y,x, t*C + c] = nearest-resized
// seg[b,c] at (y + t/3 - 1, x + t%3 - 1), zero outside the image (the conv's zero padding) and for channels >= 9*C.
// The segmentation map has 3-12 channels; unfolding it once per resolution turns every mlp_shared conv of that resolution
// into a 1x1 GEMM with K = 64 or 128 that TMA can feed (a 16-byte-chunk gather over 8/16 channels cannot).
template <typename T>
__global__ void seg_unfold3_kernel(const float *__restrict__ seg, int B, int C, int Hi, int Wi, T *__restrict__ dst,
                                   int64_t ldd, int Kpad, int Ho, int Wo)
{
    const int chunks = Kpad / 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Ho * Wo * chunks) return;
    const int ch = (int)(i % chunks);
    const int64_t bp = i / chunks;
    const int x = (int)(bp % Wo), y = (int)((bp / Wo) % Ho), b = (int)(bp / ((int64_t)Wo * Ho));
    const float sy = (float)Hi / (float)Ho, sx = (float)Wi / (float)Wo;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int kc = ch * 8 + j, t = kc / C, c = kc - t * C;
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        float val = 0.f;
        if (t < 9 && yy >= 0 && yy < Ho && xx >= 0 && xx < Wo) {
            const int iy = min((int)floorf(yy * sy), Hi - 1), ix = min((int)floorf(xx * sx), Wi - 1);
            val = seg[(((int64_t)b * C + c) * Hi + iy) * Wi + ix];
        }
        v[j] = val;
    }
    store8(dst + bp * ldd + ch * 8, v);
}

// Same result, staged: a CTA takes PX consecutive pixels of one output row, gathers the three resized source rows (plus one
// halo pixel each side) into shared memory once, and writes every pixel's Kpad channels as consecutive 16-byte chunks.
// A thread keeps ONE 8-channel chunk for all its pixels, so the (tap, channel) decode of its 8 columns happens once.
template <typename T, int PX>
__global__ void __launch_bounds__(256) seg_unfold3_row_kernel(const float *__restrict__ seg, int C, int Hi, int Wi, T *__restrict__ dst,
                                                              int64_t ldd, int Kpad, int Ho, int Wo)
{
    constexpr int MAXC = 16, TW = PX + 2;
    __shared__ float tile[3 * TW * (MAXC + 1)];     // [r][j][c], c padded to MAXC + 1
    const int segs = Wo / PX;
    const int x0 = (blockIdx.x % segs) * PX, y = (blockIdx.x / segs) % Ho, b = blockIdx.x / (segs * Ho);
    const float sy = (float)Hi / (float)Ho, sx = (float)Wi / (float)Wo;
    for (int i = threadIdx.x; i < 3 * TW * C; i += blockDim.x) {
        const int j = i % TW, rc = i / TW, c = rc % C, r = rc / C;   // consecutive lanes walk along x
        const int yy = y + r - 1, xx = x0 + j - 1;
        float val = 0.f;
        if (yy >= 0 && yy < Ho && xx >= 0 && xx < Wo) {
            const int iy = min((int)floorf(yy * sy), Hi - 1), ix = min((int)floorf(xx * sx), Wi - 1);
            val = __ldg(seg + (((int64_t)b * C + c) * Hi + iy) * Wi + ix);
        }
        tile[(r * TW + j) * (MAXC + 1) + c] = val;
    }
    const int chunks = Kpad / 8;                     // a power of two <= 32 for the shipped shapes; any divisor of 256 works
    const int ch = threadIdx.x % chunks, plane = threadIdx.x / chunks, lanes = blockDim.x / chunks;
    int off[8];                                      // smem offset of column j of this chunk at pixel 0, -1: zero padding
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int kc = ch * 8 + j, t = kc / C, c = kc - t * C;
        off[j] = t < 9 ? ((t / 3) * TW + t % 3) * (MAXC + 1) + c : -1;
    }
    __syncthreads();
    T *drow = dst + (((int64_t)b * Ho + y) * Wo + x0) * ldd + ch * 8;
    for (int px = plane; px < PX; px += lanes) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = off[j] >= 0 ? tile[off[j] + px * (MAXC + 1)] : 0.f;
        store8(drow + (int64_t)px * ldd, v);
    }
}

// ------------------------------------------------------------- instance norm
// grid (slabs, N); thread = (pixel lane, channel chunk)
template <typename T>
__global__ void plane_stats_kernel(const T *__restrict__ x, int64_t ldx, int HW, int C, int pix_per_block, double *__restrict__ stats)
{
    extern __shared__ float sm[];  // [lanes][C][2]
    const int chunks = C / 8;
    const int lanes = blockDim.x / chunks;
    const int cc = threadIdx.x % chunks, pl = threadIdx.x / chunks;
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    if (pl < lanes) {
        for (int p = p0 + pl; p < p1; p += lanes) {
            float v[8];
            load8(x + ((int64_t)n * HW + p) * ldx + cc * 8, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sm[((size_t)pl * C + cc * 8 + j) * 2] = s[j];
            sm[((size_t)pl * C + cc * 8 + j) * 2 + 1] = q[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
        float a = 0.f;
        for (int l = 0; l < lanes; ++l) a += sm[(size_t)l * C * 2 + i];
        atomicAdd(&stats[(int64_t)n * C * 2 + i], (double)a);
    }
}

// grid (slabs, N).  mean / rstd of the C planes of image n are rebuilt in smem per CTA.
template <typename T>
__global__ void __launch_bounds__(256, 4) instnorm_apply_kernel(const T *__restrict__ x, int64_t ldx, const double *__restrict__ stats,
                                      const float *__restrict__ gamma, const float *__restrict__ beta,
                                      const T *__restrict__ gb, int64_t ldgb, const T *__restrict__ res, int64_t ldr,
                                      int relu, T *__restrict__ dst, int64_t ldd, int HW, int C, int pix_per_block, float eps)
{
    extern __shared__ float sm[];  // mean[C] | scale[C] | shift[C]
    float *meanv = sm, *scale = sm + C, *shift = sm + 2 * C;
    const int n = blockIdx.y;
    const double inv_hw = 1.0 / (double)HW;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        // the cancellation-prone part (E[x^2] - mean^2) stays in double; 1/sqrt runs in fp32 like the reference's
        // instance norm (a double sqrt + divide per channel per CTA used to cost as much as the CTA's pixels)
        const double2 sq = *reinterpret_cast<const double2 *>(stats + ((int64_t)n * C + c) * 2);
        const double mean = sq.x * inv_hw;
        double var = sq.y * inv_hw - mean * mean;
        if (var < 0) var = 0;
        const float rstd = 1.0f / sqrtf((float)var + eps);
        const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
        meanv[c] = (float)mean;
        scale[c] = rstd * g;
        shift[c] = b;
    }
    __syncthreads();
    const int chunks = C / 8;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
    if (blockDim.x % chunks == 0) {
        // fast path: a thread keeps ONE 8-channel chunk for all its pixels, so the per-channel constants live in
        // registers and the loop has no index division and no shared-memory traffic
        const int lanes = blockDim.x / chunks;
        const int c8 = (threadIdx.x % chunks) * 8, pl = threadIdx.x / chunks;
        float mu[8], sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { mu[j] = meanv[c8 + j]; sc[j] = scale[c8 + j]; sh[j] = shift[c8 + j]; }
        constexpr int U = 2;
        for (int pb = p0 + pl; pb < p1; pb += lanes * U) {
            float v[U][8];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int p = min(pb + u * lanes, p1 - 1);
                load8(x + ((int64_t)n * HW + p) * ldx + c8, v[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int p = pb + u * lanes;
                if (p >= p1) break;
                const int64_t row = (int64_t)n * HW + p;
#pragma unroll
                for (int j = 0; j < 8; ++j) v[u][j] = fmaf(v[u][j] - mu[j], sc[j], sh[j]);
                if (gb) {  // spade.py:36  normalized * (1 + gamma) + beta
                    float g[8], b[8];
                    load8(gb + row * ldgb + c8, g);
                    load8(gb + row * ldgb + C + c8, b);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[u][j] = fmaf(v[u][j], 1.f + g[j], b[j]);
                }
                if (res) {
                    float r[8];
                    load8(res + row * ldr + c8, r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[u][j] += r[j];
                }
                if (relu) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[u][j] = fmaxf(v[u][j], 0.f);
                }
                store8(dst + row * ldd + c8, v[u]);
            }
        }
        return;
    }
    for (int i = threadIdx.x; i < (p1 - p0) * chunks; i += blockDim.x) {
        const int cc = i % chunks, p = p0 + i / chunks;
        const int64_t row = (int64_t)n * HW + p;
        float v[8];
        load8(x + row * ldx + cc * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j] - meanv[cc * 8 + j], scale[cc * 8 + j], shift[cc * 8 + j]);
        if (gb) {
            float g[8], b[8];
            load8(gb + row * ldgb + cc * 8, g);
            load8(gb + row * ldgb + C + cc * 8, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], 1.f + g[j], b[j]);
        }
        if (res) {
            float r[8];
            load8(res + row * ldr + cc * 8, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += r[j];
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        store8(dst + row * ldd + cc * 8, v);
    }
}

// ------------------------------------------------------------------- warping
// generator.py:466-473 (upsample_bilinear2d, align_corners=True) + generator.py:484-488.
__global__ void resize_flow_kernel(const float *__restrict__ T, int B, int Hi, int Wi, int h, int sub_idt, float *__restrict__ flow)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * h * h) return;
    const int x = (int)(i % h), y = (int)((i / h) % h), b = (int)(i / ((int64_t)h * h));
    const float rh = h > 1 ? (float)(Hi - 1) / (float)(h - 1) : 0.f;
    const float rw = h > 1 ? (float)(Wi - 1) / (float)(h - 1) : 0.f;
    const float h1r = rh * y, w1r = rw * x;
    const int h1 = (int)h1r, w1 = (int)w1r;
    const int h1p = h1 < Hi - 1 ? 1 : 0, w1p = w1 < Wi - 1 ? 1 : 0;
    const float h1l = h1r - h1, h0l = 1.f - h1l, w1l = w1r - w1, w0l = 1.f - w1l;
    const float *t = T + (int64_t)b * Hi * Wi * 2;
    // 'ij' identity grid: channel 0 follows the ROW index, channel 1 the COLUMN index (quirk Q2)
    const float idt0 = (float)(-1.0 + y * (2.0 / h)), idt1 = (float)(-1.0 + x * (2.0 / h));
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const float v00 = t[((int64_t)h1 * Wi + w1) * 2 + c], v01 = t[((int64_t)h1 * Wi + w1 + w1p) * 2 + c];
        const float v10 = t[((int64_t)(h1 + h1p) * Wi + w1) * 2 + c], v11 = t[((int64_t)(h1 + h1p) * Wi + w1 + w1p) * 2 + c];
        const float val = h0l * (w0l * v00 + w1l * v01) + h1l * (w0l * v10 + w1l * v11);
        flow[i * 2 + c] = sub_idt ? val - (c == 0 ? idt0 : idt1) : val;
    }
}

// extract_attn.py:24-28 tail.  One warp per pixel: conv1x1 -> softmax -> weighted average of the
// 25 bilinear taps of the source, plus the residual add of generator.py:407/427/446.
template <typename T, int KK>
__global__ void __launch_bounds__(128)
attn_finish_kernel(const T *__restrict__ hidden, int64_t ldh, int Chid, const float *__restrict__ w2,
                   const float *__restrict__ b2, const T *__restrict__ src, int64_t lds, const float *__restrict__ flow,
                   const T *__restrict__ tgt, int64_t ldt, T *__restrict__ dst, int64_t ldd, int64_t npix_total, int h, int C,
                   const T *__restrict__ unfold, int64_t ldu)
{
    constexpr int K = (KK == 25) ? 5 : 3;
    __shared__ float s_attn[4][KK];
    __shared__ int s_idx[4][KK][4];
    __shared__ float s_w[4][KK][4];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int64_t pix = (int64_t)blockIdx.x * 4 + warp;
    if (pix >= npix_total) return;  // whole warp exits together; no block-level sync below
    const int x = (int)(pix % h), y = (int)((pix / h) % h);
    const int64_t plane = (pix / ((int64_t)h * h)) * h * h;

    // logits: lane-strided partial dot products, butterfly-reduced
    float logit = -INFINITY;
    {
        float part[KK];
#pragma unroll
        for (int t = 0; t < KK; ++t) part[t] = 0.f;
        for (int c = lane; c < Chid; c += 32) {
            const float hv = DT<T>::ld(hidden + pix * ldh + c);
#pragma unroll
            for (int t = 0; t < KK; ++t) part[t] = fmaf(hv, __ldg(w2 + t * Chid + c), part[t]);
        }
#pragma unroll
        for (int t = 0; t < KK; ++t) {
            const float tot = warp_sum(part[t]);
            if (lane == t) logit = tot + b2[t];
        }
    }
    float mx = logit;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e = lane < KK ? expf(logit - mx) : 0.f;
    const float den = warp_sum(e);
    if (lane < KK) {
        s_attn[warp][lane] = e / den;
        const BETap tp = be_tap(flow[pix * 2], flow[pix * 2 + 1], y, x, lane / K, lane % K, K, h, h);
#pragma unroll
        for (int q = 0; q < 4; ++q) { s_idx[warp][lane][q] = tp.idx[q]; s_w[warp][lane][q] = tp.w[q]; }
    }
    __syncwarp();
    for (int cc = lane; cc < C / 8; cc += 32) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int t = 0; t < KK; ++t) {
            float smp[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) smp[j] = 0.f;
            if (unfold) {   // taps already extracted by attn_unfold: [tap][tgt C | src C] per pixel
                load8(unfold + pix * ldu + (int64_t)t * 2 * C + C + cc * 8, smp);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float u[8];
                    load8(src + (plane + s_idx[warp][t][q]) * lds + cc * 8, u);
                    const float wq = s_w[warp][t][q];
#pragma unroll
                    for (int j = 0; j < 8; ++j) smp[j] = __fmaf_rn(wq, u[j], smp[j]);
                }
            }
            const float a = s_attn[warp][t];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(a, smp[j], acc[j]);
        }
        float tv[8];
        load8(tgt + pix * ldt + cc * 8, tv);
#pragma unroll
        for (int j = 0; j < 8; ++j) tv[j] += acc[j] * (1.0f / KK);
        store8(dst + pix * ldd + cc * 8, tv);
    }
}

// generator.py:475-478: F.grid_sample bilinear / zeros / align_corners=False
template <typename T>
__global__ void grid_sample_kernel(const T *__restrict__ x, int64_t ldx, const float *__restrict__ grid,
                                   const T *__restrict__ tgt, int64_t ldt, T *__restrict__ dst, int64_t ldd,
                                   int64_t npix_total, int h, int C)
{
    const int chunks = C / 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix_total * chunks) return;
    const int cc = (int)(i % chunks);
    const int64_t pix = i / chunks;
    const int64_t plane = (pix / ((int64_t)h * h)) * h * h;
    const float gx = grid[pix * 2], gy = grid[pix * 2 + 1];
    const float ix = ((gx + 1.f) * h - 1.f) * 0.5f, iy = ((gy + 1.f) * h - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int xs[2] = {x0, x0 + 1}, ys[2] = {y0, y0 + 1};
    const float wxs[2] = {wx0, wx1}, wys[2] = {wy0, wy1};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            if (ys[a] < 0 || ys[a] >= h || xs[b] < 0 || xs[b] >= h) continue;
            float u[8];
            load8(x + (plane + (int64_t)ys[a] * h + xs[b]) * ldx + cc * 8, u);
            const float w = wys[a] * wxs[b];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, u[j], acc[j]);
        }
    if (tgt) {
        float tv[8];
        load8(tgt + pix * ldt + cc * 8, tv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += tv[j];
    }
    store8(dst + pix * ldd + cc * 8, acc);
}

__global__ void composite_kernel(const float *__restrict__ bg, const float *__restrict__ obj, const float *__restrict__ hand,
                                 const float *__restrict__ mbg, const float *__restrict__ mhand, float *__restrict__ out, int B, int HW)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * 3 * HW) return;
    const int p = (int)(i % HW);
    const int b = (int)(i / ((int64_t)3 * HW));
    const float mb = mbg[(int64_t)b * HW + p], mh = mhand[(int64_t)b * HW + p];
    out[i] = mb * bg[i] + (1.f - mb) * (obj[i] * mh + hand[i] * (1.f - mh));
}

// ------------------------------------------ reference op boundary (NCHW f32)
// block_extractor_kernel.cu:21-85, one thread per output element.
__global__ void block_extract_kernel(const float *__restrict__ src, const float *__restrict__ flow, float *__restrict__ out,
                                     int64_t n, int C, int Hs, int Ws, int Hf, int Wf, int k)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int Wo = k * Wf, Ho = k * Hf;
    const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho);
    const int c = (int)((i / ((int64_t)Wo * Ho)) % C), b = (int)(i / ((int64_t)Wo * Ho * C));
    const int yf = y / k, xf = x / k;
    const float fx = flow[(((int64_t)b * 2 + 0) * Hf + yf) * Wf + xf];
    const float fy = flow[(((int64_t)b * 2 + 1) * Hf + yf) * Wf + xf];
    const BETap t = be_tap(fx, fy, yf, xf, y % k, x % k, k, Hs, Ws);
    const float *s = src + ((int64_t)b * C + c) * Hs * Ws;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc = __fmaf_rn(t.w[q], s[t.idx[q]], acc);
    out[i] = acc;
}

// local_attn_reshape_kernel.cu:21-61
__global__ void local_attn_reshape_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t n, int k, int H, int W)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int Wo = k * W, Ho = k * H;
    const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho), b = (int)(i / ((int64_t)Wo * Ho));
    out[i] = in[(((int64_t)b * k * k + (y % k) * k + x % k) * H + y / k) * W + x / k];
}

// ---------------------------------------------------------------- backward of the two reference ops (boundary B2, row N3)
// block_extractor_kernel.cu:86-166.  Source gradient: one thread per output element scatters to its four taps (atomicAdd, as
// the reference).  Flow gradient: the reference issues 2 atomics per output element onto the same (b, yf, xf) cell, C*k*k
// colliding updates per cell; here one warp owns a cell, accumulates over channels and taps in registers and writes once.
__global__ void block_extract_bwd_src_kernel(const float *__restrict__ flow, const float *__restrict__ gout, float *__restrict__ gsrc,
                                             int64_t n, int C, int Hs, int Ws, int Hf, int Wf, int k)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int Wo = k * Wf, Ho = k * Hf;
    const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho);
    const int c = (int)((i / ((int64_t)Wo * Ho)) % C), b = (int)(i / ((int64_t)Wo * Ho * C));
    const int yf = y / k, xf = x / k;
    const float fx = flow[(((int64_t)b * 2 + 0) * Hf + yf) * Wf + xf];
    const float fy = flow[(((int64_t)b * 2 + 1) * Hf + yf) * Wf + xf];
    const BETap t = be_tap(fx, fy, yf, xf, y % k, x % k, k, Hs, Ws);
    float *gs = gsrc + ((int64_t)b * C + c) * Hs * Ws;
    const float g = gout[i];
#pragma unroll
    for (int q = 0; q < 4; ++q) atomicAdd(gs + t.idx[q], g * t.w[q]);   // w = xP * yP, :152-155
}

__global__ void __launch_bounds__(128)
block_extract_bwd_flow_kernel(const float *__restrict__ src, const float *__restrict__ flow, const float *__restrict__ gout,
                              float *__restrict__ gflow, int64_t ncells, int C, int Hs, int Ws, int Hf, int Wf, int k)
{
    const int lane = threadIdx.x % 32;
    const int64_t cell = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    if (cell >= ncells) return;                       // whole warp exits together
    const int xf = (int)(cell % Wf), yf = (int)((cell / Wf) % Hf), b = (int)(cell / ((int64_t)Wf * Hf));
    const int Wo = k * Wf, Ho = k * Hf;
    const float fx = flow[(((int64_t)b * 2 + 0) * Hf + yf) * Wf + xf];
    const float fy = flow[(((int64_t)b * 2 + 1) * Hf + yf) * Wf + xf];
    float gx = 0.f, gy = 0.f;
    for (int j = lane; j < C * k * k; j += 32) {
        const int c = j / (k * k), tap = j % (k * k), ky = tap / k, kx = tap % k;
        const float dy = __fadd_rn(__fadd_rn(fy, (float)(ky - k / 2)), (float)yf), dx = __fadd_rn(__fadd_rn(fx, (float)(kx - k / 2)), (float)xf);
        const float fdx = floorf(dx), fdy = floorf(dy);
        const int xL = max(min((int)fdx, Ws - 1), 0), xR = max(min((int)(fdx + 1.f), Ws - 1), 0);
        const int yT = max(min((int)fdy, Hs - 1), 0), yB = max(min((int)(fdy + 1.f), Hs - 1), 0);
        const float xLp = 1.f - (dx - fdx), xRp = dx - fdx, yTp = 1.f - (dy - fdy), yBp = dy - fdy;
        const float *s = src + ((int64_t)b * C + c) * Hs * Ws;
        const float vLT = s[yT * Ws + xL], vRT = s[yT * Ws + xR], vLB = s[yB * Ws + xL], vRB = s[yB * Ws + xR];
        const float g = gout[(((int64_t)b * C + c) * Ho + yf * k + ky) * Wo + xf * k + kx];
        gy += g * (-xLp * vLT - xRp * vRT + xLp * vLB + xRp * vRB);       // :157
        gx += g * (-yTp * vLT - yBp * vLB + yTp * vRT + yBp * vRB);       // :158
    }
    gx = warp_sum(gx);
    gy = warp_sum(gy);
    if (lane == 0) {
        gflow[(((int64_t)b * 2 + 0) * Hf + yf) * Wf + xf] += gx;
        gflow[(((int64_t)b * 2 + 1) * Hf + yf) * Wf + xf] += gy;
    }
}

// local_attn_reshape_kernel.cu:62-104: the forward is a permutation, so is the backward (no atomics needed)
__global__ void local_attn_reshape_bwd_kernel(const float *__restrict__ gout, float *__restrict__ gin, int64_t n, int k, int H, int W)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int Wo = k * W, Ho = k * H;
    const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho), b = (int)(i / ((int64_t)Wo * Ho));
    gin[(((int64_t)b * k * k + (y % k) * k + x % k) * H + y / k) * W + x / k] += gout[i];
}

// extract_attn.py:24-25 materialised for the tensor-core path: U[pix][t*2C + c] = BlockExtractor(tgt, 0) tap t
// for c < C and BlockExtractor(src, flow) tap t for c >= C (block_extractor_kernel.cu:52-84, same float op
// order as the fused gather), so the k x k stride-k conv over cat[block_target, block_source] becomes a plain
// GEMM over K = k*k*2C that TMA can feed.  One CTA = 4 pixels; tap tables are built once per pixel in smem.
template <typename T, int KK>
__global__ void __launch_bounds__(256)
attn_unfold_kernel(const T *__restrict__ src, int64_t lds, const T *__restrict__ tgt, int64_t ldt, const float *__restrict__ flow,
                   T *__restrict__ out, int64_t ldo, int64_t npix_total, int h, int C)
{
    constexpr int K = (KK == 25) ? 5 : 3;
    constexpr int P = 4;
    __shared__ int s_idx[P][KK][4];
    __shared__ float s_w[P][KK][4];
    __shared__ int s_tidx[P][KK];
    const int64_t pix0 = (int64_t)blockIdx.x * P;
    for (int i = threadIdx.x; i < P * KK; i += blockDim.x) {
        const int pp = i / KK, t = i % KK;
        const int64_t pix = pix0 + pp;
        if (pix < npix_total) {
            const int x = (int)(pix % h), y = (int)((pix / h) % h);
            const BETap tp = be_tap(flow[pix * 2], flow[pix * 2 + 1], y, x, t / K, t % K, K, h, h);
#pragma unroll
            for (int q = 0; q < 4; ++q) { s_idx[pp][t][q] = tp.idx[q]; s_w[pp][t][q] = tp.w[q]; }
            const int iy = max(min(y + t / K - K / 2, h - 1), 0), ix = max(min(x + t % K - K / 2, h - 1), 0);
            s_tidx[pp][t] = iy * h + ix;
        }
    }
    __syncthreads();
    const int chunks = C / 8;
    for (int i = threadIdx.x; i < P * chunks; i += blockDim.x) {
        const int pp = i / chunks, cc = i % chunks;
        const int64_t pix = pix0 + pp;
        if (pix >= npix_total) continue;
        const int64_t plane = (pix / ((int64_t)h * h)) * h * h;
        T *o = out + pix * ldo + cc * 8;
        for (int t = 0; t < KK; ++t) {
            float v[8];
            load8(tgt + (plane + s_tidx[pp][t]) * ldt + cc * 8, v);
            store8(o + (int64_t)t * 2 * C, v);
            float smp[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) smp[j] = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float u[8];
                load8(src + (plane + s_idx[pp][t][q]) * lds + cc * 8, u);
                const float wq = s_w[pp][t][q];
#pragma unroll
                for (int j = 0; j < 8; ++j) smp[j] = __fmaf_rn(wq, u[j], smp[j]);
            }
            store8(o + (int64_t)t * 2 * C + C, smp);
        }
    }
}

// block_extractor_kernel.cu:62-69 clamping materialised once: dst[n,y,x,:] = src[n, clamp(y-pad), clamp(x-pad), :]
template <typename T>
__global__ void replicate_pad_kernel(const T *__restrict__ src, int64_t lds, T *__restrict__ dst, int64_t ldd, int N, int h, int C, int pad)
{
    const int hp = h + 2 * pad, chunks = C / 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * hp * hp * chunks) return;
    const int cc = (int)(i % chunks);
    const int64_t pix = i / chunks;
    const int x = (int)(pix % hp), y = (int)((pix / hp) % hp), n = (int)(pix / ((int64_t)hp * hp));
    const int sy = max(min(y - pad, h - 1), 0), sx = max(min(x - pad, h - 1), 0);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + (((int64_t)n * h + sy) * h + sx) * lds + cc * 8);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + pix * ldd + cc * 8);
    if (sizeof(T) == 2) d4[0] = s4[0];
    else { d4[0] = s4[0]; d4[1] = s4[1]; }
}

__device__ __forceinline__ void load4(const __nv_bfloat16 *p, float v[4])
{
    const uint2 r = *reinterpret_cast<const uint2 *>(p);
    unpack2<__nv_bfloat16>(r.x, v[0], v[1]); unpack2<__nv_bfloat16>(r.y, v[2], v[3]);
}
__device__ __forceinline__ void load4(const __half *p, float v[4])
{
    const uint2 r = *reinterpret_cast<const uint2 *>(p);
    unpack2<__half>(r.x, v[0], v[1]); unpack2<__half>(r.y, v[2], v[3]);
}
__device__ __forceinline__ void load4(const float *p, float v[4])
{
    const float4 r = *reinterpret_cast<const float4 *>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}

// extract_attn.py:24-28 with the k5s5 conv commuted through the bilinear interpolation (include/hoig_b200.h,
// "local attention, tensor-core formulation").  Eight lanes per pixel, four pixels per warp:
//   hidden (128 ch, 16 per lane) = LeakyReLU(Gt(p) + sum_q w_q Gs(p0+q) + b1)
//   logits = W2 hidden + b2 (partial dot products per lane, butterfly over the 8 lanes), softmax in registers
//   coef   = the KK attention weights x the 4 bilinear weights folded into one (K+1)^2 patch
//   dst    = tgt + (1/KK) sum_u coef[u] * src[clamp(p0 - K/2 + u)]
template <typename T, int KK>
__global__ void __launch_bounds__(256, 3)
attn_combine_kernel(const T *__restrict__ gt, int64_t ldgt, const T *__restrict__ gs, int64_t ldgs, const float *__restrict__ b1,
                    const float *__restrict__ w2, const float *__restrict__ b2, const T *__restrict__ src, int64_t lds,
                    const float *__restrict__ flow, const T *__restrict__ tgt, int64_t ldt, T *__restrict__ dst, int64_t ldd,
                    int64_t npix_total, int h, int C)
{
    constexpr int K = (KK == 25) ? 5 : 3;
    constexpr int HID = 128, PK = K + 1, R = K / 2;
    __shared__ __align__(16) float s_w2[KK][HID];
    __shared__ __align__(16) float s_b1[HID];
    __shared__ float s_b2[KK];
    __shared__ float s_coef[8][4][PK * PK + 4];
    for (int i = threadIdx.x; i < KK * HID; i += blockDim.x) (&s_w2[0][0])[i] = w2[i];
    for (int i = threadIdx.x; i < HID; i += blockDim.x) s_b1[i] = b1[i];
    for (int i = threadIdx.x; i < KK; i += blockDim.x) s_b2[i] = b2[i];
    __syncthreads();
    const int lane = threadIdx.x % 32, sub = lane & 7, grp = lane >> 3;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32, nwarps = (int64_t)gridDim.x * blockDim.x / 32;
    const int hpt = h + 2 * R, hps = h + 4 * R;
    for (int64_t base = warp0 * 4; base < npix_total; base += nwarps * 4) {
        const bool valid = base + grp < npix_total;
        const int64_t pix = valid ? base + grp : npix_total - 1;   // idle groups shadow the last pixel: the warp stays converged
        const int x = (int)(pix % h), y = (int)((pix / h) % h);
        const int64_t n = pix / ((int64_t)h * h);
        // centre tap of be_tap(): all k*k taps share these fractions
        const float dx = __fadd_rn(__fadd_rn(flow[pix * 2], 0.f), (float)x), dy = __fadd_rn(__fadd_rn(flow[pix * 2 + 1], 0.f), (float)y);
        const float fdx = floorf(dx), fdy = floorf(dy);
        const float wx1 = __fsub_rn(dx, fdx), wx0 = __fsub_rn(1.f, wx1), wy1 = __fsub_rn(dy, fdy), wy0 = __fsub_rn(1.f, wy1);
        const int x0 = (int)fminf(fmaxf(fdx, -(float)(K + 2)), (float)(h + K + 2));
        const int y0 = (int)fminf(fmaxf(fdy, -(float)(K + 2)), (float)(h + K + 2));

        // ---- hidden: channels 32*i + 4*sub + j
        float hv[16];
        {
            const T *g = gt + ((n * hpt + y + R) * hpt + x + R) * ldgt + 4 * sub;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                load4(g + 32 * i, hv + 4 * i);
                const float4 bv = *reinterpret_cast<const float4 *>(&s_b1[32 * i + 4 * sub]);
                hv[4 * i] += bv.x; hv[4 * i + 1] += bv.y; hv[4 * i + 2] += bv.z; hv[4 * i + 3] += bv.w;
            }
#pragma unroll
            for (int qy = 0; qy < 2; ++qy)
#pragma unroll
                for (int qx = 0; qx < 2; ++qx) {
                    const int cy = max(min(y0 + qy, h - 1 + R), -R) + 2 * R, cx = max(min(x0 + qx, h - 1 + R), -R) + 2 * R;
                    const float w = __fmul_rn(qx ? wx1 : wx0, qy ? wy1 : wy0);
                    const T *gq = gs + ((n * hps + cy) * hps + cx) * ldgs + 4 * sub;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float u[4];
                        load4(gq + 32 * i, u);
#pragma unroll
                        for (int j = 0; j < 4; ++j) hv[4 * i + j] = fmaf(w, u[j], hv[4 * i + j]);
                    }
                }
#pragma unroll
            for (int j = 0; j < 16; ++j) hv[j] = hv[j] > 0.f ? hv[j] : 0.01f * hv[j];
        }
        // ---- logits + softmax (every lane of the group ends up with all KK weights)
        float a[KK];
#pragma unroll
        for (int t = 0; t < KK; ++t) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 wv = *reinterpret_cast<const float4 *>(&s_w2[t][32 * i + 4 * sub]);
                acc = fmaf(hv[4 * i], wv.x, acc); acc = fmaf(hv[4 * i + 1], wv.y, acc);
                acc = fmaf(hv[4 * i + 2], wv.z, acc); acc = fmaf(hv[4 * i + 3], wv.w, acc);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            a[t] = acc + s_b2[t];
        }
        float mx = a[0];
#pragma unroll
        for (int t = 1; t < KK; ++t) mx = fmaxf(mx, a[t]);
        float den = 0.f;
#pragma unroll
        for (int t = 0; t < KK; ++t) { a[t] = expf(a[t] - mx); den += a[t]; }
        const float inv = 1.0f / (den * (float)KK);
        // ---- fold attention weights x bilinear weights into the (K+1)^2 patch (kept in shared memory: the gather
        //      loop below then needs few registers and more warps fit on the SM to hide its load latency)
        float *cf = &s_coef[threadIdx.x / 32][grp][0];
        __syncwarp();
#pragma unroll
        for (int uy = 0; uy < PK; ++uy)
#pragma unroll
            for (int ux = 0; ux < PK; ++ux) {
                float c = 0.f;
                if (uy < K && ux < K) c = fmaf(a[uy * K + ux], wy0 * wx0, c);
                if (uy < K && ux > 0) c = fmaf(a[uy * K + ux - 1], wy0 * wx1, c);
                if (uy > 0 && ux < K) c = fmaf(a[(uy - 1) * K + ux], wy1 * wx0, c);
                if (uy > 0 && ux > 0) c = fmaf(a[(uy - 1) * K + ux - 1], wy1 * wx1, c);
                if (((uy * PK + ux) & 7) == sub) cf[uy * PK + ux] = c * inv;
            }
        __syncwarp();
        const T *splane = src + n * (int64_t)h * h * lds;
        const int xb = x0 - R, yb = y0 - R;
        for (int cc = sub; cc < C / 8; cc += 8) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 2
            for (int uy = 0; uy < PK; ++uy) {
                const int py = max(min(yb + uy, h - 1), 0);
                const T *srow = splane + (int64_t)py * h * lds + cc * 8;
#pragma unroll
                for (int ux = 0; ux < PK; ++ux) {
                    float u[8];
                    load8(srow + (int64_t)max(min(xb + ux, h - 1), 0) * lds, u);
                    const float c = cf[uy * PK + ux];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(c, u[j], acc[j]);
                }
            }
            float tv[8];
            load8(tgt + pix * ldt + cc * 8, tv);
#pragma unroll
            for (int j = 0; j < 8; ++j) tv[j] += acc[j];
            if (valid) store8(dst + pix * ldd + cc * 8, tv);
        }
    }
}

// ---- attn_combine on the tensor cores, source tiles staged in shared memory -------------------------------------------------------
// The weighted patch sum  out[p][c] = sum_u coef[p][u] * src[patch_p(u)][c]  (36 taps per pixel, all channels) is 85 % of attn_combine's
// work and was instruction-issue bound as per-pixel gathers.  For a TILE of 8 x 8 pixels whose (unclamped) patches fall into a window of
// WW x WH <= 256 source positions -- the case for HOGAN's flows, which are normalised coordinates used as pixel offsets (quirk Q1,
// |flow| <= 3) and vary smoothly -- it is a small dense GEMM:   OUT[64 px][C] = A[64 px][WW*WH] * S[WW*WH][C]
//   A: the pixels' 36 coefficients scattered to their window positions (zeros elsewhere), built once per tile in shared memory (fp16/bf16);
//   S: the window's source pixels (border positions replicate the clamped pixel, so unclamped patch coordinates index it directly), staged
//      in shared memory 64 channels at a time with 16-byte cp.async copies (XOR-swizzled rows, conflict-free ldmatrix);
//   warp-level mma.sync.m16n8k16 (fp32 accumulate): 8 warps = 4 row tiles x 2 halves of the 64-channel slab.
// The L2->SM traffic per pixel drops from 36 source pixels to WW*WH/64 (~3-5), and the instruction count ~7x.  Tiles whose window is larger
// (arbitrary flows) take the per-pixel gather path inside the same kernel, so the result is defined for every input.
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    if (sizeof(T) == 2 && std::is_same<T, __half>::value)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int ATC_TILE = 8, ATC_PX = 64, ATC_KMAX = 256, ATC_AP = 264;       // tile side, pixels per tile, window capacity (16 x 16), A row pitch (halves)
constexpr int ATC_CONST_BYTES = 15360, ATC_A_BYTES = ATC_PX * ATC_AP * 2, ATC_S_BYTES = 2 * ATC_KMAX * 128;   // S: two buffers (slab s + 1 staged during slab s)
constexpr int ATC_SMEM = ATC_CONST_BYTES + ATC_A_BYTES + ATC_S_BYTES;           // 114688 B: two CTAs per SM

template <typename T>
__global__ void __launch_bounds__(256, 2)
attn_combine_tc_kernel(const T *__restrict__ gt, int64_t ldgt, const T *__restrict__ gs, int64_t ldgs, const float *__restrict__ b1,
                       const float *__restrict__ w2, const float *__restrict__ b2, const T *__restrict__ src, int64_t lds,
                       const float *__restrict__ flow, const T *__restrict__ tgt, int64_t ldt, T *__restrict__ dst, int64_t ldd,
                       int N, int h, int C, int debug, int g_phase1_tc)
{
    constexpr int K = 5, KK = 25, HID = 128, PK = K + 1, R = K / 2;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float *s_w2 = reinterpret_cast<float *>(smem_raw);              // [KK][HID]
    float *s_b1 = s_w2 + KK * HID;                                  // [HID]
    float *s_b2 = s_b1 + HID;                                       // [KK] (32 reserved)
    int *s_box = reinterpret_cast<int *>(s_b2 + 32);                // bx0, by0, WW, fits
    int *s_xy = s_box + 4;                                          // [64][2]: x0, y0 of every pixel of the tile
    int *s_off = s_xy + 2 * ATC_PX;                                 // [ATC_KMAX]: element offset of the (clamped) source pixel of every window position
    T *sA = reinterpret_cast<T *>(smem_raw + ATC_CONST_BYTES);
    uint8_t *sS = smem_raw + ATC_CONST_BYTES + ATC_A_BYTES;
    float *s_coef = reinterpret_cast<float *>(sS);                  // fallback path only: [8 warps][4][PK*PK+4] (S is idle then)
    const uint32_t sA_u = (uint32_t)__cvta_generic_to_shared(sA), sS_u = (uint32_t)__cvta_generic_to_shared(sS);
    for (int i = threadIdx.x; i < KK * HID; i += blockDim.x) s_w2[i] = w2[i];
    for (int i = threadIdx.x; i < HID; i += blockDim.x) s_b1[i] = b1[i];
    for (int i = threadIdx.x; i < KK; i += blockDim.x) s_b2[i] = b2[i];
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32, sub = lane & 7, grp = lane >> 3;
    const int tiles_x = h / ATC_TILE, tiles_per_img = tiles_x * tiles_x;
    const int hpt = h + 2 * R, hps = h + 4 * R;
    for (int tile = blockIdx.x; tile < N * tiles_per_img; tile += gridDim.x) {
        const int n = tile / tiles_per_img, t_in = tile - n * tiles_per_img;
        const int ty = t_in / tiles_x, tx = t_in - ty * tiles_x;
        __syncthreads();                       // constants loaded / previous tile fully done with A, S and s_xy
        if (threadIdx.x < ATC_PX) {
            const int x = tx * ATC_TILE + (threadIdx.x & 7), y = ty * ATC_TILE + (threadIdx.x >> 3);
            const int64_t pix = ((int64_t)n * h + y) * h + x;
            const float dx = __fadd_rn(__fadd_rn(flow[pix * 2], 0.f), (float)x), dy = __fadd_rn(__fadd_rn(flow[pix * 2 + 1], 0.f), (float)y);
            s_xy[2 * threadIdx.x] = (int)fminf(fmaxf(floorf(dx), -(float)(K + 2)), (float)(h + K + 2));
            s_xy[2 * threadIdx.x + 1] = (int)fminf(fmaxf(floorf(dy), -(float)(K + 2)), (float)(h + K + 2));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int mnx = s_xy[0], mxx = s_xy[0], mny = s_xy[1], mxy = s_xy[1];
            for (int i = 1; i < ATC_PX; ++i) {
                mnx = min(mnx, s_xy[2 * i]); mxx = max(mxx, s_xy[2 * i]);
                mny = min(mny, s_xy[2 * i + 1]); mxy = max(mxy, s_xy[2 * i + 1]);
            }
            const int WW = mxx - mnx + PK, WH = mxy - mny + PK;
            s_box[0] = mnx - R; s_box[1] = mny - R; s_box[2] = WW;
            s_box[3] = (WW * WH <= ATC_KMAX) ? WW * WH : 0;
        }
        __syncthreads();
        const int bx0 = s_box[0], by0 = s_box[1], WW = s_box[2], wpos = s_box[3];
        const bool fits = wpos > 0;
        const int ksteps = (wpos + 15) / 16, kcols = ksteps * 16;
        if (fits) {           // zero the used columns of A (16-byte stores; AP * 2 and kcols * 2 are multiples of 16)
            for (int pos = threadIdx.x; pos < wpos; pos += blockDim.x) {
                const int wy = pos / WW, wx = pos - wy * WW;
                s_off[pos] = (max(min(by0 + wy, h - 1), 0) * h + max(min(bx0 + wx, h - 1), 0)) * (int)lds;
            }
            const int per_row = kcols / 8;
            for (int i = threadIdx.x; i < ATC_PX * per_row; i += blockDim.x)
                *reinterpret_cast<uint4 *>(sA + (i / per_row) * ATC_AP + (i % per_row) * 8) = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        // ---- phase 1 (staged path): hidden -> shared memory, logits as a 64 x 32 x 128 warp-level GEMM, softmax + coefficients 4 threads/pixel
        if (fits && g_phase1_tc && !(debug & 2)) {
            T *sH = reinterpret_cast<T *>(sS);                                   // [64 px][136]
            T *sW = reinterpret_cast<T *>(sS + 17408);                           // [32 logits][136] (rows >= KK are zero)
            float *sL = reinterpret_cast<float *>(sS + 26112);                   // [64 px][32]
            float *s_fr = reinterpret_cast<float *>(sS + 34304);                 // [64 px][2]: bilinear fractions wx1, wy1
            constexpr int HP = 136;
            for (int i = threadIdx.x; i < 32 * (HID / 8); i += blockDim.x) {     // W2 -> 16-bit, once per tile (the S region is reused by phase 2)
                const int r = i / (HID / 8), c8 = (i % (HID / 8)) * 8;
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = r < KK ? s_w2[r * HID + c8 + j] : 0.f;
                store8(sW + r * HP + c8, v);
            }
            if (threadIdx.x < ATC_PX) {
                const int x = tx * ATC_TILE + (threadIdx.x & 7), y = ty * ATC_TILE + (threadIdx.x >> 3);
                const int64_t pix = ((int64_t)n * h + y) * h + x;
                const float dx = __fadd_rn(__fadd_rn(flow[pix * 2], 0.f), (float)x), dy = __fadd_rn(__fadd_rn(flow[pix * 2 + 1], 0.f), (float)y);
                s_fr[2 * threadIdx.x] = __fsub_rn(dx, floorf(dx));
                s_fr[2 * threadIdx.x + 1] = __fsub_rn(dy, floorf(dy));
            }
            __syncthreads();
            // (a) hidden[px][128] = LeakyReLU(Gt + bilinear(Gs) + b1): one (pixel, 8-channel chunk) item per thread and step
            for (int i = threadIdx.x; i < ATC_PX * (HID / 8); i += blockDim.x) {
                const int pi = i / (HID / 8), c8 = (i % (HID / 8)) * 8;
                const int x = tx * ATC_TILE + (pi & 7), y = ty * ATC_TILE + (pi >> 3);
                const int x0 = s_xy[2 * pi], y0 = s_xy[2 * pi + 1];
                const float wx1 = s_fr[2 * pi], wy1 = s_fr[2 * pi + 1], wx0 = __fsub_rn(1.f, wx1), wy0 = __fsub_rn(1.f, wy1);
                float hv[8], u[8];
                load8(gt + (((int64_t)n * hpt + y + R) * hpt + x + R) * ldgt + c8, hv);
#pragma unroll
                for (int j = 0; j < 8; ++j) hv[j] += s_b1[c8 + j];
#pragma unroll
                for (int qy = 0; qy < 2; ++qy)
#pragma unroll
                    for (int qx = 0; qx < 2; ++qx) {
                        const int cy = max(min(y0 + qy, h - 1 + R), -R) + 2 * R, cx = max(min(x0 + qx, h - 1 + R), -R) + 2 * R;
                        const float w = __fmul_rn(qx ? wx1 : wx0, qy ? wy1 : wy0);
                        load8(gs + (((int64_t)n * hps + cy) * hps + cx) * ldgs + c8, u);
#pragma unroll
                        for (int j = 0; j < 8; ++j) hv[j] = fmaf(w, u[j], hv[j]);
                    }
#pragma unroll
                for (int j = 0; j < 8; ++j) hv[j] = hv[j] > 0.f ? hv[j] : 0.01f * hv[j];
                store8(sH + pi * HP + c8, hv);
            }
            __syncthreads();
            // (b) logits[64][32] = hidden[64][128] * W2^T: warp = (row tile, half of the 32 logit columns), 8 k-steps
            {
                const int mt = warp & 3, nh = warp >> 2, g = lane >> 2, tq = lane & 3;
                const uint32_t sH_u = (uint32_t)__cvta_generic_to_shared(sH), sW_u = (uint32_t)__cvta_generic_to_shared(sW);
                const uint32_t a_ad = sH_u + (uint32_t)(((mt * 16 + (lane & 15)) * HP + (lane >> 4) * 8) * 2);
                const int mi = lane >> 3;
                const uint32_t b_ad = sW_u + (uint32_t)(((nh * 16 + (mi >> 1) * 8 + (lane & 7)) * HP + (mi & 1) * 8) * 2);
                float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
                for (int ks = 0; ks < HID / 16; ++ks) {
                    uint32_t af[4], bf[4];
                    ldmatrix_x4(af, a_ad + (uint32_t)(ks * 32));
                    ldmatrix_x4(bf, b_ad + (uint32_t)(ks * 32));
                    mma_16816<T>(acc[0], af, bf[0], bf[1]);
                    mma_16816<T>(acc[1], af, bf[2], bf[3]);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int col = (nh * 2 + j) * 8 + 2 * tq;
                    *reinterpret_cast<float2 *>(sL + (mt * 16 + g) * 32 + col) = make_float2(acc[j][0], acc[j][1]);
                    *reinterpret_cast<float2 *>(sL + (mt * 16 + g + 8) * 32 + col) = make_float2(acc[j][2], acc[j][3]);
                }
            }
            __syncthreads();
            // (c) softmax and the 36 patch coefficients, four threads per pixel (nine coefficients each), scattered into A
            {
                const int pi = threadIdx.x >> 2, part = threadIdx.x & 3;
                float a[KK];
                float mx = -3.0e38f;
#pragma unroll
                for (int t = 0; t < KK; ++t) { a[t] = sL[pi * 32 + t] + s_b2[t]; mx = fmaxf(mx, a[t]); }
                float den = 0.f;
#pragma unroll
                for (int t = 0; t < KK; ++t) { a[t] = expf(a[t] - mx); den += a[t]; }
                const float inv = 1.0f / (den * (float)KK);
                const float wx1 = s_fr[2 * pi], wy1 = s_fr[2 * pi + 1], wx0 = __fsub_rn(1.f, wx1), wy0 = __fsub_rn(1.f, wy1);
                T *arow = sA + pi * ATC_AP + (s_xy[2 * pi + 1] - R - by0) * WW + (s_xy[2 * pi] - R - bx0);
#pragma unroll
                for (int uy = 0; uy < PK; ++uy)
#pragma unroll
                    for (int ux = 0; ux < PK; ++ux) {
                        if ((uy * PK + ux) / 9 != part) continue;
                        float c = 0.f;
                        if (uy < K && ux < K) c = fmaf(a[uy * K + ux], wy0 * wx0, c);
                        if (uy < K && ux > 0) c = fmaf(a[uy * K + ux - 1], wy0 * wx1, c);
                        if (uy > 0 && ux < K) c = fmaf(a[(uy - 1) * K + ux], wy1 * wx0, c);
                        if (uy > 0 && ux > 0) c = fmaf(a[(uy - 1) * K + ux - 1], wy1 * wx1, c);
                        DT<T>::st(arow + uy * WW + ux, c * inv);
                    }
            }
        } else
        // ---- phase 1 (gather path, or staged path with HOIG_ATTN_PHASE1_TC=0): per pixel (8 lanes each): hidden, logits, softmax, coefficients
#pragma unroll 1
        for (int pass = 0; pass < ((debug & 2) ? 0 : 2); ++pass) {
            const int pi = pass * 32 + warp * 4 + grp;          // pixel of the tile
            const int x = tx * ATC_TILE + (pi & 7), y = ty * ATC_TILE + (pi >> 3);
            const int64_t pix = ((int64_t)n * h + y) * h + x;
            const float dx = __fadd_rn(__fadd_rn(flow[pix * 2], 0.f), (float)x), dy = __fadd_rn(__fadd_rn(flow[pix * 2 + 1], 0.f), (float)y);
            const float fdx = floorf(dx), fdy = floorf(dy);
            const float wx1 = __fsub_rn(dx, fdx), wx0 = __fsub_rn(1.f, wx1), wy1 = __fsub_rn(dy, fdy), wy0 = __fsub_rn(1.f, wy1);
            const int x0 = s_xy[2 * pi], y0 = s_xy[2 * pi + 1];
            float hv[16];
            {
                const T *g = gt + (((int64_t)n * hpt + y + R) * hpt + x + R) * ldgt + 4 * sub;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    load4(g + 32 * i, hv + 4 * i);
                    const float4 bv = *reinterpret_cast<const float4 *>(&s_b1[32 * i + 4 * sub]);
                    hv[4 * i] += bv.x; hv[4 * i + 1] += bv.y; hv[4 * i + 2] += bv.z; hv[4 * i + 3] += bv.w;
                }
#pragma unroll
                for (int qy = 0; qy < 2; ++qy)
#pragma unroll
                    for (int qx = 0; qx < 2; ++qx) {
                        const int cy = max(min(y0 + qy, h - 1 + R), -R) + 2 * R, cx = max(min(x0 + qx, h - 1 + R), -R) + 2 * R;
                        const float w = __fmul_rn(qx ? wx1 : wx0, qy ? wy1 : wy0);
                        const T *gq = gs + (((int64_t)n * hps + cy) * hps + cx) * ldgs + 4 * sub;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float u[4];
                            load4(gq + 32 * i, u);
#pragma unroll
                            for (int j = 0; j < 4; ++j) hv[4 * i + j] = fmaf(w, u[j], hv[4 * i + j]);
                        }
                    }
#pragma unroll
                for (int j = 0; j < 16; ++j) hv[j] = hv[j] > 0.f ? hv[j] : 0.01f * hv[j];
            }
            float a[KK];
#pragma unroll
            for (int t = 0; t < KK; ++t) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 wv = *reinterpret_cast<const float4 *>(&s_w2[t * HID + 32 * i + 4 * sub]);
                    acc = fmaf(hv[4 * i], wv.x, acc); acc = fmaf(hv[4 * i + 1], wv.y, acc);
                    acc = fmaf(hv[4 * i + 2], wv.z, acc); acc = fmaf(hv[4 * i + 3], wv.w, acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                a[t] = acc + s_b2[t];
            }
            float mx = a[0];
#pragma unroll
            for (int t = 1; t < KK; ++t) mx = fmaxf(mx, a[t]);
            float den = 0.f;
#pragma unroll
            for (int t = 0; t < KK; ++t) { a[t] = expf(a[t] - mx); den += a[t]; }
            const float inv = 1.0f / (den * (float)KK);
            float *cf = s_coef + ((warp * 4 + grp) * (PK * PK + 4));
            T *arow = sA + pi * ATC_AP + (y0 - R - by0) * WW + (x0 - R - bx0);
#pragma unroll
            for (int uy = 0; uy < PK; ++uy)
#pragma unroll
                for (int ux = 0; ux < PK; ++ux) {
                    float c = 0.f;
                    if (uy < K && ux < K) c = fmaf(a[uy * K + ux], wy0 * wx0, c);
                    if (uy < K && ux > 0) c = fmaf(a[uy * K + ux - 1], wy0 * wx1, c);
                    if (uy > 0 && ux < K) c = fmaf(a[(uy - 1) * K + ux], wy1 * wx0, c);
                    if (uy > 0 && ux > 0) c = fmaf(a[(uy - 1) * K + ux - 1], wy1 * wx1, c);
                    if (((uy * PK + ux) & 7) == sub) {
                        if (fits) DT<T>::st(arow + uy * WW + ux, c * inv);
                        else cf[uy * PK + ux] = c * inv;
                    }
                }
            if (!fits) {      // per-pixel gathers (window too large for the staged GEMM): same arithmetic as attn_combine_kernel
                __syncwarp();
                const T *splane = src + (int64_t)n * h * h * lds;
                const int xb = x0 - R, yb = y0 - R;
                for (int cc = sub; cc < C / 8; cc += 8) {
                    float acc[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 2
                    for (int uy = 0; uy < PK; ++uy) {
                        const T *srow = splane + (int64_t)max(min(yb + uy, h - 1), 0) * h * lds + cc * 8;
#pragma unroll
                        for (int ux = 0; ux < PK; ++ux) {
                            float u[8];
                            load8(srow + (int64_t)max(min(xb + ux, h - 1), 0) * lds, u);
                            const float c = cf[uy * PK + ux];
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[j] = fmaf(c, u[j], acc[j]);
                        }
                    }
                    float tv[8];
                    load8(tgt + pix * ldt + cc * 8, tv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) tv[j] += acc[j];
                    store8(dst + pix * ldd + cc * 8, tv);
                }
                __syncwarp();
            }
        }
        if (!fits || (debug & 1)) continue;
        // ---- phase 2: OUT[64][C] = A[64][kcols] * S[kcols][C], 64 channels per slab; slab s + 1 is staged (cp.async) while slab s is
        //      multiplied when the window fits twice into the S region
        const int mt = warp & 3, nh = warp >> 2, g = lane >> 2, tq = lane & 3;
        const uint32_t a_addr = sA_u + (uint32_t)(((mt * 16 + (lane & 15)) * ATC_AP + (lane >> 4) * 8) * 2);
        const int b_krow = (lane & 7) + ((lane >> 3) & 1) * 8, b_csel = lane >> 4;     // ldmatrix.trans source row / chunk select of this lane
        const T *splane = src + (int64_t)n * h * h * lds;
        const int nslab = C / 64;
        auto stage = [&](int slab, uint32_t base) {
            for (int i = threadIdx.x; i < kcols * 8; i += blockDim.x) {
                const int pos = i >> 3, c = i & 7;
                const uint32_t d = base + (uint32_t)(pos * 128 + ((c ^ (pos & 7)) << 4));
                if (pos < wpos) cp_async16(d, splane + s_off[pos] + slab * 64 + c * 8);
                else asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(d), "r"(0u) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        int64_t pix_r[2];
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int r = mt * 16 + g + hrow * 8;
            pix_r[hrow] = ((int64_t)n * h + ty * ATC_TILE + (r >> 3)) * h + tx * ATC_TILE + (r & 7);
        }
        __syncthreads();                       // A and s_off complete
        stage(0, sS_u);
#pragma unroll 1
        for (int slab = 0; slab < nslab; ++slab) {
            const uint32_t sbuf = sS_u + (uint32_t)((slab & 1) * ATC_KMAX * 128);
            if (slab + 1 < nslab) {
                stage(slab + 1, sS_u + (uint32_t)(((slab + 1) & 1) * ATC_KMAX * 128));
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            // the target values this lane will add to (loaded now, consumed after the MMAs)
            uint32_t tv[2][4];
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    tv[hrow][j] = *reinterpret_cast<const uint32_t *>(tgt + pix_r[hrow] * ldt + slab * 64 + (nh * 4 + j) * 8 + 2 * tq);
            __syncthreads();                   // this slab's S is visible to every warp
            float acc[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll 3
            for (int ks = 0; ks < ksteps; ++ks) {
                uint32_t af[4], bf0[4], bf1[4];
                ldmatrix_x4(af, a_addr + (uint32_t)(ks * 32));
                const int krow = ks * 16 + b_krow;
                const uint32_t rb = sbuf + (uint32_t)(krow * 128);
                ldmatrix_x4_trans(bf0, rb + (uint32_t)((((nh * 4 + 0 + b_csel) ^ (krow & 7))) << 4));
                ldmatrix_x4_trans(bf1, rb + (uint32_t)((((nh * 4 + 2 + b_csel) ^ (krow & 7))) << 4));
                mma_16816<T>(acc[0], af, bf0[0], bf0[1]);
                mma_16816<T>(acc[1], af, bf0[2], bf0[3]);
                mma_16816<T>(acc[2], af, bf1[0], bf1[1]);
                mma_16816<T>(acc[3], af, bf1[2], bf1[3]);
            }
            // epilogue: dst = tgt + OUT; lane holds rows g, g + 8 of its row tile and channels 2*tq, 2*tq + 1 of each 8-channel group
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float lo, hi;
                    unpack2<T>(tv[hrow][j], lo, hi);
                    *reinterpret_cast<uint32_t *>(dst + pix_r[hrow] * ldd + slab * 64 + (nh * 4 + j) * 8 + 2 * tq) =
                        pack2<T>(lo + acc[j][hrow * 2], hi + acc[j][hrow * 2 + 1]);
                }
            __syncthreads();                   // every warp is done with this slab's S before it is overwritten
        }
    }
}

// x7[b,y,x, s*C + c] = x[b,c,y,x+s-k/2]  (zero outside the row / beyond k*C), from the NCHW f32 input
template <typename T>
__global__ void hunfold_kernel(const float *__restrict__ src, int B, int C, int H, int W, int k, T *__restrict__ dst, int64_t ldd, int Cpad)
{
    // one thread per pixel: loads are coalesced along x (consecutive lanes = consecutive pixels of a row)
    const int64_t bp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (bp >= (int64_t)B * H * W) return;
    const int x = (int)(bp % W), y = (int)((bp / W) % H), b = (int)(bp / ((int64_t)W * H));
    const float *row = src + ((int64_t)b * C * H + y) * W;
    for (int ch = 0; ch < Cpad / 8; ++ch) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kc = ch * 8 + j;
            const int s = kc / C, c = kc - s * C;
            const int xx = x + s - k / 2;
            v[j] = (s < k && xx >= 0 && xx < W) ? __ldg(row + (int64_t)c * H * W + xx) : 0.f;
        }
        store8(dst + bp * ldd + ch * 8, v);
    }
}

struct FoldSegs8 {          // per output channel g: its NCHW plane base (image 0) and the image stride in elements
    float *out[8];
    int64_t bstride[8];
};
// Same result, staged through shared memory: a CTA takes 128 consecutive pixels of one image row, loads the C input rows
// (plus k/2 halo pixels each side) with coalesced reads, and writes each pixel's Cpad channels as consecutive 16-byte
// chunks from consecutive lanes (the per-pixel kernel above scatters 16-byte pieces 128+ bytes apart).
template <typename T>
__global__ void __launch_bounds__(256) hunfold_row_kernel(const float *__restrict__ src, int C, int H, int W, int k, T *__restrict__ dst,
                                                          int64_t ldd, int Cpad, int rows)
{
    // A CTA takes `rows` consecutive image rows of one 128-pixel column segment.  Every thread owns ONE 8-channel chunk position
    // (its (tap, input channel) sources stay in registers) and walks the pixels, so a warp writes four pixels' 128 contiguous bytes
    // per store instruction and the per-row work is: C x 134 coalesced loads, a barrier, 8 shared loads + one 16-byte store per chunk.
    constexpr int PX = 128, MAXC = 16, MAXK = 7, TW = PX + MAXK - 1;
    __shared__ float tile[2][MAXC][TW];
    const int segs = W / PX, rgroups = H / rows;
    const int x0 = (blockIdx.x % segs) * PX, y0 = ((blockIdx.x / segs) % rgroups) * rows, b = blockIdx.x / (segs * rgroups);
    const int halo = k / 2, tw = PX + k - 1;
    const int chunks = Cpad / 8, ch = threadIdx.x % chunks, px0 = threadIdx.x / chunks, pstep = blockDim.x / chunks;
    int off[8];                                   // source of channel ch*8 + j inside a tile row set: c * TW + tap, or -1 (zero padding)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int kc = ch * 8 + j, sft = kc / C;
        off[j] = sft < k ? (kc - sft * C) * TW + sft : -1;
    }
    auto load_row = [&](int y, int buf) {
        const float *row = src + ((int64_t)b * C * H + y) * W;
        for (int i = threadIdx.x; i < C * tw; i += blockDim.x) {
            const int c = i / tw, j = i - c * tw, xx = x0 - halo + j;
            tile[buf][c][j] = (xx >= 0 && xx < W) ? __ldg(row + (int64_t)c * H * W + xx) : 0.f;
        }
    };
    load_row(y0, 0);
    for (int r = 0; r < rows; ++r) {
        __syncthreads();                          // row r is in tile[r & 1]; everyone is done reading tile[(r + 1) & 1]
        if (r + 1 < rows) load_row(y0 + r + 1, (r + 1) & 1);
        const float *tl = &tile[r & 1][0][0];
        T *drow = dst + (((int64_t)b * H + y0 + r) * W + x0) * ldd + ch * 8;
        for (int px = px0; px < PX; px += pstep) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = off[j] >= 0 ? tl[off[j] + px] : 0.f;
            store8(drow + (int64_t)px * ldd, v);
        }
    }
}

// hfold for G == 8 (the merged heads): one thread per pixel reads the k neighbours' 8 partial sums as one 16-byte load each
// and writes the 8 output planes with coalesced stores.
template <typename T>
__global__ void hfold8_kernel(const T *__restrict__ z, int64_t ldz, int B, int H, int W, int k, const int *__restrict__ act_table,
                              FoldSegs8 segs)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * H * W) return;
    const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((int64_t)W * H));
    float acc[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) acc[g] = 0.f;
    const T *zrow = z + ((int64_t)b * H + y) * W * ldz;
    for (int sft = 0; sft < k; ++sft) {
        const int xx = x + sft - k / 2;
        if (xx < 0 || xx >= W) continue;
        float v[8];
        load8(zrow + (int64_t)xx * ldz + sft * 8, v);
#pragma unroll
        for (int g = 0; g < 8; ++g) acc[g] += v[g];
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float r = apply_act(acc[g], act_table ? act_table[g] : HOIG_ACT_NONE);
        if (segs.out[g]) segs.out[g][(int64_t)b * segs.bstride[g] + (int64_t)y * W + x] = r;
    }
}

struct FoldSegs {
    float *out[4];
    int c0[4], n[4], nseg;
};
// y[b,g,y,x] = act_g( sum_s Z[b,y,x+s-k/2, s*G + g] ); one thread per (pixel, g)
template <typename T>
__global__ void hfold_kernel(const T *__restrict__ z, int64_t ldz, int B, int H, int W, int G, int k, const int *__restrict__ act_table,
                             FoldSegs segs)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * G * H * W) return;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const int g = (int)((i / ((int64_t)W * H)) % G), b = (int)(i / ((int64_t)W * H * G));
    float acc = 0.f;
    for (int s = 0; s < k; ++s) {
        const int xx = x + s - k / 2;
        if (xx < 0 || xx >= W) continue;
        acc += DT<T>::ld(z + (((int64_t)b * H + y) * W + xx) * ldz + s * G + g);
    }
    acc = apply_act(acc, act_table ? act_table[g] : HOIG_ACT_NONE);
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (q < segs.nseg && g >= segs.c0[q] && g < segs.c0[q] + segs.n[q])
            segs.out[q][(((int64_t)b * segs.n[q] + (g - segs.c0[q])) * H + y) * W + x] = acc;
}

template <typename F> int dispatch(int dtype, F f)
{
    if (dtype == HOIG_F32) return f((float *)nullptr);
    if (dtype == HOIG_BF16) return f((__nv_bfloat16 *)nullptr);
    if (dtype == HOIG_F16) return f((__half *)nullptr);
    set_error("bad dtype %d", dtype);
    return HOIG_ERR_INVALID;
}

}  // namespace
}  // namespace hoig

using namespace hoig;

extern "C" int hoig_nchw_to_nhwc(const float *src, int B, int C, int H, int W, void *dst, int64_t ldd, int Cpad, int dtype,
                                 hoigStream_t stream)
{
    HOIG_REQUIRE(src && dst && Cpad % 8 == 0 && Cpad >= C && ldd >= Cpad && ldd % 8 == 0, "nchw_to_nhwc: bad argument");
    const int64_t n = (int64_t)B * H * W * (Cpad / 8);
    if (n == 0) return HOIG_OK;
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        nchw_to_nhwc_kernel<T><<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>(src, B, C, H * W, (T *)dst, ldd, Cpad);
        return check_launch("nchw_to_nhwc_kernel");
    });
}

extern "C" int hoig_nhwc_to_nchw(const void *src, int64_t lds, int dtype, int B, int C, int H, int W, float *dst, hoigStream_t stream)
{
    HOIG_REQUIRE(src && dst && lds >= C, "nhwc_to_nchw: bad argument");
    const int64_t n = (int64_t)B * C * H * W;
    if (n == 0) return HOIG_OK;
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        nhwc_to_nchw_kernel<T><<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>((const T *)src, lds, B, C, H * W, dst);
        return check_launch("nhwc_to_nchw_kernel");
    });
}

extern "C" int hoig_seg_resize_nearest(const float *seg, int B, int C, int Hi, int Wi, void *dst, int64_t ldd, int Cpad,
                                       int Ho, int Wo, int dtype, hoigStream_t stream)
{
    HOIG_REQUIRE(seg && dst && Cpad % 8 == 0 && Cpad >= C && ldd >= Cpad && ldd % 8 == 0, "seg_resize: bad argument");
    const int64_t n = (int64_t)B * Ho * Wo * (Cpad / 8);
    if (n == 0) return HOIG_OK;
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        seg_resize_kernel<T><<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>(seg, B, C, Hi, Wi, (T *)dst, ldd, Cpad, Ho, Wo);
        return check_launch("seg_resize_kernel");
    });
}

static int slab_pixels(int HW, int N)
{
    // aim for >= 16 CTAs per SM over the (slab, image) grid (4 resident): >= 4 waves keeps the tail small
    int slabs = (148 * 16 + N - 1) / N;
    if (slabs < 1) slabs = 1;
    int pp = (HW + slabs - 1) / slabs;
    if (pp < 32) pp = 32;
    return pp;
}

extern "C" int hoig_plane_stats(const void *x, int64_t ldx, int dtype, int N, int HW, int C, double *stats, hoigStream_t stream)
{
    HOIG_REQUIRE(x && stats && C % 8 == 0 && C / 8 <= TPB && ldx >= C && ldx % 8 == 0, "plane_stats: bad argument (C=%d)", C);
    if (N == 0 || HW == 0) return HOIG_OK;
    const int pp = slab_pixels(HW, N);
    const int lanes = TPB / (C / 8);
    const size_t smem = (size_t)lanes * C * 2 * sizeof(float);
    dim3 grid(ceil_div(HW, pp), N);
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        plane_stats_kernel<T><<<grid, TPB, smem, as_stream(stream)>>>((const T *)x, ldx, HW, C, pp, stats);
        return check_launch("plane_stats_kernel");
    });
}

extern "C" int hoig_instnorm_apply(const void *x, int64_t ldx, const double *stats, const float *gamma, const float *beta,
                                   const void *gb, int64_t ldgb, const void *residual, int64_t ldr, int relu, void *dst,
                                   int64_t ldd, int dtype, int N, int HW, int C, float eps, hoigStream_t stream)
{
    HOIG_REQUIRE(x && stats && dst && C % 8 == 0 && ldx % 8 == 0 && ldd % 8 == 0, "instnorm_apply: bad argument");
    HOIG_REQUIRE(!gb || (ldgb >= 2 * C && ldgb % 8 == 0), "instnorm_apply: gb needs 2C channels");
    HOIG_REQUIRE(!residual || ldr % 8 == 0, "instnorm_apply: bad residual stride");
    if (N == 0 || HW == 0) return HOIG_OK;
    const int pp = slab_pixels(HW, N);
    dim3 grid(ceil_div(HW, pp), N);
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        instnorm_apply_kernel<T><<<grid, TPB, 3 * C * sizeof(float), as_stream(stream)>>>(
            (const T *)x, ldx, stats, gamma, beta, (const T *)gb, ldgb, (const T *)residual, ldr, relu, (T *)dst, ldd, HW, C, pp, eps);
        return check_launch("instnorm_apply_kernel");
    });
}

extern "C" int hoig_resize_flow(const float *T, int B, int Hi, int Wi, int h, int subtract_identity, float *flow, hoigStream_t stream)
{
    HOIG_REQUIRE(T && flow && h >= 1, "resize_flow: bad argument");
    const int64_t n = (int64_t)B * h * h;
    if (n == 0) return HOIG_OK;
    resize_flow_kernel<<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>(T, B, Hi, Wi, h, subtract_identity, flow);
    return check_launch("resize_flow_kernel");
}

extern "C" int hoig_attn_finish(const void *hidden, int64_t ldh, int Chid, const float *w2, const float *b2, const void *src,
                                int64_t lds, const float *flow, const void *tgt, int64_t ldt, void *dst, int64_t ldd, int dtype,
                                int N, int h, int C, int k, const void *unfold, int64_t ldu, hoigStream_t stream)
{
    HOIG_REQUIRE(hidden && w2 && b2 && src && flow && tgt && dst, "attn_finish: null pointer");
    HOIG_REQUIRE(!unfold || (ldu >= (int64_t)2 * k * k * C && ldu % 8 == 0), "attn_finish: unfold buffer needs 2*k*k*C channels");
    HOIG_REQUIRE(k == 5 || k == 3, "attn_finish: kernel size %d not supported (3 or 5)", k);
    HOIG_REQUIRE(C % 8 == 0 && lds % 8 == 0 && ldt % 8 == 0 && ldd % 8 == 0, "attn_finish: channels / strides must be multiples of 8");
    const int64_t npix = (int64_t)N * h * h;
    if (npix == 0) return HOIG_OK;
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        if (k == 5)
            attn_finish_kernel<T, 25><<<ceil_div(npix, 4), 128, 0, as_stream(stream)>>>(
                (const T *)hidden, ldh, Chid, w2, b2, (const T *)src, lds, flow, (const T *)tgt, ldt, (T *)dst, ldd, npix, h, C,
                (const T *)unfold, ldu);
        else
            attn_finish_kernel<T, 9><<<ceil_div(npix, 4), 128, 0, as_stream(stream)>>>(
                (const T *)hidden, ldh, Chid, w2, b2, (const T *)src, lds, flow, (const T *)tgt, ldt, (T *)dst, ldd, npix, h, C,
                (const T *)unfold, ldu);
        return check_launch("attn_finish_kernel");
    });
}

extern "C" int hoig_grid_sample(const void *x, int64_t ldx, const float *grid, const void *tgt, int64_t ldt, void *dst,
                                int64_t ldd, int dtype, int N, int h, int C, hoigStream_t stream)
{
    HOIG_REQUIRE(x && grid && dst && C % 8 == 0 && ldx % 8 == 0 && ldd % 8 == 0, "grid_sample: bad argument");
    const int64_t npix = (int64_t)N * h * h;
    if (npix == 0) return HOIG_OK;
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        grid_sample_kernel<T><<<ceil_div(npix * (C / 8), TPB), TPB, 0, as_stream(stream)>>>(
            (const T *)x, ldx, grid, (const T *)tgt, ldt, (T *)dst, ldd, npix, h, C);
        return check_launch("grid_sample_kernel");
    });
}

extern "C" int hoig_composite(const float *img_bg, const float *obj, const float *hand, const float *mask_bg,
                              const float *mask_hand, float *out, int B, int HW, hoigStream_t stream)
{
    HOIG_REQUIRE(img_bg && obj && hand && mask_bg && mask_hand && out, "composite: null pointer");
    const int64_t n = (int64_t)B * 3 * HW;
    if (n == 0) return HOIG_OK;
    composite_kernel<<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>(img_bg, obj, hand, mask_bg, mask_hand, out, B, HW);
    return check_launch("composite_kernel");
}

extern "C" int hoig_block_extract_f32(const float *source, const float *flow, float *out, int B, int C, int Hs, int Ws, int Hf,
                                      int Wf, int k, hoigStream_t stream)
{
    HOIG_REQUIRE(source && flow && out && k >= 1, "block_extract: bad argument");
    const int64_t n = (int64_t)B * C * k * Hf * k * Wf;
    if (n == 0) return HOIG_OK;
    block_extract_kernel<<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>(source, flow, out, n, C, Hs, Ws, Hf, Wf, k);
    return check_launch("block_extract_kernel");
}

extern "C" int hoig_local_attn_reshape_f32(const float *in, float *out, int B, int k, int H, int W, hoigStream_t stream)
{
    HOIG_REQUIRE(in && out && k >= 1, "local_attn_reshape: bad argument");
    const int64_t n = (int64_t)B * k * H * k * W;
    if (n == 0) return HOIG_OK;
    local_attn_reshape_kernel<<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>(in, out, n, k, H, W);
    return check_launch("local_attn_reshape_kernel");
}

extern "C" int hoig_attn_unfold(const void *src, int64_t lds, const void *tgt, int64_t ldt, const float *flow, void *out,
                                int64_t ldo, int dtype, int N, int h, int C, int k, hoigStream_t stream)
{
    HOIG_REQUIRE(src && tgt && flow && out, "attn_unfold: null pointer");
    HOIG_REQUIRE(k == 5 || k == 3, "attn_unfold: kernel size %d not supported (3 or 5)", k);
    HOIG_REQUIRE(C % 8 == 0 && lds % 8 == 0 && ldt % 8 == 0 && ldo % 8 == 0 && ldo >= (int64_t)2 * k * k * C,
                 "attn_unfold: channels / strides must be multiples of 8 and ldo >= 2*k*k*C");
    const int64_t npix = (int64_t)N * h * h;
    if (npix == 0) return HOIG_OK;
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        if (k == 5)
            attn_unfold_kernel<T, 25><<<ceil_div(npix, 4), 256, 0, as_stream(stream)>>>((const T *)src, lds, (const T *)tgt, ldt, flow,
                                                                                       (T *)out, ldo, npix, h, C);
        else
            attn_unfold_kernel<T, 9><<<ceil_div(npix, 4), 256, 0, as_stream(stream)>>>((const T *)src, lds, (const T *)tgt, ldt, flow,
                                                                                      (T *)out, ldo, npix, h, C);
        return check_launch("attn_unfold_kernel");
    });
}

extern "C" int hoig_hunfold_nchw(const float *src, int B, int C, int H, int W, int k, void *dst, int64_t ldd, int Cpad, int dtype,
                                 hoigStream_t stream)
{
    HOIG_REQUIRE(src && dst && k >= 1 && (k & 1) && Cpad % 8 == 0 && Cpad >= k * C && ldd >= Cpad && ldd % 8 == 0, "hunfold: bad argument");
    const int64_t n = (int64_t)B * H * W;
    if (n == 0) return HOIG_OK;
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        const int chunks = Cpad / 8;
        if (W % 128 == 0 && C <= 16 && k <= 7 && Cpad <= 256 && 256 % chunks == 0) {
            const int rows = H % 8 == 0 ? 8 : (H % 4 == 0 ? 4 : 1);
            hunfold_row_kernel<T><<<(unsigned)(n / 128 / rows), 256, 0, as_stream(stream)>>>(src, C, H, W, k, (T *)dst, ldd, Cpad, rows);
        }
        else
            hunfold_kernel<T><<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>(src, B, C, H, W, k, (T *)dst, ldd, Cpad);
        return check_launch("hunfold_kernel");
    });
}

extern "C" int hoig_hfold_nchw(const void *z, int64_t ldz, int dtype, int B, int H, int W, int G, int k, const int *act_table,
                               int nseg, float *const *outs, const int *seg_c0, const int *seg_n, hoigStream_t stream)
{
    HOIG_REQUIRE(z && outs && seg_c0 && seg_n && nseg >= 1 && nseg <= 4 && k >= 1 && (k & 1) && ldz >= (int64_t)k * G, "hfold: bad argument");
    FoldSegs segs;
    segs.nseg = nseg;
    for (int q = 0; q < 4; ++q) {
        segs.out[q] = q < nseg ? outs[q] : nullptr;
        segs.c0[q] = q < nseg ? seg_c0[q] : 0;
        segs.n[q] = q < nseg ? seg_n[q] : 0;
        HOIG_REQUIRE(q >= nseg || outs[q], "hfold: null output");
    }
    const int64_t n = (int64_t)B * G * H * W;
    if (n == 0) return HOIG_OK;
    FoldSegs8 s8;
    for (int g = 0; g < 8; ++g) {
        s8.out[g] = nullptr; s8.bstride[g] = 0;
        for (int q = 0; q < nseg; ++q)
            if (g >= seg_c0[q] && g < seg_c0[q] + seg_n[q]) {
                s8.out[g] = outs[q] + (int64_t)(g - seg_c0[q]) * H * W;
                s8.bstride[g] = (int64_t)seg_n[q] * H * W;
            }
    }
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        if (G == 8 && ldz % 8 == 0)
            hfold8_kernel<T><<<ceil_div((int64_t)B * H * W, TPB), TPB, 0, as_stream(stream)>>>((const T *)z, ldz, B, H, W, k, act_table, s8);
        else
            hfold_kernel<T><<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>((const T *)z, ldz, B, H, W, G, k, act_table, segs);
        return check_launch("hfold_kernel");
    });
}

extern "C" int hoig_replicate_pad(const void *src, int64_t lds, void *dst, int64_t ldd, int dtype, int N, int h, int C, int pad,
                                  hoigStream_t stream)
{
    HOIG_REQUIRE(src && dst && pad >= 0, "replicate_pad: bad argument");
    HOIG_REQUIRE(C % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && lds >= C && ldd >= C, "replicate_pad: channels / strides must be multiples of 8");
    const int64_t n = (int64_t)N * (h + 2 * pad) * (h + 2 * pad) * (C / 8);
    if (n == 0) return HOIG_OK;
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        replicate_pad_kernel<T><<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>((const T *)src, lds, (T *)dst, ldd, N, h, C, pad);
        return check_launch("replicate_pad_kernel");
    });
}

static int g_attn_debug = 0;
static int g_attn_phase1_tc = 1;      // HOIG_ATTN_PHASE1_TC: logits of the staged path as a warp-level GEMM (1) or per-pixel FMAs (0)
static int g_attn_tc = 1;      // HOIG_ATTN_TC / hoig_set_attn_tc_mode: 1 = tensor-core attn_combine with staged source windows, 0 = per-pixel gathers
extern "C" void hoig_set_attn_tc_mode(int on) { g_attn_tc = on ? 1 : 0; }

extern "C" int hoig_attn_combine(const void *gt, int64_t ldgt, const void *gs, int64_t ldgs, int Chid, const float *b1, const float *w2,
                                 const float *b2, const void *src, int64_t lds, const float *flow, const void *tgt, int64_t ldt,
                                 void *dst, int64_t ldd, int dtype, int N, int h, int C, int k, hoigStream_t stream)
{
    HOIG_REQUIRE(gt && gs && b1 && w2 && b2 && src && flow && tgt && dst, "attn_combine: null pointer");
    HOIG_REQUIRE(Chid == 128, "attn_combine: hidden width %d not supported (extract_attn.py:11 uses 128)", Chid);
    HOIG_REQUIRE(k == 5 || k == 3, "attn_combine: kernel size %d not supported (3 or 5)", k);
    HOIG_REQUIRE(C % 8 == 0 && lds % 8 == 0 && ldt % 8 == 0 && ldd % 8 == 0 && ldgt % 4 == 0 && ldgs % 4 == 0,
                 "attn_combine: channels / strides must be multiples of 8");
    const int64_t npix = (int64_t)N * h * h;
    if (npix == 0) return HOIG_OK;
    const int sms = device_sm_count();
    const int64_t want = (npix + 31) / 32;
    const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    static bool env_parsed = false;
    if (!env_parsed) {
        env_parsed = true;
        const char *e = getenv("HOIG_ATTN_TC");
        if (e) g_attn_tc = atoi(e) ? 1 : 0;
        const char *dbg = getenv("HOIG_ATTN_DEBUG");       // timing knock-outs (garbage results): 1 = no patch-sum GEMM, 2 = no per-pixel phase
        if (dbg) g_attn_debug = atoi(dbg);
        const char *p1 = getenv("HOIG_ATTN_PHASE1_TC");
        if (p1) g_attn_phase1_tc = atoi(p1) ? 1 : 0;
    }
    if (g_attn_tc && (dtype == HOIG_BF16 || dtype == HOIG_F16) && k == 5 && h % ATC_TILE == 0 && C % 64 == 0) {
        // tensor-core formulation with shared-memory staging of the source window (attn_combine_tc_kernel)
        const int64_t tiles = (int64_t)N * (h / ATC_TILE) * (h / ATC_TILE);
        const int tgrid = (int)(tiles < (int64_t)sms * 2 ? tiles : (int64_t)sms * 2);
        if (dtype == HOIG_F16) {
            if (first_use_on_device(SLOT_ATTN_TC_F16) &&
                cudaFuncSetAttribute(attn_combine_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM) != cudaSuccess)
                return check_launch("attn_combine_tc smem attribute");
            attn_combine_tc_kernel<__half><<<tgrid, 256, ATC_SMEM, as_stream(stream)>>>(
                (const __half *)gt, ldgt, (const __half *)gs, ldgs, b1, w2, b2, (const __half *)src, lds, flow, (const __half *)tgt, ldt, (__half *)dst, ldd, N, h, C, g_attn_debug, g_attn_phase1_tc);
        } else {
            if (first_use_on_device(SLOT_ATTN_TC_BF16) &&
                cudaFuncSetAttribute(attn_combine_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM) != cudaSuccess)
                return check_launch("attn_combine_tc smem attribute");
            attn_combine_tc_kernel<__nv_bfloat16><<<tgrid, 256, ATC_SMEM, as_stream(stream)>>>(
                (const __nv_bfloat16 *)gt, ldgt, (const __nv_bfloat16 *)gs, ldgs, b1, w2, b2, (const __nv_bfloat16 *)src, lds, flow, (const __nv_bfloat16 *)tgt, ldt,
                (__nv_bfloat16 *)dst, ldd, N, h, C, g_attn_debug, g_attn_phase1_tc);
        }
        return check_launch("attn_combine_tc_kernel");
    }
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        if (k == 5)
            attn_combine_kernel<T, 25><<<grid, 256, 0, as_stream(stream)>>>((const T *)gt, ldgt, (const T *)gs, ldgs, b1, w2, b2, (const T *)src,
                                                                           lds, flow, (const T *)tgt, ldt, (T *)dst, ldd, npix, h, C);
        else
            attn_combine_kernel<T, 9><<<grid, 256, 0, as_stream(stream)>>>((const T *)gt, ldgt, (const T *)gs, ldgs, b1, w2, b2, (const T *)src,
                                                                          lds, flow, (const T *)tgt, ldt, (T *)dst, ldd, npix, h, C);
        return check_launch("attn_combine_kernel");
    });
}

extern "C" int hoig_seg_unfold3(const float *seg, int B, int C, int Hi, int Wi, void *dst, int64_t ldd, int Kpad, int Ho, int Wo,
                                int dtype, hoigStream_t stream)
{
    HOIG_REQUIRE(seg && dst && C > 0 && Kpad % 8 == 0 && Kpad >= 9 * C && ldd >= Kpad && ldd % 8 == 0, "seg_unfold3: bad argument");
    const int64_t n = (int64_t)B * Ho * Wo * (Kpad / 8);
    if (n == 0) return HOIG_OK;
    return dispatch(dtype, [&](auto *tag) {
        using T = std::remove_pointer_t<decltype(tag)>;
        const bool staged = C <= 16 && Kpad <= 256 && 256 % (Kpad / 8) == 0;
        if (staged && Wo % 64 == 0)
            seg_unfold3_row_kernel<T, 64><<<(unsigned)((int64_t)B * Ho * (Wo / 64)), 256, 0, as_stream(stream)>>>(seg, C, Hi, Wi, (T *)dst, ldd,
                                                                                                             Kpad, Ho, Wo);
        else if (staged && Wo % 32 == 0)
            seg_unfold3_row_kernel<T, 32><<<(unsigned)((int64_t)B * Ho * (Wo / 32)), 256, 0, as_stream(stream)>>>(seg, C, Hi, Wi, (T *)dst, ldd,
                                                                                                             Kpad, Ho, Wo);
        else
            seg_unfold3_kernel<T><<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>(seg, B, C, Hi, Wi, (T *)dst, ldd, Kpad, Ho, Wo);
        return check_launch("seg_unfold3_kernel");
    });
}

extern "C" int hoig_block_extract_backward_f32(const float *source, const float *flow, const float *grad_out, float *grad_source,
                                               float *grad_flow, int B, int C, int Hs, int Ws, int Hf, int Wf, int k, hoigStream_t stream)
{
    HOIG_REQUIRE(source && flow && grad_out && grad_source && grad_flow && k >= 1, "block_extract_backward: bad argument");
    const int64_t n = (int64_t)B * C * k * Hf * k * Wf, ncells = (int64_t)B * Hf * Wf;
    if (n == 0) return HOIG_OK;
    block_extract_bwd_src_kernel<<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>(flow, grad_out, grad_source, n, C, Hs, Ws, Hf, Wf, k);
    block_extract_bwd_flow_kernel<<<ceil_div(ncells * 32, 128), 128, 0, as_stream(stream)>>>(source, flow, grad_out, grad_flow, ncells, C, Hs, Ws,
                                                                                          Hf, Wf, k);
    return check_launch("block_extract_backward");
}

extern "C" int hoig_local_attn_reshape_backward_f32(const float *grad_out, float *grad_in, int B, int k, int H, int W, hoigStream_t stream)
{
    HOIG_REQUIRE(grad_out && grad_in && k >= 1, "local_attn_reshape_backward: bad argument");
    const int64_t n = (int64_t)B * k * H * k * W;
    if (n == 0) return HOIG_OK;
    local_attn_reshape_bwd_kernel<<<ceil_div(n, TPB), TPB, 0, as_stream(stream)>>>(grad_out, grad_in, n, k, H, W);
    return check_launch("local_attn_reshape_bwd_kernel");
}
