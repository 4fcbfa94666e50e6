// conv_common.cuh -- implicit-GEMM problem description shared by the SIMT fp32
// kernel (conv_simt.cu) and the tcgen05 kernel (conv_umma.cu).
//
// GEMM view:  D[m][n] = sum_k A[m][k] * Wp[n][k]
//   m = output pixel (image-major: m = n_img*OH*OW + oy*OW + ox; an M tile never
//       straddles two images so per-plane statistics reduce inside a tile)
//   n = output channel
//   k = tap*Cin + c,  tap = r*KW + s,  Cin = C0 + C1 (c < C0 reads src0, else src1)
// A is never materialised: 8-channel chunks (k multiple of 8) are gathered on
// the fly by mode:
//   HOIG_CONV            iy = oy*stride - pad + r                       (nn.Conv2d)
//   HOIG_CONV_TRANSPOSED iy = (oy + pad - r)/stride when divisible      (nn.ConvTranspose2d)
//   HOIG_CONV_LOCAL_ATTN 5x5 taps of BlockExtractor(tgt,0) | BlockExtractor(src,flow)
//                        (extract_attn.py:24-26 + block_extractor_kernel.cu:52-84)
#pragma once
#include "common.cuh"

namespace hoig {

struct ConvParams {
    int mode;
    int N, H, W, C0, C1, Cin;
    int OH, OW, Cout;
    int KH, KW, stride, pad;
    int K;       // KH*KW*Cin (logical)
    int Kpad;    // padded to 64
    int Npad;    // Cout padded to 16
    const void *src0; int64_t ld0;
    const void *src1; int64_t ld1;
    const void *weight;
    const float *bias;
    int act;
    const void *residual; int64_t ldr;
    void *dst; int64_t ldd;
    double *stats;
    const float *flow;
    int tiles_per_image;  // ceil(OH*OW / BM)
};

// Bilinear tap set of BlockExtractor for one (pixel, 5x5 tap): indices into the
// (H,W) source plane and the four products xP*yP, in the kernel's order LT,RT,LB,RB
// (block_extractor_kernel.cu:57-82, same float op order).
struct BETap {
    int idx[4];
    float w[4];
};
__device__ __forceinline__ BETap be_tap(float flow_x, float flow_y, int yf, int xf, int ky, int kx, int k, int Hs, int Ws)
{
    const float fy = __fadd_rn(flow_y, (float)(ky - k / 2));
    const float fx = __fadd_rn(flow_x, (float)(kx - k / 2));
    const float dy = __fadd_rn(fy, (float)yf);
    const float dx = __fadd_rn(fx, (float)xf);
    const float fdx = floorf(dx), fdy = floorf(dy);
    const int xL = max(min((int)fdx, Ws - 1), 0);
    const int xR = max(min((int)__fadd_rn(fdx, 1.f), Ws - 1), 0);
    const int yT = max(min((int)fdy, Hs - 1), 0);
    const int yB = max(min((int)__fadd_rn(fdy, 1.f), Hs - 1), 0);
    const float xLp = __fsub_rn(1.f, __fsub_rn(dx, fdx)), xRp = __fsub_rn(dx, fdx);
    const float yTp = __fsub_rn(1.f, __fsub_rn(dy, fdy)), yBp = __fsub_rn(dy, fdy);
    BETap t;
    t.idx[0] = yT * Ws + xL; t.idx[1] = yT * Ws + xR; t.idx[2] = yB * Ws + xL; t.idx[3] = yB * Ws + xR;
    t.w[0] = __fmul_rn(xLp, yTp); t.w[1] = __fmul_rn(xRp, yTp); t.w[2] = __fmul_rn(xLp, yBp); t.w[3] = __fmul_rn(xRp, yBp);
    return t;
}

// Gather 8 consecutive k (one chunk) of A row (n_img, oy, ox) as floats.
template <typename T>
__device__ __forceinline__ void gather_chunk(const ConvParams &p, int n_img, int oy, int ox, int kchunk, float v[8])
{
    const int k0 = kchunk * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (k0 >= p.K) return;
    const int tap = k0 / p.Cin;
    int c = k0 - tap * p.Cin;
    const int r = tap / p.KW, s = tap - r * p.KW;
    const T *base;
    int64_t ld;
    const bool second = c >= p.C0;
    if (second) { base = static_cast<const T *>(p.src1); ld = p.ld1; c -= p.C0; }
    else        { base = static_cast<const T *>(p.src0); ld = p.ld0; }
    if (p.mode == HOIG_CONV) {
        const int iy = oy * p.stride - p.pad + r, ix = ox * p.stride - p.pad + s;
        if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) return;
        load8(base + ((int64_t)(n_img * p.H + iy) * p.W + ix) * ld + c, v);
    } else if (p.mode == HOIG_CONV_TRANSPOSED) {
        const int ty = oy + p.pad - r, tx = ox + p.pad - s;
        if (ty < 0 || tx < 0 || (ty % p.stride) || (tx % p.stride)) return;
        const int iy = ty / p.stride, ix = tx / p.stride;
        if (iy >= p.H || ix >= p.W) return;
        load8(base + ((int64_t)(n_img * p.H + iy) * p.W + ix) * ld + c, v);
    } else {  // HOIG_CONV_LOCAL_ATTN: src0 = target (zero flow), src1 = source (flow)
        const int64_t plane = (int64_t)n_img * p.H * p.W;
        if (!second) {
            const int iy = max(min(oy + r - p.KH / 2, p.H - 1), 0), ix = max(min(ox + s - p.KW / 2, p.W - 1), 0);
            load8(base + (plane + (int64_t)iy * p.W + ix) * ld + c, v);
        } else {
            const float *fl = p.flow + (plane + (int64_t)oy * p.W + ox) * 2;
            const BETap t = be_tap(fl[0], fl[1], oy, ox, r, s, p.KH, p.H, p.W);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float u[8];
                load8(base + (plane + t.idx[q]) * ld + c, u);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __fmaf_rn(t.w[q], u[j], v[j]);
            }
        }
    }
}

int fill_conv_params(const hoigConvDesc *d, int BM, ConvParams *p);  // validates; returns hoigStatus

}  // namespace hoig
