// conv_common.cuh -- implicit-GEMM problem description shared by the SIMT fp32
// kernel (conv_simt.cu) and the tcgen05 kernel (conv_umma.cu).
//
// GEMM view of one launch:  D[m][n] = sum_k A[m][k] * Wp[n][k]
//   m = pixel of the launch grid (image-major; an M tile never straddles two images so
//       per-plane statistics reduce inside a tile)
//   n = output channel
//   k = t*Cin + c,  t = tap index into the launch's tap table,  Cin = C0 + C1
// A is never materialised.  A tap is an input offset (dy, dx) plus the id of the input view
// ("phase") it reads:   input pixel = (gy*stride + dy, gx*stride + dx) of view `tap_map[t]`.
// That one table covers
//   nn.Conv2d              taps (r - pad, s - pad); stride-2 convs read the four parity
//                          sub-images of the input as four strided views with offsets in {-1,0},
//                          which makes them stride-1 (TMA-box friendly) problems;
//   nn.ConvTranspose2d     ONE stride-1 conv over the INPUT grid with the 2x2 taps (dy,dx) in {0,1}^2 and
//                          N = 4*Cout columns, one block per output parity (a,b); block (a,b) lands at
//                          (2*gy+a, 2*gx+b).  Taps a parity does not use carry zero weights (9 of 16
//                          blocks are live): no zero-insertion, 2.25x fewer MACs than the gather
//                          formulation and a wide-N tile for the tensor core;
//   local attention        5x5 taps of BlockExtractor(tgt,0) | BlockExtractor(src,flow)
//                          (extract_attn.py:24-26 + block_extractor_kernel.cu:52-84).
#pragma once
#include "common.cuh"

namespace hoig {

constexpr int kMaxTaps = 64;
constexpr int kMaxViews = 4;

struct InputView {          // strided NHWC view of a source tensor
    const void *base;
    int H, W;               // extent of the view
    int64_t sx, sy, sn;     // element strides between view pixels / rows / images
};

struct ConvParams {
    int mode;               // HOIG_CONV (also used for transposed phases) or HOIG_CONV_LOCAL_ATTN
    int N, C0, C1, Cin;
    int GH, GW;             // launch grid (pixels enumerated by m)
    int Cout;
    int ntaps, stride;      // stride of the input walk (1 after phase decomposition)
    int8_t tap_dy[kMaxTaps], tap_dx[kMaxTaps], tap_map[kMaxTaps];
    InputView view[kMaxViews];   // views of src0 (channels [0,C0))
    InputView view1;             // single view of src1 (channels [C0,Cin)), only with nviews == 1
    int nviews;
    int K, Kpad, Npad;
    const void *weight;     // [Npad][ldw], this launch's columns start at `weight`
    int64_t ldw;
    const float *bias;
    int act;
    const int *act_table;   // optional per-output-channel activation codes
    const void *residual; int64_t ldr;
    void *dst; int64_t ldd;
    int OHf, OWf, os, ooy, oox;   // output pixel = (gy*os + ooy, gx*os + oox) in an OHf x OWf image
    int phase_cout;               // > 0 (transposed conv): GEMM column n = phase*phase_cout + channel, phase (a,b) = (n/pc >> 1, & 1)
                                  //      lands at pixel (gy*2 + a, gx*2 + b); Cout then counts all 4 phases
    double *stats;
    const float *flow;
    int KH;                 // local attention: kernel size (5)
    int tiles_per_image;
    int kh, kw, pad_h, pad_w;   // regular stride-1 tap grid (tap (r,s) = offset (r - pad_h, s - pad_w) of view 0), else kw = 0
    const void *spade_x; int64_t ld_spade_x;   // SPADE-modulating epilogue (hoigConvDesc::spade_x), else NULL
    const double *spade_stats;
    float spade_eps;
};

// Bilinear tap set of BlockExtractor for one (pixel, k x k tap): indices into the
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

// element offset of pixel (n, y, x) in a view
__device__ __forceinline__ int64_t view_off(const InputView &v, int n, int y, int x)
{
    return (int64_t)n * v.sn + (int64_t)y * v.sy + (int64_t)x * v.sx;
}

// Gather 8 consecutive k (one chunk) of the A row of grid pixel (n_img, gy, gx) as floats.
template <typename T>
__device__ __forceinline__ void gather_chunk(const ConvParams &p, int n_img, int gy, int gx, int kchunk, float v[8])
{
    const int k0 = kchunk * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (k0 >= p.K) return;
    const int tap = k0 / p.Cin;
    int c = k0 - tap * p.Cin;
    const bool second = c >= p.C0;
    if (p.mode != HOIG_CONV_LOCAL_ATTN) {
        const InputView &vw = second ? p.view1 : p.view[p.tap_map[tap]];
        if (second) c -= p.C0;
        const int iy = gy * p.stride + p.tap_dy[tap], ix = gx * p.stride + p.tap_dx[tap];
        if (iy < 0 || iy >= vw.H || ix < 0 || ix >= vw.W) return;
        load8(static_cast<const T *>(vw.base) + view_off(vw, n_img, iy, ix) + c, v);
    } else {  // view[0] = target (zero flow), view1 = source (flow)
        const int r = tap / p.KH, s = tap - r * p.KH;
        if (!second) {
            const InputView &vw = p.view[0];
            const int iy = max(min(gy + r - p.KH / 2, vw.H - 1), 0), ix = max(min(gx + s - p.KH / 2, vw.W - 1), 0);
            load8(static_cast<const T *>(vw.base) + view_off(vw, n_img, iy, ix) + c, v);
        } else {
            const InputView &vw = p.view1;
            c -= p.C0;
            const float *fl = p.flow + (((int64_t)n_img * p.GH + gy) * p.GW + gx) * 2;
            const BETap t = be_tap(fl[0], fl[1], gy, gx, r, s, p.KH, vw.H, vw.W);
            const T *base = static_cast<const T *>(vw.base) + (int64_t)n_img * vw.sn + c;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float u[8];
                load8(base + (int64_t)t.idx[q] * vw.sx, u);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __fmaf_rn(t.w[q], u[j], v[j]);
            }
        }
    }
}

// output element offset of grid pixel (n_img, pix) for channel 0
__device__ __forceinline__ int64_t out_pixel(const ConvParams &p, int n_img, int pix)
{
    if (p.os == 1) return (int64_t)n_img * p.GH * p.GW + pix;
    const int gy = pix / p.GW, gx = pix - gy * p.GW;
    return ((int64_t)n_img * p.OHf + gy * p.os + p.ooy) * p.OWf + gx * p.os + p.oox;
}

__device__ __forceinline__ int act_of(const ConvParams &p, int n) { return p.act_table ? p.act_table[n] : p.act; }

// Host side: expands a hoigConvDesc into 1 launch (conv / local attention) or 4 (transposed).
struct ConvPlan {
    int n;
    ConvParams launch[4];
};
int plan_conv(const hoigConvDesc *d, int BM, ConvPlan *plan);  // validates; returns hoigStatus
// columns of the packed weight matrix occupied by launch `i` of the plan for this geometry
void packed_layout(int mode, int Cout, int KH, int KW, int Cin, int stride, int pad, int *rows, int *cols, int col_off[4],
                   int col_len[4]);

}  // namespace hoig
