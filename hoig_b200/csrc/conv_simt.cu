// conv_simt.cu -- fp32 parity path of the implicit-GEMM convolution (SIMT FFMA).
//
// Same problem description, operand gather and epilogue contract as the
// tcgen05 kernel (conv_umma.cu); used for dtype HOIG_F32 (max-abs 1e-3 parity
// configuration, BASELINE config 1).  Also instantiated for bf16 storage so
// tests can cross-check the tensor-core kernel layer by layer.
//
// Tile: 64 pixels x 64 output channels per CTA, K step 16, 256 threads, 4x4
// register micro-tile, fp32 accumulation in k order within a CTA.
#include "conv_common.cuh"

namespace hoig {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(THREADS)
conv_simt_kernel(const ConvParams p)
{
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    __shared__ float red[2][16][BN];

    const int tid = threadIdx.x;
    const int n_img = blockIdx.x / p.tiles_per_image;
    const int tile = blockIdx.x % p.tiles_per_image;
    const int pix0 = tile * BM;
    const int n0 = blockIdx.y * BN;
    const int npix = p.OH * p.OW;

    // loader roles: threads 0..127 gather A chunks (64 rows x 2 chunks), 128..255 load B chunks
    const int lrow = (tid & 127) >> 1, lchunk = tid & 1;
    int a_oy = 0, a_ox = 0;
    bool a_valid = false;
    if (tid < 128) {
        const int pix = pix0 + lrow;
        a_valid = pix < npix;
        a_oy = a_valid ? pix / p.OW : 0;
        a_ox = a_valid ? pix % p.OW : 0;
    }
    const T *wrow = static_cast<const T *>(p.weight) + (int64_t)min(n0 + lrow, p.Npad - 1) * p.Kpad;
    const bool b_valid = (n0 + lrow) < p.Npad;

    const int ty = tid / 16, tx = tid % 16;  // micro-tile: rows ty*4.., cols tx*4..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.Kpad; k0 += BK) {
        float v[8];
        if (tid < 128) {
            if (a_valid) gather_chunk<T>(p, n_img, a_oy, a_ox, k0 / 8 + lchunk, v);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) As[lchunk * 8 + j][lrow] = v[j];
        } else {
            if (b_valid) load8(wrow + k0 + lchunk * 8, v);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) Bs[lchunk * 8 + j][lrow] = v[j];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // epilogue: bias, residual, activation, store, statistics of the stored values
    float csum[4] = {0.f, 0.f, 0.f, 0.f}, csq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pix = pix0 + ty * 4 + i;
        if (pix >= npix) continue;
        const int64_t m = (int64_t)n_img * npix + pix;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.Cout) continue;
            float val = acc[i][j];
            if (p.bias) val += p.bias[n];
            if (p.residual) val += DT<T>::ld(static_cast<const T *>(p.residual) + m * p.ldr + n);
            val = apply_act(val, p.act);
            val = round_to<T>(val);
            DT<T>::st(static_cast<T *>(p.dst) + m * p.ldd + n, val);
            csum[j] += val;
            csq[j] += val * val;
        }
    }
    if (p.stats) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[0][ty][tx * 4 + j] = csum[j];
            red[1][ty][tx * 4 + j] = csq[j];
        }
        __syncthreads();
        if (tid < 2 * BN) {
            const int which = tid / BN, col = tid % BN;
            const int n = n0 + col;
            if (n < p.Cout) {
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < 16; ++r) s += red[which][r][col];
                atomicAdd(&p.stats[((int64_t)n_img * p.Cout + n) * 2 + which], (double)s);
            }
        }
    }
}

}  // namespace

int fill_conv_params(const hoigConvDesc *d, int bm, ConvParams *p)
{
    HOIG_REQUIRE(d && d->src0 && d->weight && d->dst, "conv2d: null pointer");
    HOIG_REQUIRE(d->dtype == HOIG_F32 || d->dtype == HOIG_BF16, "conv2d: bad dtype %d", d->dtype);
    HOIG_REQUIRE(d->mode >= HOIG_CONV && d->mode <= HOIG_CONV_LOCAL_ATTN, "conv2d: bad mode %d", d->mode);
    HOIG_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->OH > 0 && d->OW > 0 && d->Cout > 0, "conv2d: bad shape");
    HOIG_REQUIRE(d->C0 > 0 && d->C0 % 8 == 0 && d->C1 >= 0 && d->C1 % 8 == 0, "conv2d: C0/C1 must be multiples of 8 (got %d,%d)", d->C0, d->C1);
    HOIG_REQUIRE(d->C1 == 0 || d->src1, "conv2d: C1 > 0 needs src1");
    HOIG_REQUIRE(d->KH > 0 && d->KW > 0 && d->stride > 0 && d->pad >= 0, "conv2d: bad kernel geometry");
    HOIG_REQUIRE(d->ld0 >= d->C0 && d->ld0 % 8 == 0 && (d->C1 == 0 || (d->ld1 >= d->C1 && d->ld1 % 8 == 0)), "conv2d: source pixel stride must be a multiple of 8 and >= channels");
    HOIG_REQUIRE(d->ldd >= d->Cout, "conv2d: ldd < Cout");
    HOIG_REQUIRE(!d->residual || d->ldr >= d->Cout, "conv2d: ldr < Cout");
    const int esz = d->dtype == HOIG_F32 ? 4 : 2;
    HOIG_REQUIRE(((uintptr_t)d->src0 % 16) == 0 && (!d->src1 || ((uintptr_t)d->src1 % 16) == 0) && ((uintptr_t)d->weight % 16) == 0,
                 "conv2d: sources and weights must be 16-byte aligned");
    (void)esz;
    if (d->mode == HOIG_CONV) {
        HOIG_REQUIRE(d->OH == (d->H + 2 * d->pad - d->KH) / d->stride + 1 && d->OW == (d->W + 2 * d->pad - d->KW) / d->stride + 1,
                     "conv2d: output size does not match geometry");
    } else if (d->mode == HOIG_CONV_LOCAL_ATTN) {
        HOIG_REQUIRE(d->flow && d->src1 && d->C0 == d->C1 && d->OH == d->H && d->OW == d->W && d->KH == d->KW,
                     "conv2d(local_attn): needs flow, src1, C0 == C1, OH == H");
    }
    p->mode = d->mode;
    p->N = d->N; p->H = d->H; p->W = d->W; p->C0 = d->C0; p->C1 = d->C1; p->Cin = d->C0 + d->C1;
    p->OH = d->OH; p->OW = d->OW; p->Cout = d->Cout;
    p->KH = d->KH; p->KW = d->KW; p->stride = d->stride; p->pad = d->pad;
    p->K = d->KH * d->KW * p->Cin;
    int rows, cols;
    hoig_conv_packed_dims(d->Cout, d->KH, d->KW, p->Cin, &rows, &cols);
    p->Npad = rows; p->Kpad = cols;
    p->src0 = d->src0; p->ld0 = d->ld0; p->src1 = d->src1; p->ld1 = d->ld1;
    p->weight = d->weight; p->bias = d->bias; p->act = d->act;
    p->residual = d->residual; p->ldr = d->ldr; p->dst = d->dst; p->ldd = d->ldd;
    p->stats = d->stats; p->flow = d->flow;
    p->tiles_per_image = ceil_div((int64_t)d->OH * d->OW, bm);
    return HOIG_OK;
}

int conv2d_simt(const hoigConvDesc *d, cudaStream_t stream)
{
    ConvParams p;
    const int st = fill_conv_params(d, BM, &p);
    if (st != HOIG_OK) return st;
    dim3 grid((unsigned)(p.N * p.tiles_per_image), (unsigned)ceil_div(p.Cout, BN));
    if (d->dtype == HOIG_F32) conv_simt_kernel<float><<<grid, THREADS, 0, stream>>>(p);
    else conv_simt_kernel<__nv_bfloat16><<<grid, THREADS, 0, stream>>>(p);
    return check_launch("conv_simt_kernel");
}

}  // namespace hoig

extern "C" int hoig_conv_packed_dims(int Cout, int KH, int KW, int Cin, int *rows, int *cols)
{
    if (rows) *rows = (Cout + 15) / 16 * 16;
    if (cols) *cols = (KH * KW * Cin + 63) / 64 * 64;
    return HOIG_OK;
}
