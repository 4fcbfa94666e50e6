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
    const int npix = p.GH * p.GW;

    // loader roles: threads 0..127 gather A chunks (64 rows x 2 chunks), 128..255 load B chunks
    const int lrow = (tid & 127) >> 1, lchunk = tid & 1;
    int a_oy = 0, a_ox = 0;
    bool a_valid = false;
    if (tid < 128) {
        const int pix = pix0 + lrow;
        a_valid = pix < npix;
        a_oy = a_valid ? pix / p.GW : 0;
        a_ox = a_valid ? pix % p.GW : 0;
    }
    const T *wrow = static_cast<const T *>(p.weight) + (int64_t)min(n0 + lrow, p.Npad - 1) * p.ldw;
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
        const int64_t m0 = out_pixel(p, n_img, pix);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ng = n0 + tx * 4 + j;      // GEMM column
            if (ng >= p.Cout) continue;
            int n = ng;                           // output channel
            int64_t m = m0;
            if (p.phase_cout) {                   // transposed conv: column = phase * Cout + channel
                const int ph = ng / p.phase_cout;
                n = ng - ph * p.phase_cout;
                m = m0 + (int64_t)(ph >> 1) * p.OWf + ((ph & 1) ^ (ph >> 1));   // block order (0,0),(0,1),(1,1),(1,0)
            }
            float val = acc[i][j];
            if (p.bias) val += p.bias[n];
            if (p.residual) val += DT<T>::ld(static_cast<const T *>(p.residual) + m * p.ldr + n);
            val = apply_act(val, act_of(p, n));
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
                const int pc = p.phase_cout;
                atomicAdd(&p.stats[((int64_t)n_img * (pc ? pc : p.Cout) + (pc ? n % pc : n)) * 2 + which], (double)s);
            }
        }
    }
}

}  // namespace

int conv2d_simt(const hoigConvDesc *d, cudaStream_t stream)
{
    HOIG_REQUIRE(!d || !d->spade_x, "conv2d: the SPADE-modulating epilogue exists on the tensor-core path only");
    ConvPlan plan;
    const int st = plan_conv(d, BM, &plan);
    if (st != HOIG_OK) return st;
    for (int i = 0; i < plan.n; ++i) {
        const ConvParams &p = plan.launch[i];
        dim3 grid((unsigned)(p.N * p.tiles_per_image), (unsigned)ceil_div(p.Cout, BN));
        if (d->dtype == HOIG_F32) conv_simt_kernel<float><<<grid, THREADS, 0, stream>>>(p);
        else if (d->dtype == HOIG_F16) conv_simt_kernel<__half><<<grid, THREADS, 0, stream>>>(p);
        else conv_simt_kernel<__nv_bfloat16><<<grid, THREADS, 0, stream>>>(p);
        const int rc = check_launch("conv_simt_kernel");
        if (rc != HOIG_OK) return rc;
    }
    return HOIG_OK;
}

}  // namespace hoig
