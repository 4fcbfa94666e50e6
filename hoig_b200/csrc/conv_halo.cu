// conv_halo.cu -- dense KH x KW stride-1 convolution over PADDED rasters on the 5th-gen tensor cores,
// with the activation tile (plus its horizontal halo) loaded ONCE per kernel row and reused by all KW taps.
//
// A padded raster is an NHWC tensor [N][Hp][Wp][ld] whose border already holds the padding values
// (replicated edge pixels for BlockExtractor's clamped taps, block_extractor_kernel.cu:62-69).  In that
// layout a 2-D convolution is a 1-D one over the pixel sequence m = (n*Hp + y)*Wp + x:
//
//     out[m][n] = sum_{r,s,c} W[n][(r*KW + s)*C + c] * in[m + (r - KH/2)*Wp + (s - KW/2)][c]
//
// so the A operand of tap (r,s) for the 128 pixels [m0, m0+128) is a CONTIGUOUS run of pixel rows.  One TMA
// box of 128 + KW-1 rows per (r, 64-channel block) lands in 128B-swizzled shared memory and the KW taps of
// that kernel row are issued as UMMAs whose A descriptors start s*128 bytes further on (one pixel row each).
// Compared with one box per tap (conv_umma.cu) this divides the L2->SM activation traffic by KW; the weight
// traffic is halved by giving each CTA TWO 128-row accumulators (BM = 256) that share every weight tile.
//
// Up to two problems ("segments") share one launch so the tail wave of one is filled by the other.
//
// Persistent, warp-specialised CTA (320 threads):
//   warp 0      TMA producer: A ring (2 x (128+KW-1) rows x 64 ch per stage) and B ring (BN x 64 per stage)
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue
//   warps 2-9   epilogue: warp w drains TMEM lane quadrant w % 4 of accumulator half (w - 2) / 4, raw fp32 ->
//               16-bit stores (no bias / activation: the consumer, hoig_attn_combine, adds them)
#include <type_traits>

#include "umma_common.cuh"

namespace hoig {
namespace {

constexpr int HM = 128;                      // rows per accumulator half
constexpr int HBM = 2 * HM;                  // rows per tile
constexpr int HBK = 64;
constexpr int A_HALF_BYTES = 17 * 1024;      // (128 + up to 7 halo rows) x 128 B, rounded up to the 1024 B swizzle period
constexpr int A_STAGE = 2 * A_HALF_BYTES;
constexpr int H_THREADS = 352;
constexpr int H_EPI_WARP0 = 2, H_EPI_WARPS = 8;
constexpr int H_MMA_WARP0 = 1, H_MMA_WARP1 = 10;   // one issuing warp per accumulator half
constexpr int H_MAX_STAGES = 8;
constexpr int H_SMEM_TOTAL = 216 * 1024;

struct HaloSeg {
    int64_t rows;      // N*Hp*Wp
    int pitch;         // Wp
    int cblocks;       // C / 64
    int C;
    void *dst;
    int64_t ldd;
    int tile0;         // first tile of this segment
    int st256;         // rows are 32-byte aligned: one 256-bit store per 16-column piece (a whole sector per lane)
};
struct HaloParams {
    HaloSeg seg[2];
    int nsegs, total_tiles;
    int KH, KW, BN, Cout;
    int a_stages, b_stages;
    int variant;       // 0: one box per kernel row, taps by shifted descriptors; 2: one box per tap (test hook)
};

template <typename T>
__global__ void __launch_bounds__(H_THREADS, 1)
conv_halo_kernel(const HaloParams P, const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_w0,
                 const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_w1)
{
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t afull[H_MAX_STAGES], aempty[H_MAX_STAGES], bfull[H_MAX_STAGES], bempty[H_MAX_STAGES];
    __shared__ __align__(8) uint64_t tfull[2], tempty[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int BN = P.BN;
    const uint32_t b_stage = (uint32_t)BN * HBK * 2;
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t a_base = smem_u32(smem);
    const uint32_t b_base = a_base + (uint32_t)P.a_stages * A_STAGE;
    const int hw = P.KW / 2, hh = P.KH / 2;
    const bool per_tap = P.variant == 2;

    if (threadIdx.x == 0) {
        // "empty" and "accumulator full" take one tcgen05.commit from each of the two MMA-issuing warps
        for (int s = 0; s < P.a_stages; ++s) { mbar_init(smem_u32(&afull[s]), 1); mbar_init(smem_u32(&aempty[s]), 2); }
        for (int s = 0; s < P.b_stages; ++s) { mbar_init(smem_u32(&bfull[s]), 1); mbar_init(smem_u32(&bempty[s]), 2); }
        for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&tfull[a]), 2); mbar_init(smem_u32(&tempty[a]), H_EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w0) : "memory");
        if (P.nsegs > 1) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a1) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ================================================================ TMA producer
        {   // whole warp runs the loop; one elected lane issues (see elect_one())
            const uint32_t box_bytes = (uint32_t)(HM + P.KW - 1) * 128u;
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;   // ring positions / phases, kept incrementally (no divisions on this thread)
            const uint32_t afull0 = smem_u32(&afull[0]), aempty0 = smem_u32(&aempty[0]);
            const uint32_t bfull0 = smem_u32(&bfull[0]), bempty0 = smem_u32(&bempty[0]);
            for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
                const int sg = (P.nsegs > 1 && tile >= P.seg[1].tile0) ? 1 : 0;
                const HaloSeg &S = P.seg[sg];
                const CUtensorMap *ma = sg ? &map_a1 : &map_a0, *mw = sg ? &map_w1 : &map_w0;
                const int64_t m0 = (int64_t)(tile - S.tile0) * HBM;
                for (int r = 0; r < P.KH; ++r)
                    for (int cb = 0; cb < S.cblocks; ++cb) {
                        for (int s = 0; s < P.KW; ++s) {
                            if (s == 0 || per_tap) {
                                mbar_wait(aempty0 + 8u * sa, pa ^ 1u);
                                const uint32_t bar = afull0 + 8u * sa;
                                const uint32_t dst = a_base + sa * (uint32_t)A_STAGE;
                                const int64_t row0 = m0 + (int64_t)(r - hh) * S.pitch - hw + (per_tap ? s : 0);
                                if (elect_one()) {
                                    mbar_arrive_expect_tx(bar, 2 * box_bytes);
                                    tma_load_2d(dst, ma, bar, cb * HBK, (int)row0);
                                    tma_load_2d(dst + A_HALF_BYTES, ma, bar, cb * HBK, (int)(row0 + HM));
                                }
                                __syncwarp();
                                if (++sa == (uint32_t)P.a_stages) { sa = 0; pa ^= 1u; }
                            }
                            mbar_wait(bempty0 + 8u * sb, pb ^ 1u);
                            const uint32_t bar = bfull0 + 8u * sb;
                            if (elect_one()) {
                                mbar_arrive_expect_tx(bar, b_stage);
                                tma_load_2d(b_base + sb * b_stage, mw, bar, ((r * P.KW + s) * S.cblocks + cb) * HBK, 0);
                            }
                            __syncwarp();
                            if (++sb == (uint32_t)P.b_stages) { sb = 0; pb ^= 1u; }
                        }
                    }
            }
        }
    } else if (warp == H_MMA_WARP0 || warp == H_MMA_WARP1) {
        // ================================================================== MMA issuers
        // One warp per 128-row accumulator half.  A single thread sustains one 128 x N x 16 MMA per ~65-100 cycles of issue
        // overhead, more than a narrow (N <= 128) MMA executes in (48-64 cycles, scripts/probes/mma_probe.cu); two
        // independent issue streams fill the tensor pipe.
        const int half = warp == H_MMA_WARP0 ? 0 : 1;
        {   // whole warp runs the loop; one elected lane issues
            constexpr uint32_t kFmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
            const uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(HM >> 4) << 24);
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0, tcount = 0;
            const uint32_t afull0 = smem_u32(&afull[0]), aempty0 = smem_u32(&aempty[0]);
            const uint32_t bfull0 = smem_u32(&bfull[0]), bempty0 = smem_u32(&bempty[0]);
            const uint64_t adesc0 = umma_desc(a_base), bdesc0 = umma_desc(b_base);   // + (byte offset >> 4)
            for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++tcount) {
                const int sg = (P.nsegs > 1 && tile >= P.seg[1].tile0) ? 1 : 0;
                const HaloSeg &S = P.seg[sg];
                const uint32_t acc = tcount & 1;
                mbar_wait(smem_u32(&tempty[acc]), ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)(2 * BN);
                uint32_t first = 1;
                for (int r = 0; r < P.KH; ++r)
                    for (int cb = 0; cb < S.cblocks; ++cb) {
                        uint64_t da = 0;
                        uint32_t a_st = 0;
                        for (int s = 0; s < P.KW; ++s) {
                            if (s == 0 || per_tap) {
                                a_st = sa;
                                mbar_wait(afull0 + 8u * sa, pa);
                                da = adesc0 + (uint64_t)(sa * (uint32_t)(A_STAGE >> 4));
                                if (++sa == (uint32_t)P.a_stages) { sa = 0; pa ^= 1u; }
                            }
                            const uint32_t b_st = sb;
                            mbar_wait(bfull0 + 8u * sb, pb);
                            if (++sb == (uint32_t)P.b_stages) { sb = 0; pb ^= 1u; }
                            tc_fence_after();
                            const uint64_t db = bdesc0 + (uint64_t)(b_st * (b_stage >> 4));
                            // Tap s reads the box s pixel rows (128 B each) further on.  Descriptor start addresses that are NOT
                            // multiples of the 1024 B swizzle period are fine as they are (base-offset field 0): the 128B swizzle
                            // is a function of the absolute smem address, for the TMA write and the UMMA read alike.  Measured on
                            // B200; setting base_offset = (addr >> 7) & 7 breaks it.
                            const uint64_t das = da + (uint64_t)(per_tap ? 0 : s * 8);
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < HBK / 16; ++k)
                                    umma_bf16(d_tmem + (uint32_t)(half * BN), das + (uint64_t)(half * (A_HALF_BYTES >> 4) + 2 * k),
                                              db + (uint64_t)(2 * k), idesc, (first && k == 0) ? 0u : 1u);
                                umma_commit(bempty0 + 8u * b_st);
                                if (s == P.KW - 1 || per_tap) umma_commit(aempty0 + 8u * a_st);
                            }
                            __syncwarp();
                            first = 0;
                        }
                    }
                if (elect_one()) umma_commit(smem_u32(&tfull[acc]));
                __syncwarp();
            }
        }
    } else if (warp >= H_EPI_WARP0 && warp < H_EPI_WARP0 + H_EPI_WARPS) {
        // ==================================================================== epilogue
        const int quad = warp & 3, half = (warp - H_EPI_WARP0) >> 2;   // warps 2-9
        const int n_chunks = P.Cout / 16;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++tcount) {
            const int sg = (P.nsegs > 1 && tile >= P.seg[1].tile0) ? 1 : 0;
            const HaloSeg &S = P.seg[sg];
            const int64_t row = (int64_t)(tile - S.tile0) * HBM + half * HM + quad * 32 + lane;
            const bool valid = row < S.rows;
            T *drow = static_cast<T *>(S.dst) + (valid ? row : 0) * S.ldd;
            const uint32_t acc = tcount & 1;
            mbar_wait(smem_u32(&tfull[acc]), (tcount >> 1) & 1);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (uint32_t)(2 * BN) + (uint32_t)(half * BN);
            uint32_t ra[16], rb[16];
            auto store16 = [&](const uint32_t (&r)[16], int ch) {
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) pk[j] = pack2<T>(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                if (valid) {
                    if (S.st256) {
                        st_global_v8(drow + ch * 16, pk);
                    } else {
                        *reinterpret_cast<uint4 *>(drow + ch * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4 *>(drow + ch * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
            };
            tmem_ld16(t_row, ra);
            for (int ch = 0; ch < n_chunks; ch += 2) {
                tmem_ld_wait(ra);
                if (ch + 1 < n_chunks) tmem_ld16(t_row + (uint32_t)((ch + 1) * 16), rb);
                store16(ra, ch);
                if (ch + 1 < n_chunks) {
                    tmem_ld_wait(rb);
                    if (ch + 2 < n_chunks) tmem_ld16(t_row + (uint32_t)((ch + 2) * 16), ra);
                    store16(rb, ch + 1);
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&tempty[acc]));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

int g_halo_variant = 0;

}  // namespace

int conv2d_halo(int dtype, int KH, int KW, int Cout, const hoigHaloConvSeg *segs, int nsegs, cudaStream_t stream)
{
    HOIG_REQUIRE(dtype == HOIG_BF16 || dtype == HOIG_F16, "conv2d_halo: tensor-core path needs bf16 or fp16 (got dtype %d)", dtype);
    HOIG_REQUIRE(segs && nsegs >= 1 && nsegs <= 2, "conv2d_halo: 1 or 2 segments");
    HOIG_REQUIRE(KH >= 1 && KW >= 1 && KH <= 7 && KW <= 7 && (KH & 1) && (KW & 1), "conv2d_halo: odd kernel sizes up to 7 (got %dx%d)", KH, KW);
    HOIG_REQUIRE(Cout >= 16 && Cout <= 128 && Cout % 16 == 0, "conv2d_halo: Cout must be a multiple of 16 in [16,128] (got %d)", Cout);
    HaloParams P;
    memset(&P, 0, sizeof(P));
    P.nsegs = nsegs; P.KH = KH; P.KW = KW; P.BN = Cout; P.Cout = Cout; P.variant = g_halo_variant;
    CUtensorMap map_a[2], map_w[2];
    int tiles = 0;
    for (int i = 0; i < nsegs; ++i) {
        const hoigHaloConvSeg &g = segs[i];
        HOIG_REQUIRE(g.src && g.weight && g.dst, "conv2d_halo: null pointer in segment %d", i);
        HOIG_REQUIRE(g.N > 0 && g.Hp >= KH && g.Wp >= KW && g.C > 0 && g.C % 64 == 0, "conv2d_halo: bad shape in segment %d (C must be a multiple of 64)", i);
        HOIG_REQUIRE(g.ld >= g.C && g.ld % 8 == 0 && g.ldd >= Cout && g.ldd % 8 == 0, "conv2d_halo: pixel strides must be multiples of 8 and >= channels");
        HOIG_REQUIRE(((uintptr_t)g.src % 16) == 0 && ((uintptr_t)g.weight % 16) == 0 && ((uintptr_t)g.dst % 16) == 0, "conv2d_halo: 16-byte alignment");
        const int64_t rows = (int64_t)g.N * g.Hp * g.Wp;
        HOIG_REQUIRE(rows + HBM < (1ll << 31), "conv2d_halo: raster too large");
        HaloSeg &S = P.seg[i];
        S.rows = rows; S.pitch = g.Wp; S.C = g.C; S.cblocks = g.C / HBK; S.dst = g.dst; S.ldd = g.ldd; S.tile0 = tiles;
        S.st256 = (getenv("HOIG_UMMA_ST256") == nullptr || (atoi(getenv("HOIG_UMMA_ST256")) & 4) != 0) && g.ldd % 16 == 0 && ((uintptr_t)g.dst % 32) == 0;
        tiles += ceil_div(rows, HBM);
        {
            const cuuint64_t dims[2] = {(cuuint64_t)g.C, (cuuint64_t)rows};
            const cuuint64_t strides[1] = {(cuuint64_t)g.ld * 2};
            const cuuint32_t box[2] = {HBK, (cuuint32_t)(HM + KW - 1)};
            const int st = make_map(&map_a[i], g.src, 2, dims, strides, box, "halo activations", dtype);
            if (st != HOIG_OK) return st;
        }
        {
            const int64_t K = (int64_t)KH * KW * g.C;
            const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
            const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
            const cuuint32_t box[2] = {HBK, (cuuint32_t)Cout};
            const int st = make_map(&map_w[i], g.weight, 2, dims, strides, box, "halo weights", dtype);
            if (st != HOIG_OK) return st;
        }
    }
    if (nsegs == 1) { map_a[1] = map_a[0]; map_w[1] = map_w[0]; }
    P.total_tiles = tiles;
    const int b_stage = Cout * HBK * 2;
    P.a_stages = 3;
    P.b_stages = (H_SMEM_TOTAL - 1024 - P.a_stages * A_STAGE) / b_stage;
    if (P.b_stages > H_MAX_STAGES) P.b_stages = H_MAX_STAGES;
    HOIG_REQUIRE(P.b_stages >= 2, "conv2d_halo: not enough shared memory");

    const int num_sms = device_sm_count();
    if (first_use_on_device(SLOT_CONV_HALO) &&
        (cudaFuncSetAttribute(conv_halo_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM_TOTAL) != cudaSuccess ||
         cudaFuncSetAttribute(conv_halo_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM_TOTAL) != cudaSuccess))
        return check_launch("conv_halo smem attribute");
    const int grid = tiles < num_sms ? tiles : num_sms;
    const size_t smem = 1024 + (size_t)P.a_stages * A_STAGE + (size_t)P.b_stages * b_stage;
    if (dtype == HOIG_F16) conv_halo_kernel<__half><<<grid, H_THREADS, smem, stream>>>(P, map_a[0], map_w[0], map_a[1], map_w[1]);
    else conv_halo_kernel<__nv_bfloat16><<<grid, H_THREADS, smem, stream>>>(P, map_a[0], map_w[0], map_a[1], map_w[1]);
    return check_launch("conv_halo_kernel");
}

}  // namespace hoig

extern "C" void hoig_set_halo_variant(int v) { hoig::g_halo_variant = v; }
