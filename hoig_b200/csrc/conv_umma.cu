// conv_umma.cu -- 16-bit (bf16 / fp16) implicit-GEMM convolution on 5th-gen tensor cores (sm_100a).
//
// D[128 pixels][BN channels] (fp32, TMEM) += A[128][64] (smem) * W[BN][64]^T (smem) per k-block,
// tcgen05.mma.kind::f16, UMMA 128 x BN x 16 (cta_group::1) or 256 x BN x 16 over a CTA pair (cta_group::2), BN <= 256.
//
// Persistent, warp-specialised CTA (one per SM, 640 threads), each walking a contiguous range of tiles:
//   warps 0-15  epilogue, four groups of four warps (warp w: TMEM lane quadrant w % 4, group w / 4; the groups share the
//               16-column chunks round-robin): tcgen05.ld (next chunk in flight while the current one is finalised) -> +bias
//               (smem) (+residual) -> activation, or SPADE modulation of a second tensor -> 16-bit 32-byte row-segment stores;
//               per-plane sum / sum of squares (instance-norm statistics) by warp-level MMA column sums, flushed per image.
//               Overlaps the next tile's main loop (2 TMEM accumulator stages, 4 in dual mode).
//   warps 8-11  double as A producers in "gather" mode (< 64 input channels, fp32-style flow gathers): one output pixel per
//               thread, 16-byte cp.async chunks written straight into the 128B-swizzled K-major layout the UMMA descriptor
//               expects (zero-fill = padding); in local-attention mode the BlockExtractor taps are blended in registers.
//   warp  16    TMA producer, activations: one 4-D box per (tap, 64 channels) of the tap's input view, out-of-bounds rows
//               zero-filled by TMA == conv padding (stride-2 convs and the transposed conv are stride-1 problems over strided
//               views, conv_common.cuh); in row-halo mode one box of 128 + kw - 1 pixels per (kernel row, 64 channels).
//   warp  18    TMA producer, weights (in gather mode warp 16 does this); resident-weights mode loads the matrix once.
//   warps 17,19 TMEM allocation + tcgen05.mma issue (one elected lane each; the second only in dual mode, where two pipelines
//               with their own ring halves and accumulator stages process alternate tiles), tcgen05.commit -> mbarriers.
// Modes chosen per launch on the host (launch_one): tma_a, halo, bres, dual, CTA pairs; DESIGN.md 3.1 has the measurements
// behind each.  smem: ring of `stages` x (A 16/17 KB + weight tiles), 1024-byte aligned for SWIZZLE_128B.
#include <cuda.h>

#include <string.h>

#include <type_traits>

#include "conv_common.cuh"
#include "umma_common.cuh"

namespace hoig {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // bf16 elements per k-block = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int A_HALO_BYTES = 17 * 1024;   // (128 + up to 7 halo pixels) x 128 B, rounded up to the 1024 B swizzle period
constexpr int VH_COLS = 8, VH_ROWS = 16;  // vhalo tile: 8 columns (1024 B of 64 channels = one swizzle period) x 16 rows = 128 pixels
constexpr int EPI_WARPS = 8, PROD_WARPS = 4;    // epilogue warp w: TMEM lane quadrant w % 4, 16-column chunks of parity w / 4
constexpr int EPI_END = 16;                     // warps 12-15: a further epilogue group
constexpr int TMA_WARP = 16, MMA_WARP = 17, WB_WARP = 18, MMA2_WARP = 19;   // WB: weight-tile producer (TMA-A mode), MMA2: second issue stream (dual mode)
constexpr int THREADS = 640;
constexpr int MAX_STAGES = 8;
constexpr int LOOKAHEAD = 3;           // cp.async groups in flight per producer thread
constexpr int STG_BYTES = 0;
constexpr int SMEM_TOTAL = 214 * 1024;   // dynamic; + ~10 KB static (barriers, stats, bias) <= 227 KB per CTA
constexpr int RING_BUDGET = SMEM_TOTAL - 1024 - STG_BYTES;

struct UmmaParams {
    ConvParams c;
    int BN;          // UMMA N (multiple of 16, <= 256)
    int n_tiles;     // ceil(Npad / BN)
    int m_tiles;     // N * tiles_per_image
    int k_blocks;    // Kpad / 64
    int stages;
    int tma_a;       // 1: activations by TMA boxes, 0: cp.async / register gather
    int cpt_shift;   // log2(Cin / 8) when Cin < 64 (chunk -> tap by shift), -1: generic division
    int tpi_shift;       // log2(tiles_per_image) when it is a power of two, else -1
    int contig;          // 1: contiguous tile range per CTA, 0: tiles strided by the grid size
    int st256;           // FAST kernels: output rows are 32-byte aligned, 16-column pieces go out as one 256-bit store per lane
    int fast_epi;        // 2: streamlined epilogue over 32-column chunks (plain bias / ReLU / statistics convs, see launch_one)
    int halo;            // 1: regular kh x kw stride-1 conv whose tiles are 128 consecutive pixels of ONE image row: a stage holds one
                         //    activation box of 128 + kw - 1 pixels per (kernel row, 64 channels) and the kw taps read it through
                         //    descriptors shifted by one pixel row (128 B) each -- L2->SM activation traffic / kw (see conv_halo.cu)
    int k_steps;         // halo: kh * Cin / 64 pipeline steps of kw taps each; vhalo: Cin / 64 steps of kh taps each
    int vhalo;           // 1: kh x 1 conv (the re-associated 7x7 stems / heads) on 8-column x 16-row tiles with RESIDENT weights: a stage
                         //    holds ONE activation box of 8 x (16 + kh - 1) pixels per 64 channels; tap r reads it r pixel-rows of the
                         //    box (8 pixels = 1024 B, a whole swizzle period) further on -- L2->SM activation traffic / (kh * 16 / 22)
    int tiles_x;         // vhalo: tiles per image row (GW / 8)
    int mma_stats;       // 1: per-plane statistics reduced with warp-level MMAs (colsum16), 0: shuffle transpose-reduce
    int bres;            // 1: the whole packed weight matrix (k_blocks tiles of BN x 64) stays RESIDENT in shared memory, loaded
                         //    once per CTA; the ring then carries activations only.  For the narrow high-resolution convs the
                         //    L2->SM path (~42 B/clk/SM), not the tensor pipe, is the bound, and weights are 1/3 of that traffic.
    int dual;            // 1: two independent pipelines per CTA (narrow N): tiles alternate between two MMA-issuing warps,
                         //    each with its own half of the smem ring and two of the four TMEM accumulator stages
    uint8_t tap_live[16]; // transposed conv: per n-tile, bit t set = tap t of the 2x2 input neighbourhood feeds some parity block of the tile;
                         //    dead (n-tile, tap) k-blocks are skipped by the producers and the MMA issuer (their weight blocks are zero)
    int tap_skip;        // 1: tap_live is in use (TMA-fed transposed conv)
    int backoff_ns;      // epilogue: maximum nanosleep between polls of the accumulator-full barrier (HOIG_UMMA_BACKOFF_NS, 0 = spin)
    int prefetch;        // epilogue: SPADE activation / residual loads issued one chunk ahead (HOIG_UMMA_PREFETCH)
    int early_release;   // epilogue: hand the accumulator stage back right after the last tcgen05.ld (HOIG_UMMA_EARLY_RELEASE)
    int relaxed_release; // epilogue: relaxed (signal-only) arrival on the accumulator barrier (HOIG_UMMA_RELAXED_RELEASE)
    int debug;           // timing knock-outs (HOIG_UMMA_DEBUG, results are garbage): 1 = epilogue only drains TMEM, 2 = A tile loaded once per tile, 4 = no statistics, 8 = no output stores
};

// ------------------------------------------------------------------------ kernel
// T = __nv_bfloat16 or __half (16-bit storage; tcgen05 kind::f16 handles both).
// NCTA = 2: CTA pair (cluster of 2, cta_group::2).  The pair computes a 256-pixel x BN tile: each CTA stages its own 128
// pixels of A and HALF of the weight tile (BN/2 rows), the leader (cluster rank 0) issues 256 x BN x 16 MMAs that read both
// CTAs' shared memory, and each CTA drains its own 128 TMEM lanes.  Weight traffic (L2->SM and smem writes) per CTA halves,
// which is what bounds the wide-N convs: TMA writes and UMMA operand reads share the 128 B/clk shared-memory port.
template <typename T, int NCTA, bool PERSIST, bool FAST>
__global__ void __launch_bounds__(THREADS, 1)
conv_umma_kernel(const UmmaParams P, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3)
{
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[4], tempty_bar[4];
    __shared__ __align__(8) uint64_t bres_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_stats[4][2][256];   // per TMEM lane quadrant: plain += by its owning warps, no shared atomics
    __shared__ __align__(16) float s_bias[256 + 32];

    const ConvParams &p = P.c;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int BN = P.BN;
    const int TPS = P.halo ? p.kw : (P.vhalo ? p.kh : 1);               // taps per pipeline step
    const int a_bytes = P.halo ? A_HALO_BYTES : (P.vhalo ? (VH_ROWS + p.kh - 1) * VH_COLS * 128 : A_STAGE_BYTES);   // activation region of a stage
    const int b_tile = (BN / NCTA) * BK * 2;                            // one tap's weight tile (this CTA's share)
    const int stage_bytes = P.bres ? a_bytes : a_bytes + TPS * b_tile;
    const int n_steps = (P.halo || P.vhalo) ? P.k_steps : P.k_blocks;
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t stg_base = smem_u32(smem);                  // epilogue staging first (1024-aligned)
    const uint32_t smem_base = stg_base + STG_BYTES;           // then the operand ring
    const int npix = p.GH * p.GW;
    const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;  // 0 = leader of the pair
    // Each CTA (pair) walks a CONTIGUOUS range of tiles: consecutive tiles then share their image (and n-tile), so the
    // epilogue flushes the per-plane statistics once per image instead of once per tile, and the vertical taps of
    // consecutive tiles re-read rows this CTA has just pulled through L2.
    const int tile_units = (int)gridDim.x / NCTA, unit = (int)blockIdx.x / NCTA;
    const int all_tiles = (P.m_tiles / NCTA) * P.n_tiles;
    const int tile0 = P.contig ? (int)((int64_t)unit * all_tiles / tile_units) : unit;
    const int total_tiles = P.contig ? (int)((int64_t)(unit + 1) * all_tiles / tile_units) : all_tiles;   // end of this CTA's range
    const int tile_step = P.contig ? 1 : tile_units;

    if (threadIdx.x == 0) {
        // TMA-A: one arrive.expect_tx per producer thread (only the activation thread when the weights are resident)
        const uint32_t full_count = P.tma_a ? (P.bres ? 1u : 2u) : (uint32_t)(PROD_WARPS * 32 + 1);
        mbar_init(smem_u32(&bres_bar), 1);
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), full_count);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int a = 0; a < 4; ++a) {
            mbar_init(smem_u32(&tfull_bar[a]), 1);
            mbar_init(smem_u32(&tempty_bar[a]), (uint32_t)(NCTA * (P.tma_a ? EPI_END : EPI_END - PROD_WARPS) * 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 4 * 2 * 256; i += blockDim.x) (&s_stats[0][0][0])[i] = 0.f;
    if (warp == MMA_WARP) {
        if (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_smem)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_smem)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (warp == TMA_WARP && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        if (P.tma_a) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
    }
    tc_fence_before();
    if (NCTA == 2) cluster_sync();   // the peer's barriers must be initialised before anything is signalled on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    // In TMA-A mode the four producer warps have nothing to gather and join the epilogue as a third group.
    if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS && !P.tma_a) {
        // =============================================== A producers (gather mode)
        {
            const int row = threadIdx.x - EPI_WARPS * 32;  // 0..127 : A row == pixel of the tile
            const uint32_t row_off = (uint32_t)row * 128u;
            const uint32_t sw = (uint32_t)(row & 7);
            const T *src0 = static_cast<const T *>(p.view[0].base);
            uint32_t it = 0;  // running k-block counter across tiles
            int pending = 0;  // k-blocks issued but not yet signalled
            uint32_t sig_it = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int mt = tile / P.n_tiles;
                const int n_img = mt / p.tiles_per_image;
                const int pix = (mt % p.tiles_per_image) * BM + row;
                const bool valid = pix < npix;
                const int gy = valid ? pix / p.GW : 0, gx = valid ? pix % p.GW : 0;
                int tap64 = 0, c64 = 0;   // (tap, channel) of the k-block start when Cin % 64 == 0
                for (int kb = 0; kb < P.k_blocks; ++kb, ++it) {
                    const int s = it % P.stages;
                    mbar_wait(smem_u32(&empty_bar[s]), ((it / P.stages) & 1) ^ 1);
                    const uint32_t a_dst = smem_base + (uint32_t)s * stage_bytes + row_off;
                    if (p.mode != HOIG_CONV_LOCAL_ATTN) {
                        if (P.cpt_shift < 0 && (p.Cin & 63) == 0) {
                            // all 8 chunks of this k-block share one tap: one bounds test, one base pointer
                            const T *src = src0;
                            uint32_t bytes = 0;
                            if (valid && tap64 < p.ntaps) {
                                const bool second = c64 >= p.C0;
                                const InputView &vw = second ? p.view1 : p.view[p.tap_map[tap64]];
                                const int iy = gy * p.stride + p.tap_dy[tap64], ix = gx * p.stride + p.tap_dx[tap64];
                                if (iy >= 0 && iy < vw.H && ix >= 0 && ix < vw.W) {
                                    src = static_cast<const T *>(vw.base) + view_off(vw, n_img, iy, ix) + (second ? c64 - p.C0 : c64);
                                    bytes = 16;
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                cp_async_16(a_dst + (((uint32_t)j ^ sw) << 4), src + (bytes ? j * 8 : 0), bytes);
                            c64 += BK;
                            if (c64 >= p.Cin) { c64 = 0; ++tap64; }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int chunk = kb * 8 + j;
                                const T *src = src0;
                                uint32_t bytes = 0;
                                int tap, c;
                                if (P.cpt_shift >= 0) { tap = chunk >> P.cpt_shift; c = (chunk - (tap << P.cpt_shift)) * 8; }
                                else { tap = (chunk * 8) / p.Cin; c = chunk * 8 - tap * p.Cin; }
                                if (valid && tap < p.ntaps) {
                                    const bool second = c >= p.C0;
                                    const InputView &vw = second ? p.view1 : p.view[p.tap_map[tap]];
                                    const int iy = gy * p.stride + p.tap_dy[tap], ix = gx * p.stride + p.tap_dx[tap];
                                    if (iy >= 0 && iy < vw.H && ix >= 0 && ix < vw.W) {
                                        src = static_cast<const T *>(vw.base) + view_off(vw, n_img, iy, ix) + (second ? c - p.C0 : c);
                                        bytes = 16;
                                    }
                                }
                                cp_async_16(a_dst + (((uint32_t)j ^ sw) << 4), src, bytes);
                            }
                        }
                    } else {
                        // local attention: chunks of [target | source] channels per k x k tap
                        // (extract_attn.py:24-26); target taps are plain copies (zero flow),
                        // source taps are the BlockExtractor bilinear blend, done in registers.
                        const InputView &vt = p.view[0];
                        const InputView &vs = p.view1;
                        int cached_tap = -1;
                        BETap t;
#pragma unroll 2
                        for (int j = 0; j < 8; ++j) {
                            const int k0 = kb * BK + j * 8;
                            const uint32_t dst_j = a_dst + (((uint32_t)j ^ sw) << 4);
                            if (!valid || k0 >= p.K) { cp_async_16(dst_j, src0, 0); continue; }
                            const int tap = k0 / p.Cin;
                            int c = k0 - tap * p.Cin;
                            const int r = tap / p.KH, sx = tap - r * p.KH;
                            if (c < p.C0) {
                                const int iy = max(min(gy + r - p.KH / 2, vt.H - 1), 0), ix = max(min(gx + sx - p.KH / 2, vt.W - 1), 0);
                                cp_async_16(dst_j, static_cast<const T *>(vt.base) + view_off(vt, n_img, iy, ix) + c, 16);
                            } else {
                                c -= p.C0;
                                if (tap != cached_tap) {
                                    const float *fl = p.flow + (((int64_t)n_img * p.GH + gy) * p.GW + gx) * 2;
                                    t = be_tap(fl[0], fl[1], gy, gx, r, sx, p.KH, vs.H, vs.W);
                                    cached_tap = tap;
                                }
                                const T *sb = static_cast<const T *>(vs.base) + (int64_t)n_img * vs.sn + c;
                                float acc[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    float u[8];
                                    load8(sb + (int64_t)t.idx[q] * vs.sx, u);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) acc[e] = __fmaf_rn(t.w[q], u[e], acc[e]);
                                }
                                st_shared_v4(dst_j, pack2<T>(acc[0], acc[1]), pack2<T>(acc[2], acc[3]),
                                             pack2<T>(acc[4], acc[5]), pack2<T>(acc[6], acc[7]));
                            }
                        }
                    }
                    cp_async_commit();
                    ++pending;
                    if (pending > LOOKAHEAD) {
                        cp_async_wait<LOOKAHEAD>();
                        fence_proxy_async();
                        mbar_arrive(smem_u32(&full_bar[sig_it % P.stages]));
                        ++sig_it; --pending;
                    }
                }
            }
            cp_async_wait<0>();
            fence_proxy_async();
            while (pending > 0) {
                mbar_arrive(smem_u32(&full_bar[sig_it % P.stages]));
                ++sig_it; --pending;
            }
        }
    } else if (warp == TMA_WARP || warp == WB_WARP) {
        // ======================================================== TMA producers
        // TMA-A mode: warp TMA_WARP streams the activation boxes, warp WB_WARP the weight tiles (two threads, because one
        // thread's wait + expect_tx + two TMA issues per k-block take longer than the four MMAs of a narrow-N k-block).
        // Gather mode: TMA_WARP streams the weight tiles, WB_WARP has nothing to do.
        const bool do_a = P.tma_a && warp == TMA_WARP;
        const bool do_b = P.tma_a ? warp == WB_WARP : warp == TMA_WARP;
        const uint32_t bres_base = smem_base + (uint32_t)P.stages * (uint32_t)stage_bytes;   // resident weights follow the ring
        if (do_b && P.bres) {
            // resident weights: all k-block tiles once, one barrier
            if (elect_one()) {
                mbar_arrive_expect_tx(smem_u32(&bres_bar), (uint32_t)P.k_blocks * (uint32_t)(BN * BK * 2));
                for (int kb = 0; kb < P.k_blocks; ++kb)
                    tma_load_2d(bres_base + (uint32_t)kb * (uint32_t)(BN * BK * 2), &map_w, smem_u32(&bres_bar), kb * BK, 0);
            }
            __syncwarp();
        } else if (do_a || do_b) {   // whole warp runs the loop; one elected lane issues
            const CUtensorMap *maps[4] = {&map_a0, &map_a1, &map_a2, &map_a3};
            // NP pipelines (1, or 2 in dual mode): pipeline w owns ring stages w, w+NP, ... and the CTA's tiles w, w+NP, ...
            const int NP = P.dual ? 2 : 1;
            const uint32_t sp = (uint32_t)(P.stages / NP);             // stages per pipeline
            uint32_t q[2] = {0, 0}, ph[2] = {0, 0};
            const uint32_t empty0 = smem_u32(&empty_bar[0]);
            // the "full" barriers live in the leader CTA: both CTAs' TMA loads complete on them
            const uint32_t full0 = NCTA == 2 ? mapa(smem_u32(&full_bar[0]), 0) : smem_u32(&full_bar[0]);
            const uint32_t b_rows = (uint32_t)(BN / NCTA);
            const uint32_t b_bytes = (uint32_t)(BN * BK * 2);          // whole pair
            for (int tbase = tile0; tbase < total_tiles; tbase += NP * tile_step) {
                int n_img[2], gy0[2], gx0[2], nt[2];
                bool live[2] = {false, false};
                for (int w = 0; w < NP; ++w) {
                    const int tile = tbase + w * tile_step;
                    live[w] = tile < total_tiles;
                    int mt = tile;
                    nt[w] = 0;
                    if (P.n_tiles > 1) { mt = tile / P.n_tiles; nt[w] = tile - mt * P.n_tiles; }
                    mt = mt * NCTA + (int)rank;
                    n_img[w] = mt / p.tiles_per_image;
                    if (P.vhalo) {
                        const int t_in = mt % p.tiles_per_image;
                        gy0[w] = (t_in / P.tiles_x) * VH_ROWS; gx0[w] = (t_in % P.tiles_x) * VH_COLS;
                    } else {
                        const int pix0 = (mt % p.tiles_per_image) * BM;
                        gy0[w] = pix0 / p.GW; gx0[w] = pix0 % p.GW;
                    }
                }
                int tap = 0, c = 0;
                if (P.vhalo) {
                    // vertical-halo mode (weights resident, so only the activation warp is here): step = 64-channel slice, ONE box of
                    // 8 columns x (16 + kh - 1) rows starting pad_h rows above the tile
                    for (int kb = 0; kb < P.k_steps; ++kb) {
                        for (int w = 0; w < NP; ++w) {
                            if (!live[w]) continue;
                            const uint32_t st = q[w] * (uint32_t)NP + (uint32_t)w;
                            mbar_wait(empty0 + 8u * st, ph[w] ^ 1u);
                            if (elect_one()) {
                                mbar_arrive_expect_tx(full0 + 8u * st, (uint32_t)a_bytes);
                                tma_load_4d(smem_base + st * (uint32_t)stage_bytes, &map_a0, full0 + 8u * st, kb * BK, gx0[w], gy0[w] - p.pad_h, n_img[w]);
                            }
                            __syncwarp();
                            if (++q[w] == sp) { q[w] = 0; ph[w] ^= 1u; }
                        }
                    }
                } else if (P.tap_skip) {
                    // transposed conv: same loop, minus the (n-tile, tap) k-blocks whose weight blocks are structurally zero
                    const int lmask[2] = {P.tap_live[nt[0]], P.tap_live[NP > 1 ? nt[1] : nt[0]]};
                    for (int kb = 0; kb < P.k_blocks; ++kb) {
                        const int tt = tap < p.ntaps ? tap : p.ntaps - 1;   // K padding blocks: any finite data (weights are zero)
                        for (int w = 0; w < NP; ++w) {
                            if (!live[w] || !((lmask[w] >> tt) & 1)) continue;
                            // ring position / phase kept incrementally: a runtime division per k-block on this single
                            // thread costs more than the MMAs of a short k-block
                            const uint32_t st = q[w] * (uint32_t)NP + (uint32_t)w;
                            mbar_wait(empty0 + 8u * st, ph[w] ^ 1u);
                            const uint32_t bar = full0 + 8u * st;
                            const uint32_t a_dst = smem_base + st * (uint32_t)stage_bytes;
                            if (do_a) {
                                const bool skip_a = NCTA == 1 && (P.debug & 2) && kb > 0;
                                if (elect_one()) {
                                    if (skip_a) {
                                        mbar_arrive(bar);
                                    } else if (NCTA == 2) {
                                        if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[0]) + 8u * st, 2u * (uint32_t)A_STAGE_BYTES);
                                        tma_load_4d_2sm(a_dst, maps[p.tap_map[tt]], bar, c, gx0[w] + p.tap_dx[tt], gy0[w] + p.tap_dy[tt], n_img[w]);
                                    } else {
                                        mbar_arrive_expect_tx(bar, (uint32_t)A_STAGE_BYTES);
                                        tma_load_4d(a_dst, maps[p.tap_map[tt]], bar, c, gx0[w] + p.tap_dx[tt], gy0[w] + p.tap_dy[tt], n_img[w]);
                                    }
                                }
                            } else if (elect_one()) {
                                if (NCTA == 2) {
                                    if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[0]) + 8u * st, b_bytes);
                                    tma_load_2d_2sm(a_dst + A_STAGE_BYTES, &map_w, bar, kb * BK, nt[w] * BN + (int)(rank * b_rows));
                                } else {
                                    mbar_arrive_expect_tx(bar, b_bytes);
                                    tma_load_2d(a_dst + A_STAGE_BYTES, &map_w, bar, kb * BK, nt[w] * BN);
                                }
                            }
                            __syncwarp();
                            if (++q[w] == sp) { q[w] = 0; ph[w] ^= 1u; }
                        }
                        c += BK;
                        if (c >= p.Cin) { c = 0; ++tap; }
                    }
                } else if (!P.halo) {
                    for (int kb = 0; kb < P.k_blocks; ++kb) {
                        const int tt = tap < p.ntaps ? tap : p.ntaps - 1;   // K padding blocks: any finite data (weights are zero)
                        for (int w = 0; w < NP; ++w) {
                            if (!live[w]) continue;
                            // ring position / phase kept incrementally: a runtime division per k-block on this single
                            // thread costs more than the MMAs of a short k-block
                            const uint32_t st = q[w] * (uint32_t)NP + (uint32_t)w;
                            mbar_wait(empty0 + 8u * st, ph[w] ^ 1u);
                            const uint32_t bar = full0 + 8u * st;
                            const uint32_t a_dst = smem_base + st * (uint32_t)stage_bytes;
                            if (do_a) {
                                const bool skip_a = NCTA == 1 && (P.debug & 2) && kb > 0;
                                if (elect_one()) {
                                    if (skip_a) {
                                        mbar_arrive(bar);
                                    } else if (NCTA == 2) {
                                        if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[0]) + 8u * st, 2u * (uint32_t)A_STAGE_BYTES);
                                        tma_load_4d_2sm(a_dst, maps[p.tap_map[tt]], bar, c, gx0[w] + p.tap_dx[tt], gy0[w] + p.tap_dy[tt], n_img[w]);
                                    } else {
                                        mbar_arrive_expect_tx(bar, (uint32_t)A_STAGE_BYTES);
                                        tma_load_4d(a_dst, maps[p.tap_map[tt]], bar, c, gx0[w] + p.tap_dx[tt], gy0[w] + p.tap_dy[tt], n_img[w]);
                                    }
                                }
                            } else if (elect_one()) {
                                if (NCTA == 2) {
                                    if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[0]) + 8u * st, b_bytes);
                                    tma_load_2d_2sm(a_dst + A_STAGE_BYTES, &map_w, bar, kb * BK, nt[w] * BN + (int)(rank * b_rows));
                                } else {
                                    mbar_arrive_expect_tx(bar, b_bytes);
                                    tma_load_2d(a_dst + A_STAGE_BYTES, &map_w, bar, kb * BK, nt[w] * BN);
                                }
                            }
                            __syncwarp();
                            if (++q[w] == sp) { q[w] = 0; ph[w] ^= 1u; }
                        }
                        c += BK;
                        if (c >= p.Cin) { c = 0; ++tap; }
                    }
                } else {
                    // row-halo mode: step = (kernel row `tap`, channel block c): ONE activation box of 128 + kw - 1 pixels starting
                    // pad_w pixels left of the tile, and the kw weight tiles of that kernel row
                    const uint32_t a_tx = (uint32_t)(BM + TPS - 1) * 128u;
                    for (int kb = 0; kb < P.k_steps; ++kb) {
                        for (int w = 0; w < NP; ++w) {
                            if (!live[w]) continue;
                            const uint32_t st = q[w] * (uint32_t)NP + (uint32_t)w;
                            mbar_wait(empty0 + 8u * st, ph[w] ^ 1u);
                            const uint32_t bar = full0 + 8u * st;
                            const uint32_t a_dst = smem_base + st * (uint32_t)stage_bytes;
                            if (do_a) {
                                if (elect_one()) {
                                    if (NCTA == 2) {
                                        if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[0]) + 8u * st, 2u * a_tx);
                                        tma_load_4d_2sm(a_dst, &map_a0, bar, c, gx0[w] - p.pad_w, gy0[w] + tap - p.pad_h, n_img[w]);
                                    } else {
                                        mbar_arrive_expect_tx(bar, a_tx);
                                        tma_load_4d(a_dst, &map_a0, bar, c, gx0[w] - p.pad_w, gy0[w] + tap - p.pad_h, n_img[w]);
                                    }
                                }
                            } else {
                                if (elect_one()) {
                                    if (NCTA == 2) {
                                        if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[0]) + 8u * st, (uint32_t)TPS * b_bytes);
                                    } else {
                                        mbar_arrive_expect_tx(bar, (uint32_t)TPS * b_bytes);
                                    }
                                }
                                for (int sft = 0; sft < TPS; ++sft) {      // warp-uniform loop, one elected lane issues each load
                                    const int kcol = (tap * TPS + sft) * p.Cin + c;
                                    const uint32_t b_dst = a_dst + (uint32_t)A_HALO_BYTES + (uint32_t)(sft * b_tile);
                                    if (elect_one()) {
                                        if (NCTA == 2) tma_load_2d_2sm(b_dst, &map_w, bar, kcol, nt[w] * BN + (int)(rank * b_rows));
                                        else tma_load_2d(b_dst, &map_w, bar, kcol, nt[w] * BN);
                                    }
                                }
                            }
                            __syncwarp();
                            if (++q[w] == sp) { q[w] = 0; ph[w] ^= 1u; }
                        }
                        c += BK;
                        if (c >= p.Cin) { c = 0; ++tap; }
                    }
                }
            }
        }
    } else if (warp == MMA_WARP || (warp == MMA2_WARP && P.dual)) {
        // ========================================================== MMA issuer
        if (rank == 0) {   // leader only.  Whole warp runs the loop (waits, fences); one elected lane issues the MMAs and commits
            // instruction descriptor: D = f32 (bit 4), A/B format bf16 = 1 / f16 = 0 (bits 7-9 / 10-12), K-major A and B,
            // N >> 3 at bit 17, M >> 4 at bit 24 (M = 256 for the CTA pair)
            constexpr uint32_t kFmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
            const uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((NCTA * BM) >> 4) << 24);
            const int NP = P.dual ? 2 : 1, w = warp == MMA_WARP ? 0 : 1;    // this warp's pipeline
            const uint32_t sp = (uint32_t)(P.stages / NP);
            uint32_t q = 0, ph = 0;
            const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
            const uint64_t desc0 = umma_desc(smem_base);             // + (byte offset >> 4) addresses any tile of the ring
            const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
            // i = index of the tile in this CTA's sequence; TMEM accumulator stage i % (2*NP), used for the (i / (2*NP))-th time
            const uint64_t bres_desc = desc0 + (uint64_t)(((uint32_t)P.stages * (uint32_t)stage_bytes) >> 4);
            if (P.bres) { mbar_wait(smem_u32(&bres_bar), 0); tc_fence_after(); }
            for (uint32_t i = (uint32_t)w; tile0 + (int)i * tile_step < total_tiles; i += (uint32_t)NP) {
                const uint32_t acc = i % (uint32_t)(2 * NP), use = i / (uint32_t)(2 * NP);
                mbar_wait(smem_u32(&tempty_bar[acc]), (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)BN;
                if (P.vhalo) {
                    for (int kb = 0; kb < P.k_steps; ++kb) {
                        const uint32_t st = q * (uint32_t)NP + (uint32_t)w;
                        mbar_wait(full0 + 8u * st, ph);
                        tc_fence_after();
                        const uint64_t da = desc0 + (uint64_t)(st * stage16);
                        // tap r reads the shared box r pixel-rows (8 pixels = 1024 B) further on; its weight k-block is r * (Cin / 64) + kb
                        for (int r = 0; r < TPS; ++r) {             // warp-uniform loop, one elected lane issues
                            const uint64_t das = da + (uint64_t)(r * (VH_COLS * 128 / 16));
                            const uint64_t dbs = bres_desc + (uint64_t)((uint32_t)(r * P.k_steps + kb) * (uint32_t)(BN * 8));
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    umma_bf16(d_tmem, das + (uint64_t)(2 * k), dbs + (uint64_t)(2 * k), idesc, (kb | r | k) ? 1u : 0u);
                                if (r == TPS - 1) umma_commit(empty0 + 8u * st);
                            }
                            __syncwarp();
                        }
                        if (++q == sp) { q = 0; ph ^= 1u; }
                    }
                } else if (P.tap_skip) {
                    // transposed conv: skip the (n-tile, tap) k-blocks the producers skipped (tap 0 is always live, so kb = 0 initialises)
                    const int tile = tile0 + (int)i * tile_step;
                    const int live_mask = P.tap_live[P.n_tiles > 1 ? tile % P.n_tiles : 0];
                    const int cin_blocks = p.Cin / BK;
                    int tap = 0, cb = 0;
                    for (int kb = 0; kb < P.k_blocks; ++kb) {
                        const int t = tap;
                        if (++cb == cin_blocks) { cb = 0; ++tap; }
                        if (!((live_mask >> t) & 1)) continue;
                        const uint32_t st = q * (uint32_t)NP + (uint32_t)w;
                        mbar_wait(full0 + 8u * st, ph);
                        tc_fence_after();
                        const uint64_t da = desc0 + (uint64_t)(st * stage16);
                        const uint64_t db = P.bres ? bres_desc + (uint64_t)((uint32_t)kb * (uint32_t)(BN * 8)) : da + (uint64_t)(A_STAGE_BYTES >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                if (NCTA == 2) umma_bf16_2sm(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                                else umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                            }
                            // frees the smem stage (in both CTAs of a pair) when these MMAs retire
                            if (NCTA == 2) umma_commit_2sm(empty0 + 8u * st);
                            else umma_commit(empty0 + 8u * st);
                        }
                        __syncwarp();
                        if (++q == sp) { q = 0; ph ^= 1u; }
                    }
                } else if (!P.halo) {
                    for (int kb = 0; kb < P.k_blocks; ++kb) {
                        const uint32_t st = q * (uint32_t)NP + (uint32_t)w;
                        mbar_wait(full0 + 8u * st, ph);
                        tc_fence_after();
                        const uint64_t da = desc0 + (uint64_t)(st * stage16);
                        const uint64_t db = P.bres ? bres_desc + (uint64_t)((uint32_t)kb * (uint32_t)(BN * 8)) : da + (uint64_t)(A_STAGE_BYTES >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                if (NCTA == 2) umma_bf16_2sm(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                                else umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                            }
                            // frees the smem stage (in both CTAs of a pair) when these MMAs retire
                            if (NCTA == 2) umma_commit_2sm(empty0 + 8u * st);
                            else umma_commit(empty0 + 8u * st);
                        }
                        __syncwarp();
                        if (++q == sp) { q = 0; ph ^= 1u; }
                    }
                } else {
                    for (int kb = 0; kb < P.k_steps; ++kb) {
                        const uint32_t st = q * (uint32_t)NP + (uint32_t)w;
                        mbar_wait(full0 + 8u * st, ph);
                        tc_fence_after();
                        const uint64_t da = desc0 + (uint64_t)(st * stage16);
                        const uint64_t db = da + (uint64_t)(A_HALO_BYTES >> 4);
                        // tap s reads the shared activation box s pixel rows (128 B) further on and its own weight tile
                        for (int sft = 0; sft < TPS; ++sft) {       // warp-uniform loop, one elected lane issues
                            const uint64_t das = da + (uint64_t)(sft * 8), dbs = db + (uint64_t)(sft * (b_tile >> 4));
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k) {
                                    const uint32_t accum = (kb | sft | k) ? 1u : 0u;
                                    if (NCTA == 2) umma_bf16_2sm(d_tmem, das + (uint64_t)(2 * k), dbs + (uint64_t)(2 * k), idesc, accum);
                                    else umma_bf16(d_tmem, das + (uint64_t)(2 * k), dbs + (uint64_t)(2 * k), idesc, accum);
                                }
                                if (sft == TPS - 1) {
                                    if (NCTA == 2) umma_commit_2sm(empty0 + 8u * st);
                                    else umma_commit(empty0 + 8u * st);
                                }
                            }
                            __syncwarp();
                        }
                        if (++q == sp) { q = 0; ph ^= 1u; }
                    }
                }
                if (elect_one()) {   // accumulator complete (signalled in both CTAs of a pair)
                    if (NCTA == 2) umma_commit_2sm(smem_u32(&tfull_bar[acc]));
                    else umma_commit(smem_u32(&tfull_bar[acc]));
                }
                __syncwarp();
            }
        }
    } else if (warp < EPI_END) {
        // ============================================================ epilogue
        // lane == tile row within the warp's TMEM lane quadrant; the two warps of a quadrant take
        // alternate 16-column chunks.  Chunk i+1 is in flight (tcgen05.ld) while chunk i is finalised.
        // warp groups {0-3}, {4-7}, ({8-11} in TMA-A mode), {12-15}: each group covers the four TMEM lane quadrants and the
        // groups share the 16-column chunks round-robin
        const int quad = warp & 3;
        const int ngrp = P.tma_a ? 4 : 3;
        const int half = (!P.tma_a && warp >= EPI_WARPS + PROD_WARPS) ? 2 : warp >> 2;   // chunk group of this warp
        const int epi_threads = ngrp * 128;
        const int epi_tid = half * 128 + quad * 32 + lane;
        const T *res = static_cast<const T *>(p.residual);
        T *dst = static_cast<T *>(p.dst);
        const bool vec_ok = ((p.ldd & 7) == 0) && (!res || (p.ldr & 7) == 0);
        const int n_chunks = BN / 16;
        const bool all_valid = npix % BM == 0;
        uint32_t tcount = 0;
        const uint32_t tempty0 = NCTA == 2 ? mapa(smem_u32(&tempty_bar[0]), 0) : smem_u32(&tempty_bar[0]);   // in the leader CTA
        int cur_img = -1, cur_nt = -1;
        uint32_t e_sel[4], e_sel_bf[4];      // 0/1 selection fragments of colsum16 in the operand type / in bf16 (squares)
        colsum_select<std::is_same<T, __half>::value>(lane, e_sel);
        colsum_select<false>(lane, e_sel_bf);
        // persistent statistics accumulators (see finalize): valid when this warp sees at most two chunks per tile and one n-tile
        // (PERSIST kernels are launched only for statistics-producing convs with one n-tile and n_chunks <= 2 * ngrp, see launch_one)
        constexpr bool persist = PERSIST;
        constexpr int NSLOT = PERSIST ? 2 : 1;
        float st_sum[NSLOT][4], st_sq[NSLOT][4];
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl)
#pragma unroll
            for (int j = 0; j < 4; ++j) { st_sum[sl][j] = 0.f; st_sq[sl][j] = 0.f; }
        auto dump_stats = [&]() {      // registers -> this warp's private s_stats columns, then clear
#pragma unroll
            for (int sl = 0; sl < NSLOT; ++sl) {
                const int ch = (FAST && P.fast_epi == 2) ? 2 * half + sl : half + sl * ngrp;      // the 16-column chunk slot sl accumulated
                float s_lo = st_sum[sl][0] + st_sum[sl][1], s_hi = st_sum[sl][2] + st_sum[sl][3];
                float q_lo = st_sq[sl][0] + st_sq[sl][1], q_hi = st_sq[sl][2] + st_sq[sl][3];
                s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1);
                q_lo += __shfl_xor_sync(0xffffffffu, q_lo, 1); q_hi += __shfl_xor_sync(0xffffffffu, q_hi, 1);
                s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
                q_lo += __shfl_xor_sync(0xffffffffu, q_lo, 2); q_hi += __shfl_xor_sync(0xffffffffu, q_hi, 2);
                if (ch < n_chunks && (lane & 3) == 0) {
                    const int col = ch * 16 + (lane >> 2);
                    s_stats[quad][0][col] += s_lo; s_stats[quad][0][col + 8] += s_hi;
                    s_stats[quad][1][col] += q_lo; s_stats[quad][1][col + 8] += q_hi;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) { st_sum[sl][j] = 0.f; st_sq[sl][j] = 0.f; }
            }
        };
        const bool spade = p.spade_x != nullptr;
        float *s_mod = &s_stats[0][0][0];   // SPADE epilogue: mean[1024] | rstd[1024] (the statistics buffer is idle in that mode)
        // per-plane sum / sum of squares of the tiles since the last flush: one atomicAdd(double) per column
        auto flush_stats = [&](int img, int ntile) {
            const int pcs = p.phase_cout;
            for (int i = epi_tid; i < 2 * BN; i += epi_threads) {
                const int which = i / BN, col = i % BN;
                const int n = ntile * BN + col;
                const float tot = s_stats[0][which][col] + s_stats[1][which][col] + s_stats[2][which][col] + s_stats[3][which][col];
                if (n < p.Cout)
                    atomicAdd(&p.stats[((int64_t)img * (pcs ? pcs : p.Cout) + (pcs ? n % pcs : n)) * 2 + which], (double)tot);
                s_stats[0][which][col] = 0.f; s_stats[1][which][col] = 0.f; s_stats[2][which][col] = 0.f; s_stats[3][which][col] = 0.f;
            }
        };
        for (int tile = tile0; tile < total_tiles; tile += tile_step, ++tcount) {
            int mt = tile, nt = 0;
            if (P.n_tiles > 1) { mt = tile / P.n_tiles; nt = tile - mt * P.n_tiles; }
            mt = mt * NCTA + (int)rank;
            const int n_img = P.tpi_shift >= 0 ? mt >> P.tpi_shift : mt / p.tiles_per_image;
            int pix = (mt - n_img * p.tiles_per_image) * BM + quad * 32 + lane;
            if (P.vhalo) {   // 8-column x 16-row tiles, row-major over the image
                const int t_in = mt - n_img * p.tiles_per_image, ty = t_in / P.tiles_x, r = quad * 32 + lane;
                pix = (ty * VH_ROWS + (r >> 3)) * p.GW + (t_in - ty * P.tiles_x) * VH_COLS + (r & 7);
            }
            const bool valid = pix < npix;
            const int64_t m = valid ? out_pixel(p, n_img, pix) : 0;
            const int pc = p.phase_cout;     // > 0: transposed conv, column n = phase * pc + channel
            const uint32_t nacc = P.dual ? 4u : 2u;
            const uint32_t acc = tcount % nacc, use = tcount / nacc;
            if (n_img != cur_img || nt != cur_nt) {
                // new (image, n-tile): hand the finished plane's statistics over and reload the bias slice
                if (persist && cur_img >= 0) dump_stats();
                epi_bar(epi_threads);                   // every warp is done with the previous tiles' s_stats / s_bias
                if (p.stats && cur_img >= 0) flush_stats(cur_img, cur_nt);
                if (spade && n_img != cur_img) {        // instance-norm constants of the modulated tensor's planes of this image
                    const int Cm = p.Cout / 2;
                    const double inv_hw = 1.0 / (double)npix;
                    for (int c = epi_tid; c < Cm; c += epi_threads) {
                        const double2 sq = *reinterpret_cast<const double2 *>(p.spade_stats + ((int64_t)n_img * Cm + c) * 2);
                        const double mean = sq.x * inv_hw;
                        double var = sq.y * inv_hw - mean * mean;
                        if (var < 0) var = 0;
                        s_mod[c] = (float)mean;
                        s_mod[1024 + c] = 1.0f / sqrtf((float)var + p.spade_eps);
                    }
                }
                if (nt != cur_nt && p.bias)
                    for (int i = epi_tid; i < BN; i += epi_threads) {
                        const int n = nt * BN + i;
                        s_bias[i] = n < p.Cout ? __ldg(p.bias + (pc ? n % pc : n)) : 0.f;
                    }
                cur_img = n_img; cur_nt = nt;
                epi_bar(epi_threads);
            }
            mbar_wait_backoff(smem_u32(&tfull_bar[acc]), use & 1, (uint32_t)P.backoff_ns);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (uint32_t)BN;
            uint32_t ra[16], rb[16];
            // The SPADE epilogue's activation tile and the residual are read once per (row, chunk) at pixel stride: their loads are
            // issued one chunk AHEAD (right after that chunk's tcgen05.ld), so the L2 latency overlaps the TMEM wait and the
            // previous chunk's arithmetic instead of stalling every chunk (long-scoreboard was the top stall of these epilogues).
            const bool res_vec = res && vec_ok;
            auto prefetch = [&](int ch, uint4 (&pf)[2]) {
                if (PERSIST) return;      // the register-persistent statistics kernels have no registers to spare (and no SPADE epilogue)
                if (!P.prefetch) return;
                const int n0 = nt * BN + ch * 16;
                if (!valid || n0 >= p.Cout) return;
                if (spade) pf[0] = *reinterpret_cast<const uint4 *>(static_cast<const T *>(p.spade_x) + m * p.ld_spade_x + (n0 >> 4) * 8);
                else if (res_vec && n0 + 16 <= p.Cout) {
                    const T *rr = res + m * p.ldr + n0;
                    pf[0] = *reinterpret_cast<const uint4 *>(rr); pf[1] = *reinterpret_cast<const uint4 *>(rr + 8);
                }
            };
            auto unpack8 = [](const uint4 &q, float (&u)[8]) {
                unpack2<T>(q.x, u[0], u[1]); unpack2<T>(q.y, u[2], u[3]); unpack2<T>(q.z, u[4], u[5]); unpack2<T>(q.w, u[6], u[7]);
            };
            uint4 pfa[2], pfb[2];
            auto finalize = [&](const uint32_t (&r)[16], int ch, auto slot_tag, const uint4 (&pf)[2]) {
                constexpr int slot = decltype(slot_tag)::value;      // which of this warp's (up to two) persistent statistics slots
                const int c0 = ch * 16;
                    const int n0 = nt * BN + c0;
                    if (n0 >= p.Cout) return;     // warp-uniform: padded output channels
                    int ncol = n0;
                    int64_t mm = m;
                    if (pc) {   // output parity (a,b) = phase: pixel (2gy+a, 2gx+b), channel n0 - phase*pc
                        const int ph = n0 / pc;
                        ncol = n0 - ph * pc;
                        mm = m + (int64_t)(ph >> 1) * p.OWf + ((ph & 1) ^ (ph >> 1));   // parity blocks are ordered (0,0),(0,1),(1,1),(1,0)
                    }
                    T *drow = dst + mm * p.ldd + ncol - c0;
                    const T *rrow = res ? res + mm * p.ldr + ncol - c0 : nullptr;
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 bv = *reinterpret_cast<const float4 *>(&s_bias[c0 + j]);
                            v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                        }
                    }
                    if (spade) {
                        // columns [g0..g7 b0..b7] of channels 8*(n0/16) .. +7: modulate the normalised activation, write 8 channels
                        const int cb = (n0 >> 4) * 8;
                        if (valid) {
                            float xv[8];
                            if (PERSIST || !P.prefetch) load8(static_cast<const T *>(p.spade_x) + m * p.ld_spade_x + cb, xv);
                            else unpack8(pf[0], xv);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float xn = (xv[j] - s_mod[cb + j]) * s_mod[1024 + cb + j];
                                xv[j] = fmaxf(fmaf(xn, 1.f + v[j], v[8 + j]), 0.f);
                            }
                            store8(dst + m * p.ldd + cb, xv);
                        }
                        return;
                    }
                    const bool full = n0 + 16 <= p.Cout;
                    if (res && valid) {
                        if (full && vec_ok) {
                            float u[8];
                            if (PERSIST || !P.prefetch) load8(rrow + c0, u);
                            else unpack8(pf[0], u);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] += u[j];
                            if (PERSIST || !P.prefetch) load8(rrow + c0 + 8, u);
                            else unpack8(pf[1], u);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[8 + j] += u[j];
                        } else {
                            for (int j = 0; j < 16; ++j)
                                if (n0 + j < p.Cout) v[j] += DT<T>::ld(rrow + c0 + j);
                        }
                    }
                    if (p.act_table) {
                        for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], p.act_table[min(n0 + j, p.Cout - 1)]);
                    } else if (p.act == HOIG_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (p.act != HOIG_ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], p.act);
                    }
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) pk[j] = pack2<T>(v[2 * j], v[2 * j + 1]);
                    if (P.st256 && full && !(P.debug & 8)) {
                        if (valid) st_global_v8(drow + c0, pk);        // a whole 32-byte sector per lane (sm_100 STG.256)
                    } else if (full && vec_ok && all_valid && !(P.debug & 8)) {
                        // Lane pairs swap one 16-byte piece so that the two lanes of a pair write the 32 contiguous bytes of ONE row per
                        // instruction (first the even lane's row, then the odd lane's): a store instruction then touches 16 lines
                        // with a full 32-byte sector each instead of 32 lines with half a sector -- half the LSU wavefronts.
                        const bool odd = lane & 1;
                        uint32_t keep[4], got[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t send = odd ? pk[j] : pk[4 + j];
                            keep[j] = odd ? pk[4 + j] : pk[j];
                            got[j] = __shfl_xor_sync(0xffffffffu, send, 1);
                        }
                        T *own = drow + c0 + (odd ? 8 : 0);
                        const uint64_t oth_bits = __shfl_xor_sync(0xffffffffu, (unsigned long long)reinterpret_cast<uintptr_t>(drow), 1);
                        T *oth = reinterpret_cast<T *>((uintptr_t)oth_bits) + c0 + (odd ? 8 : 0);
                        // even lane: row r <- own piece 0, row r+1 <- partner's piece 0;  odd lane: row r-1 <- partner's piece 1, row r <- own piece 1
                        *reinterpret_cast<uint4 *>(odd ? oth : own) = odd ? make_uint4(got[0], got[1], got[2], got[3]) : make_uint4(keep[0], keep[1], keep[2], keep[3]);
                        *reinterpret_cast<uint4 *>(odd ? own : oth) = odd ? make_uint4(keep[0], keep[1], keep[2], keep[3]) : make_uint4(got[0], got[1], got[2], got[3]);
                    } else if (valid && !(P.debug & 8)) {
                        if (full && vec_ok) {
                            *reinterpret_cast<uint4 *>(drow + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            *reinterpret_cast<uint4 *>(drow + c0 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        } else {
                            for (int j = 0; j < 16; ++j)
                                if (n0 + j < p.Cout) DT<T>::st(drow + c0 + j, v[j]);
                        }
                    }
                    if (p.stats && !(P.debug & 4)) {
                        if (P.mma_stats) {
                            // Column sums of the stored 16-bit values and of their squares (rounded to bf16) on the warp-level
                            // tensor-core path (umma_common.cuh colsum16): ~60 instructions instead of ~140 per chunk.
                            constexpr bool kF16 = std::is_same<T, __half>::value;
                            uint32_t sq[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {     // squares of the fp32 values, rounded to bf16 (range of fp32)
                                if (!all_valid && !valid) { pk[j] = 0u; v[2 * j] = 0.f; v[2 * j + 1] = 0.f; }
                                sq[j] = pack2<__nv_bfloat16>(v[2 * j] * v[2 * j], v[2 * j + 1] * v[2 * j + 1]);
                            }
                            if (persist) {
                                // narrow tiles: this warp owns the same (at most two) 16-column chunks on every tile, so the warp-level
                                // MMAs keep accumulating in registers; shuffles + shared-memory adds happen once per image (dump_stats)
                                colsum16_acc<kF16>(pk, e_sel, st_sum[slot < NSLOT ? slot : 0]);
                                colsum16_acc<false>(sq, e_sel_bf, st_sq[slot < NSLOT ? slot : 0]);
                            } else {
                            float s_lo, s_hi, q_lo, q_hi;
                            colsum16<kF16>(pk, e_sel, s_lo, s_hi);
                            colsum16<false>(sq, e_sel_bf, q_lo, q_hi);
                            if ((lane & 3) == 0) {      // (quad, column) is touched by this warp only
                                const int col = c0 + (lane >> 2);
                                s_stats[quad][0][col] += s_lo; s_stats[quad][0][col + 8] += s_hi;
                                s_stats[quad][1][col] += q_lo; s_stats[quad][1][col + 8] += q_hi;
                            }
                            }
                        } else {
                        // Statistics of the fp32 values (before the 16-bit store rounding); rows past the end of the plane
                        // contribute nothing.
                        float q[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (!all_valid) v[j] = valid ? v[j] : 0.f;
                            q[j] = v[j] * v[j];
                        }
                        const float cs = transpose_reduce16(v, lane);
                        const float cq = transpose_reduce16(q, lane);
                        if ((lane & 1) == 0) {
                            const int col = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                            s_stats[quad][0][col] += cs;   // (quad, column) is touched by this warp only
                            s_stats[quad][1][col] += cq;
                        }
                        }
                    }
            };
            if constexpr (FAST) {
                // Streamlined epilogue (launch_one: no residual / SPADE / per-channel activations, 16-byte aligned rows, whole 16- or
                // 32-column chunks).  Its own kernel instantiation, so neither the general epilogue's code nor its registers are carried.
                // P.fast_epi = 2: 32-column chunks -- the lane writes 64 contiguous bytes (two full sectors, no lane-pair exchange) and the
                // per-chunk address / bounds code runs once per 32 columns; 1: 16-column chunks with the lane-pair store of the general path.
                T *const drow0 = dst + m * p.ldd;
                const bool relu = p.act == HOIG_ACT_RELU;
                auto fast_chunk = [&](auto wtag, int ch, auto slot_tag) {
                    constexpr int W = decltype(wtag)::value;
                    uint32_t r[W];
                    if constexpr (W == 32) { tmem_ld32(t_row + (uint32_t)(ch * W), r); tmem_ld_wait32(r); }
                    else { tmem_ld16(t_row + (uint32_t)(ch * W), r); tmem_ld_wait(r); }
                    const int c0 = ch * W, n0 = nt * BN + c0;
                    if (n0 >= p.Cout) return;         // warp-uniform: padded output channels
                    T *drow = drow0 + n0;
                    if (pc) {       // transposed conv: parity block ph of pc channels, blocks ordered (0,0),(0,1),(1,1),(1,0)
                        const int ph = n0 / pc;
                        drow = dst + (m + (int64_t)(ph >> 1) * p.OWf + ((ph & 1) ^ (ph >> 1))) * p.ldd + (n0 - ph * pc);
                    }
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < W; j += 4) {
                            const float4 bv = *reinterpret_cast<const float4 *>(&s_bias[c0 + j]);
                            r[j] = __float_as_uint(__uint_as_float(r[j]) + bv.x);
                            r[j + 1] = __float_as_uint(__uint_as_float(r[j + 1]) + bv.y);
                            r[j + 2] = __float_as_uint(__uint_as_float(r[j + 2]) + bv.z);
                            r[j + 3] = __float_as_uint(__uint_as_float(r[j + 3]) + bv.w);
                        }
                    }
                    if (relu) {
#pragma unroll
                        for (int j = 0; j < W; ++j) r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]), 0.f));
                    }
#pragma unroll
                    for (int hh = 0; hh < W / 16; ++hh) {       // 16-column pieces: pack, store, statistics
                        constexpr int slot0 = decltype(slot_tag)::value;
                        const int slot = W == 32 ? hh : slot0;   // PERSIST: a 32-column chunk is the warp's only one (slots 0, 1)
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) pk[j] = pack2<T>(__uint_as_float(r[hh * 16 + 2 * j]), __uint_as_float(r[hh * 16 + 2 * j + 1]));
                        if (P.debug & 8) {
                        } else if (P.st256) {
                            if (valid) st_global_v8(drow + hh * 16, pk);      // one whole 32-byte sector per lane
                        } else if (W == 32 || !all_valid) {
                            if (valid) {
                                *reinterpret_cast<uint4 *>(drow + hh * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                *reinterpret_cast<uint4 *>(drow + hh * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                            }
                        } else {
                            // lane pairs swap one 16-byte piece: each store instruction writes 16 full 32-byte sectors (see the general path)
                            const bool odd = lane & 1;
                            uint32_t keep[4], got[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t send = odd ? pk[j] : pk[4 + j];
                                keep[j] = odd ? pk[4 + j] : pk[j];
                                got[j] = __shfl_xor_sync(0xffffffffu, send, 1);
                            }
                            T *own = drow + (odd ? 8 : 0);
                            const uint64_t oth_bits = __shfl_xor_sync(0xffffffffu, (unsigned long long)reinterpret_cast<uintptr_t>(drow), 1);
                            T *oth = reinterpret_cast<T *>((uintptr_t)oth_bits) + (odd ? 8 : 0);
                            *reinterpret_cast<uint4 *>(odd ? oth : own) = odd ? make_uint4(got[0], got[1], got[2], got[3]) : make_uint4(keep[0], keep[1], keep[2], keep[3]);
                            *reinterpret_cast<uint4 *>(odd ? own : oth) = odd ? make_uint4(keep[0], keep[1], keep[2], keep[3]) : make_uint4(got[0], got[1], got[2], got[3]);
                        }
                        if (p.stats && !(P.debug & 4)) {
                            const int cc = c0 + hh * 16;
                            if (P.mma_stats) {
                                constexpr bool kF16 = std::is_same<T, __half>::value;
                                uint32_t sq[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    float a = __uint_as_float(r[hh * 16 + 2 * j]), b = __uint_as_float(r[hh * 16 + 2 * j + 1]);
                                    if (!all_valid && !valid) { pk[j] = 0u; a = 0.f; b = 0.f; }
                                    sq[j] = pack2<__nv_bfloat16>(a * a, b * b);
                                }
                                if (persist) {
                                    colsum16_acc<kF16>(pk, e_sel, st_sum[slot < NSLOT ? slot : 0]);
                                    colsum16_acc<false>(sq, e_sel_bf, st_sq[slot < NSLOT ? slot : 0]);
                                } else {
                                    float s_lo, s_hi, q_lo, q_hi;
                                    colsum16<kF16>(pk, e_sel, s_lo, s_hi);
                                    colsum16<false>(sq, e_sel_bf, q_lo, q_hi);
                                    if ((lane & 3) == 0) {
                                        const int col = cc + (lane >> 2);
                                        s_stats[quad][0][col] += s_lo; s_stats[quad][0][col + 8] += s_hi;
                                        s_stats[quad][1][col] += q_lo; s_stats[quad][1][col + 8] += q_hi;
                                    }
                                }
                            } else {
                                float v[16], q[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    v[j] = (all_valid || valid) ? __uint_as_float(r[hh * 16 + j]) : 0.f;
                                    q[j] = v[j] * v[j];
                                }
                                const float cs = transpose_reduce16(v, lane);
                                const float cq = transpose_reduce16(q, lane);
                                if ((lane & 1) == 0) {
                                    const int col = cc + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                                    s_stats[quad][0][col] += cs;
                                    s_stats[quad][1][col] += cq;
                                }
                            }
                        }
                    }
                };
                if (P.fast_epi == 2) {
                    for (int ch = half; ch < BN / 32; ch += ngrp) fast_chunk(std::integral_constant<int, 32>(), ch, std::integral_constant<int, 0>());
                } else {
                    for (int ch = half; ch < n_chunks; ch += 2 * ngrp) {
                        fast_chunk(std::integral_constant<int, 16>(), ch, std::integral_constant<int, 0>());
                        if (ch + ngrp < n_chunks) fast_chunk(std::integral_constant<int, 16>(), ch + ngrp, std::integral_constant<int, 1>());
                    }
                }
                tc_fence_before();
                if (NCTA == 2) mbar_arrive_cluster_relaxed(tempty0 + 8u * acc);
                else mbar_arrive_relaxed(tempty0 + 8u * acc);
            } else {
            if (P.debug & 1) {
                for (int ch = half; ch < n_chunks; ch += ngrp) { tmem_ld16(t_row + (uint32_t)(ch * 16), ra); tmem_ld_wait(ra); }
                if (ra[0] == 0x7fc12345u && valid) dst[m * p.ldd] = T(0);
            } else {
            // HOIG_UMMA_EARLY_RELEASE (P.early_release): 1 = the accumulator stage goes back to the MMA warp as soon as this warp's LAST
            // tcgen05.ld has landed in registers (before the arithmetic and stores of that chunk), 0 = after the whole tile.  Either
            // way the arrival is relaxed: it only signals, so nothing waits for the outstanding global stores.
            auto release = [&]() {
                tc_fence_before();
                if (NCTA == 2) mbar_arrive_cluster_relaxed(tempty0 + 8u * acc);
                else mbar_arrive_relaxed(tempty0 + 8u * acc);
            };
            const bool early = P.early_release != 0;
            if (half < n_chunks) { tmem_ld16(t_row + (uint32_t)(half * 16), ra); prefetch(half, pfa); }
            else if (early) release();
            for (int ch = half; ch < n_chunks; ch += 2 * ngrp) {
                tmem_ld_wait(ra);
                const bool more1 = ch + ngrp < n_chunks;
                if (more1) { tmem_ld16(t_row + (uint32_t)((ch + ngrp) * 16), rb); prefetch(ch + ngrp, pfb); }
                else if (early) release();
                finalize(ra, ch, std::integral_constant<int, 0>(), pfa);
                if (more1) {
                    tmem_ld_wait(rb);
                    const bool more2 = ch + 2 * ngrp < n_chunks;
                    if (more2) { tmem_ld16(t_row + (uint32_t)((ch + 2 * ngrp) * 16), ra); prefetch(ch + 2 * ngrp, pfa); }
                    else if (early) release();
                    finalize(rb, ch + ngrp, std::integral_constant<int, 1>(), pfb);
                }
            }
            if (!early) {
                tc_fence_before();
                if (P.early_release == 0 && P.relaxed_release) { if (NCTA == 2) mbar_arrive_cluster_relaxed(tempty0 + 8u * acc); else mbar_arrive_relaxed(tempty0 + 8u * acc); }
                else { if (NCTA == 2) mbar_arrive_cluster(tempty0 + 8u * acc); else mbar_arrive(tempty0 + 8u * acc); }
            }
            continue;
            }
            // (debug drain path) accumulator drained: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            if (NCTA == 2) mbar_arrive_cluster(tempty0 + 8u * acc);
            else mbar_arrive(tempty0 + 8u * acc);
            }   // !FAST
        }
        if (p.stats && cur_img >= 0) {   // the last plane this CTA touched
            if (persist) dump_stats();
            epi_bar(epi_threads);
            flush_stats(cur_img, cur_nt);
        }
    }

    tc_fence_before();
    if (NCTA == 2) cluster_sync();   // neither CTA may exit (or free TMEM) while its peer can still touch its smem / barriers
    else __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

int g_umma_debug = 0;

int g_st256 = 7;            // HOIG_UMMA_ST256 (bit mask): 256-bit epilogue stores in 1 = the FAST kernels, 2 = the general epilogue, 4 = conv_halo
int g_contig_mode = 1;      // HOIG_UMMA_CONTIG
int g_fast_epi = 1;         // HOIG_UMMA_FAST_EPI: streamlined 32-column epilogue for plain convs
int g_mma_stats = 1;        // HOIG_UMMA_MMA_STATS
int g_halo_mode = 1;        // HOIG_UMMA_HALO: row-halo activation reuse for full-row tiles
int g_small_split_mode = 1; // HOIG_UMMA_SMALL_SPLIT: narrower n-tiles when a launch has fewer work units than half the SMs (small batches)
int g_tap_skip_mode = 1;    // HOIG_UMMA_TAP_SKIP: transposed convs skip (n-tile, tap) k-blocks whose weight blocks are structurally zero
int g_backoff_ns = 0;       // HOIG_UMMA_BACKOFF_NS
int g_prefetch = 1;         // HOIG_UMMA_PREFETCH: 0 = never, 1 = SPADE epilogue only (default), 2 = SPADE and residual
int g_early_release = 0;    // HOIG_UMMA_EARLY_RELEASE
int g_relaxed_release = 1;  // HOIG_UMMA_RELAXED_RELEASE
int g_vhalo_mode = 1;       // HOIG_UMMA_VHALO: vertical-halo activation reuse for kh x 1 convs
int g_bres_mode = 1;        // HOIG_UMMA_BRES: resident weights for small weight matrices
int g_dual_mode = 1;        // 1: narrow-N TMA convs run two MMA issue pipelines per CTA (HOIG_UMMA_DUAL=0 disables)
int g_pair_mode = 1;        // 0: one CTA per tile; 1: CTA pairs (cta_group::2) where they pay off; 2: pairs wherever legal (tests)

template <typename T, int NCTA, bool PERSIST, bool FAST>
int launch_kernel(const UmmaParams &P, const CUtensorMap &map_w, const CUtensorMap *map_a, int grid, size_t smem, cudaStream_t stream)
{
    constexpr int slot = (std::is_same<T, __half>::value ? SLOT_CONV_UMMA_F16_1 : SLOT_CONV_UMMA_BF16_1) + 2 * (NCTA - 1) + (PERSIST ? 1 : 0) +
                         (FAST ? (int)SLOT_CONV_UMMA_FAST_FIRST - (int)SLOT_CONV_UMMA_BF16_1 : 0);
    if (first_use_on_device(slot) &&
        cudaFuncSetAttribute(conv_umma_kernel<T, NCTA, PERSIST, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL) != cudaSuccess)
        return check_launch("conv_umma smem attribute");
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = NCTA > 1 ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, conv_umma_kernel<T, NCTA, PERSIST, FAST>, P, map_w, map_a[0], map_a[1], map_a[2], map_a[3]) != cudaSuccess)
        return check_launch("conv_umma_kernel launch");
    return check_launch("conv_umma_kernel");
}

int launch_one(const ConvParams &cp, cudaStream_t stream, int force_gather, int dtype)
{
    UmmaParams P;
    P.c = cp;
    const ConvParams &p = P.c;
    HOIG_REQUIRE(p.Npad <= 4096, "conv2d: Cout too large");
    P.BN = p.Npad <= 256 ? p.Npad : 256;
    P.m_tiles = p.N * p.tiles_per_image;
    // Small batches (the eval.py case, batch 1): a 32x32 layer has only 8 pixel tiles, so wide n-tiles leave most SMs idle.
    // Narrower n-tiles multiply the number of work units (each re-reads the L2-resident activations): 512 -> 512 at batch 1 runs on
    // 64 SMs instead of 16.  Never taken at batch >= 8 (the unit count is already above half the SM count).
    if (g_small_split_mode)
        while (P.BN > 64 && (P.BN / 2) % 16 == 0 && p.Npad % (P.BN / 2) == 0 &&
               (int64_t)P.m_tiles * ceil_div(p.Npad, P.BN) < device_sm_count() / 2)
            P.BN /= 2;
    P.n_tiles = ceil_div(p.Npad, P.BN);
    P.k_blocks = p.Kpad / BK;
    P.cpt_shift = -1;
    if (p.Cin < 64) {
        const int cpt = p.Cin / 8;
        if ((cpt & (cpt - 1)) == 0) { P.cpt_shift = 0; while ((1 << P.cpt_shift) < cpt) ++P.cpt_shift; }
    }

    const int64_t npix = (int64_t)p.GH * p.GW;
    const bool rect = (p.GW % BM == 0) || (BM % p.GW == 0);
    // (stride-2 convs arrive here as stride-1 convs over four input-parity views, see conv_plan.cu)
    P.tma_a = (!force_gather && p.mode == HOIG_CONV && p.stride == 1 && p.C1 == 0 && p.C0 % 64 == 0 && rect && npix % BM == 0) ? 1 : 0;
    if (P.tma_a)   // view strides must be 16-byte multiples for a tensor map
        for (int v = 0; v < p.nviews; ++v)
            if ((p.view[v].sx * 2) % 16 || ((uintptr_t)p.view[v].base % 16)) P.tma_a = 0;
    // CTA pairs: TMA-fed activations, an even number of 128-pixel tiles, and a weight tile that splits into two
    // halves of whole 16-row groups
    // Short reductions stay on single CTAs: a pair pays cluster-scope barrier latency per tile, measured to cost more than
    // the halved weight traffic saves below ~12 k-blocks (7x1 stems, 64->128 stride-2, 128->64 transposed).
    const bool pair_ok = P.tma_a && P.m_tiles % 2 == 0 && P.BN % 32 == 0 && p.Npad % P.BN == 0;
    int ncta = (pair_ok && (g_pair_mode == 2 || (g_pair_mode == 1 && P.k_blocks >= 12))) ? 2 : 1;
    // Vertical-halo mode for the kh x 1 convs (7x7 stems / heads after re-association): single CTAs, resident weights
    const int vh_a_bytes = (VH_ROWS + p.kh - 1) * VH_COLS * 128;
    P.vhalo = (g_vhalo_mode && g_pair_mode != 2 && P.tma_a && p.nviews == 1 && p.kw == 1 && p.kh >= 2 && p.kh <= 8 && p.GW % VH_COLS == 0 &&
               p.GH % VH_ROWS == 0 && p.Cin % BK == 0 && P.n_tiles == 1 && p.Kpad == p.kh * p.Cin &&
               (size_t)P.k_blocks * P.BN * BK * 2 + 4 * (size_t)vh_a_bytes <= (size_t)RING_BUDGET) ? 1 : 0;
    P.tiles_x = P.vhalo ? p.GW / VH_COLS : 0;
    if (P.vhalo) ncta = 1;
    // Resident weights (single-CTA tiles, one n-tile): worth it when the whole matrix fits beside >= 4 activation stages
    // Row-halo mode: regular stride-1 kh x kw conv, tiles = 128 consecutive pixels of one image row
    P.halo = (g_halo_mode && P.tma_a && p.nviews == 1 && p.kw >= 2 && p.kw <= 7 && p.GW % BM == 0 && p.Cin % BK == 0) ? 1 : 0;
    if (P.halo && RING_BUDGET / (A_HALO_BYTES + p.kw * (P.BN / ncta) * BK * 2) < 3) P.halo = 0;   // needs >= 3 stages of kw weight tiles
    P.k_steps = P.halo ? p.kh * (p.Cin / BK) : (P.vhalo ? p.Cin / BK : 0);
    const size_t w_bytes = (size_t)P.k_blocks * P.BN * BK * 2;
    // (not when the conv would run on CTA pairs: measured slower for 128->64 @256^2, whose 147 KB of weights leave 4 stages)
    P.bres = (P.vhalo || (g_bres_mode && !P.halo && P.tma_a && P.n_tiles == 1 && ncta == 1 && w_bytes + 4 * (size_t)A_STAGE_BYTES <= (size_t)RING_BUDGET)) ? 1 : 0;
    const int stage_bytes = P.vhalo ? vh_a_bytes : P.bres ? A_STAGE_BYTES
                                   : (P.halo ? A_HALO_BYTES + p.kw * (P.BN / ncta) * BK * 2 : A_STAGE_BYTES + (P.BN / ncta) * BK * 2);
    P.stages = (int)((RING_BUDGET - (P.bres ? w_bytes : 0)) / stage_bytes);
    if (P.stages > MAX_STAGES) P.stages = MAX_STAGES;
    // Narrow tiles (N <= 128): one thread cannot issue 128 x N x 16 MMAs as fast as the tensor pipe retires them
    // (scripts/probes/mma_probe.cu: ~65-100 cycles of issue overhead vs 48-64 cycles of execution), so two warps
    // issue, each driving its own pipeline (half of the ring, two of four accumulator stages, alternate tiles).
    P.contig = g_contig_mode;
    P.dual = (g_dual_mode && P.tma_a && P.BN <= 128 && P.stages >= 4) ? 1 : 0;
    if (P.dual) P.stages &= ~1;
    HOIG_REQUIRE(P.stages >= (P.tma_a ? 2 : LOOKAHEAD + 1), "conv2d: not enough shared memory stages");

    // transposed conv: which taps feed which n-tile (packing.parity_block order (0,0),(0,1),(1,1),(1,0); tap (dy,dx) feeds (a,b) iff
    // dy <= a and dx <= b)
    P.tap_skip = 0;
    memset(P.tap_live, 0xff, sizeof(P.tap_live));
    if (g_tap_skip_mode && p.phase_cout > 0 && P.tma_a && p.ntaps == 4 && P.n_tiles <= 16 && p.Cin % BK == 0) {
        P.tap_skip = 1;
        for (int nt = 0; nt < P.n_tiles; ++nt) {
            const int j0 = (nt * P.BN) / p.phase_cout;
            int j1 = ((nt + 1) * P.BN - 1) / p.phase_cout;
            if (j1 > 3) j1 = 3;
            uint8_t mask = 0;
            for (int j = j0; j <= j1; ++j) {
                const int a = j >> 1, b = (j & 1) ^ (j >> 1);
                for (int t = 0; t < 4; ++t)
                    if ((t >> 1) <= a && (t & 1) <= b) mask |= (uint8_t)(1u << t);
            }
            P.tap_live[nt] = mask;
        }
    }
    P.debug = g_umma_debug;
    P.early_release = g_early_release;
    P.backoff_ns = P.k_blocks >= 16 ? g_backoff_ns : 0;      // only where a tile's main loop is long
    P.prefetch = g_prefetch == 2 ? 1 : (g_prefetch == 1 && p.spade_x ? 1 : 0);   // measured: SPADE epilogue -2.6 %, residual epilogue +1.8 % (off there)
    P.relaxed_release = g_relaxed_release;
    P.tpi_shift = -1;
    if ((p.tiles_per_image & (p.tiles_per_image - 1)) == 0) { P.tpi_shift = 0; while ((1 << P.tpi_shift) < p.tiles_per_image) ++P.tpi_shift; }

    CUtensorMap map_w, map_a[4];
    int st;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)p.Kpad, (cuuint64_t)p.Npad};
        const cuuint64_t strides[1] = {(cuuint64_t)p.ldw * 2};
        const cuuint32_t box[2] = {BK, (cuuint32_t)(P.BN / ncta)};
        st = make_map(&map_w, p.weight, 2, dims, strides, box, "weights", dtype);
        if (st != HOIG_OK) return st;
    }
    for (int v = 0; v < 4; ++v) map_a[v] = map_w;
    if (P.tma_a) {
        const int bw = P.vhalo ? VH_COLS : P.halo ? BM + p.kw - 1 : (p.GW < BM ? p.GW : BM);
        const int bh = P.vhalo ? VH_ROWS + p.kh - 1 : P.halo ? 1 : BM / bw;
        const cuuint32_t box[4] = {BK, (cuuint32_t)bw, (cuuint32_t)bh, 1};
        for (int v = 0; v < p.nviews; ++v) {
            const InputView &vw = p.view[v];
            const cuuint64_t dims[4] = {(cuuint64_t)p.C0, (cuuint64_t)vw.W, (cuuint64_t)vw.H, (cuuint64_t)p.N};
            const cuuint64_t strides[3] = {(cuuint64_t)vw.sx * 2, (cuuint64_t)vw.sy * 2, (cuuint64_t)vw.sn * 2};
            st = make_map(&map_a[v], vw.base, 4, dims, strides, box, "activations", dtype);
            if (st != HOIG_OK) return st;
        }
    }

    const int num_sms = device_sm_count();
    const int total = (P.m_tiles / ncta) * P.n_tiles;
    const int units = num_sms / ncta;                       // CTAs or CTA pairs that fit the GPU
    const int grid = (total < units ? total : units) * ncta;
    const size_t smem = (size_t)STG_BYTES + (size_t)P.stages * stage_bytes + (P.bres ? w_bytes : 0) + 1024;
    // persistent register statistics: at most two 16-column chunks per epilogue warp, one n-tile, MMA column sums enabled
    const int ngrp = P.tma_a ? 4 : 3;
    const bool persist = p.stats && g_mma_stats && P.n_tiles == 1 && P.BN / 16 <= 2 * ngrp && !p.spade_x;
    // Statistics method (HOIG_UMMA_MMA_STATS: 0 = shuffle transpose-reduce everywhere, 1 = auto, 2 = warp-level MMAs everywhere).
    // Auto: the warp-level MMA column sums only where they keep accumulating in registers (persist).  Elsewhere -- the wide / long-K
    // convs -- legacy HMMA shares the tensor pipe with the UMMAs that bound those convs, while their issue slots are idle: the shuffle
    // version measured 2 % (512->512) to 18 % (256->128 @128^2) faster there.
    P.mma_stats = g_mma_stats == 2 ? 1 : (g_mma_stats == 1 ? (persist ? 1 : 0) : 0);
    // Streamlined epilogue (HOIG_UMMA_FAST_EPI; kernels instantiated with FAST): plain TMA-fed convs -- optional bias, no / ReLU activation,
    // optional statistics -- with whole 16-column chunks.  32-column chunks where every epilogue warp gets at least one (exactly one in
    // PERSIST kernels).  Measured (batch 64, fp16): stem -10 %, 128->64 @256^2 -6 %, ConvT 128->64 -13 %, mlp_shared GEMMs -18..-27 %
    // (and, with the 256-bit stores, stride-2 64->128 -12 %).
    P.fast_epi = 0;
    P.st256 = (g_st256 != 0 && !p.spade_x && p.ldd % 16 == 0 && reinterpret_cast<uintptr_t>(p.dst) % 32 == 0 &&
               (!p.residual || p.ldr % 8 == 0)) ? 1 : 0;     // also the general epilogue's full chunks
    if (g_fast_epi && P.tma_a && !p.residual && !p.spade_x && !p.act_table && (p.act == HOIG_ACT_NONE || p.act == HOIG_ACT_RELU) && p.ldd % 8 == 0 &&
        reinterpret_cast<uintptr_t>(p.dst) % 16 == 0 && p.Cout % 16 == 0 && (p.phase_cout == 0 || p.phase_cout % 16 == 0)) {
        const bool wide = P.BN % 32 == 0 && P.BN / 32 >= ngrp && p.Cout % 32 == 0 && (p.phase_cout == 0 || p.phase_cout % 32 == 0) &&
                          (!persist || P.BN / 32 == ngrp);
        P.fast_epi = wide ? 2 : 1;
        if (g_fast_epi == 2 && !wide) P.fast_epi = 0;       // HOIG_UMMA_FAST_EPI=2: only the 32-column form (A/B runs)
    }
    if (!((P.fast_epi ? 1 : 2) & g_st256)) P.st256 = 0;
    auto go = [&](auto tag, auto nc, auto ps) {
        using T = std::remove_pointer_t<decltype(tag)>;
        return P.fast_epi ? launch_kernel<T, decltype(nc)::value, decltype(ps)::value, true>(P, map_w, map_a, grid, smem, stream)
                          : launch_kernel<T, decltype(nc)::value, decltype(ps)::value, false>(P, map_w, map_a, grid, smem, stream);
    };
    auto pick_ps = [&](auto tag, auto nc) {
        return persist ? go(tag, nc, std::true_type()) : go(tag, nc, std::false_type());
    };
    auto pick_nc = [&](auto tag) {
        return ncta == 2 ? pick_ps(tag, std::integral_constant<int, 2>()) : pick_ps(tag, std::integral_constant<int, 1>());
    };
    return dtype == HOIG_F16 ? pick_nc((__half *)nullptr) : pick_nc((__nv_bfloat16 *)nullptr);
}

}  // namespace

int g_force_gather = -1;

int conv2d_umma(const hoigConvDesc *d, cudaStream_t stream)
{
    ConvPlan plan;
    int st = plan_conv(d, BM, &plan);
    if (st != HOIG_OK) return st;
    static bool env_parsed = false;       // independent of the setters below: calling one of them first must not skip this block
    if (!env_parsed) {
        env_parsed = true;
        const char *e = getenv("HOIG_UMMA_GATHER_ONLY");
        if (g_force_gather < 0) g_force_gather = (e && e[0] == '1') ? 1 : 0;
        const char *ms = getenv("HOIG_UMMA_MMA_STATS");
        if (ms) g_mma_stats = atoi(ms);
        const char *hm = getenv("HOIG_UMMA_HALO");
        if (hm) g_halo_mode = atoi(hm);
        const char *ss = getenv("HOIG_UMMA_SMALL_SPLIT");
        if (ss) g_small_split_mode = atoi(ss);
        const char *ts = getenv("HOIG_UMMA_TAP_SKIP");
        if (ts) g_tap_skip_mode = atoi(ts);
        const char *bo = getenv("HOIG_UMMA_BACKOFF_NS");
        if (bo) g_backoff_ns = atoi(bo);
        const char *pf = getenv("HOIG_UMMA_PREFETCH");
        if (pf) g_prefetch = atoi(pf);
        const char *er = getenv("HOIG_UMMA_EARLY_RELEASE");
        if (er) g_early_release = atoi(er);
        const char *rr = getenv("HOIG_UMMA_RELAXED_RELEASE");
        if (rr) g_relaxed_release = atoi(rr);
        const char *vm = getenv("HOIG_UMMA_VHALO");
        if (vm) g_vhalo_mode = atoi(vm);
        const char *bm = getenv("HOIG_UMMA_BRES");
        if (bm) g_bres_mode = atoi(bm);
        const char *cm = getenv("HOIG_UMMA_CONTIG");
        if (cm) g_contig_mode = atoi(cm);
        const char *s2 = getenv("HOIG_UMMA_ST256");
        if (s2) g_st256 = atoi(s2);
        const char *fe = getenv("HOIG_UMMA_FAST_EPI");
        if (fe) g_fast_epi = atoi(fe);
        const char *dm = getenv("HOIG_UMMA_DUAL");
        if (dm) g_dual_mode = atoi(dm);
        const char *pm = getenv("HOIG_UMMA_2CTA");
        if (pm) g_pair_mode = atoi(pm);
        const char *dbg = getenv("HOIG_UMMA_DEBUG");
        if (dbg) g_umma_debug = atoi(dbg);
    }
    for (int i = 0; i < plan.n; ++i) {
        st = launch_one(plan.launch[i], stream, g_force_gather, d->dtype);
        if (st != HOIG_OK) return st;
    }
    return HOIG_OK;
}

}  // namespace hoig

// Diagnostic switch: route every bf16 conv through the cp.async gather A-operand path
// (1) or let eligible convs use TMA boxes (0).
extern "C" void hoig_set_umma_gather_only(int on) { hoig::g_force_gather = on ? 1 : 0; }
// Diagnostic switch: 0 = one CTA per tile, 1 = CTA pairs (cta_group::2) where they pay off (default), 2 = wherever legal.
extern "C" void hoig_set_umma_pair_mode(int on) { hoig::g_pair_mode = on; }
// Diagnostic switch: 1 = full-row tiles of regular stride-1 convs share one activation box per kernel row (default), 0 = one box per tap.
extern "C" void hoig_set_umma_halo_mode(int on) { hoig::g_halo_mode = on ? 1 : 0; }
// Diagnostic switch: 1 = kh x 1 convs share one activation box per 64 channels across their vertical taps (default), 0 = one box per tap.
extern "C" void hoig_set_umma_vhalo_mode(int on) { hoig::g_vhalo_mode = on ? 1 : 0; }
// Diagnostic switch: 1 = small weight matrices stay resident in shared memory (default), 0 = always streamed.
extern "C" void hoig_set_umma_bres_mode(int on) { hoig::g_bres_mode = on ? 1 : 0; }
// Diagnostic switch: 1 = two MMA issue pipelines per CTA for narrow-N tiles (default), 0 = one.
extern "C" void hoig_set_umma_dual_mode(int on) { hoig::g_dual_mode = on ? 1 : 0; }
