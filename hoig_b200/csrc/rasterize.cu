// rasterize.cu -- stage R: bit-exact condition rasterizer + condition-map kernels.
//
// Replaces (paths relative to /root/reference/HOIG_HOv3):
//   thirdparty/neural_renderer/neural_renderer/cuda/rasterize_cuda_kernel.cu:41-186
//   (forward_face_index_map kernels _1/_2), rasterize.py:50-52 (output init),
//   rasterize.py:335-338 (vertical flip).
//
// Design (B200-first, not the reference's per-pixel x all-faces loop):
//   * face-parallel scatter.  One CTA owns (mesh, band of rows); its z-buffer
//     lives in shared memory as 64-bit keys (ordered depth bits << 32 | face
//     index) resolved with shared-memory atomicMin, so the winner is the
//     minimum depth with the LOWEST face index on ties -- exactly the result
//     of the reference's in-order strict '<' z-test.
//   * Candidate pixels of a face are found EXACTLY, with no epsilon: each of
//     the reference's three edge tests compares A(y) = (yp-ya)*dx against
//     B(x) = (xp-xa)*dy; IEEE rounding is monotone, so A is monotone in the
//     row index and B in the column index.  The set of rows that can pass and,
//     per row, the column interval that passes are therefore found by binary
//     search on the reference's own float predicate.  Every pixel inside the
//     interval passes all three tests bit-identically to the reference.
//   * Per-(pixel,face) arithmetic replays the reference's compiled op
//     sequence with explicit __f*_rn intrinsics, which
//     nvcc never contracts.
//   * A second phase of the same kernel turns keys into fim / wim / depth and
//     writes them coalesced (flip fused).
#include "common.cuh"

namespace hoig {
namespace {

constexpr int kRastThreads = 512;
constexpr int kMaxBandPixels = 16384;  // 128 KB of 64-bit keys
constexpr unsigned long long kEmptyKey = 0xffffffffffffffffull;
constexpr float kWild = 1e15f;

struct FaceInv { float v[9]; };

// rasterize_cuda_kernel.cu:62-79 with the contraction pattern nvcc emits.
__device__ __forceinline__ void face_inverse(const float f[9], float isf, float fi[9])
{
    float p[3][2];
#pragma unroll
    for (int n = 0; n < 3; ++n)
#pragma unroll
        for (int d = 0; d < 2; ++d)
            p[n][d] = __fmul_rn(__fadd_rn(__fmaf_rn(f[3 * n + d], isf, isf), -1.0f), 0.5f);
    fi[0] = __fsub_rn(p[1][1], p[2][1]);
    fi[1] = __fsub_rn(p[2][0], p[1][0]);
    fi[2] = __fmaf_rn(p[1][0], p[2][1], -__fmul_rn(p[2][0], p[1][1]));
    fi[3] = __fsub_rn(p[2][1], p[0][1]);
    fi[4] = __fsub_rn(p[0][0], p[2][0]);
    fi[5] = __fmaf_rn(p[2][0], p[0][1], -__fmul_rn(p[0][0], p[2][1]));
    fi[6] = __fsub_rn(p[0][1], p[1][1]);
    fi[7] = __fsub_rn(p[1][0], p[0][0]);
    fi[8] = __fmaf_rn(p[0][0], p[1][1], -__fmul_rn(p[1][0], p[0][1]));
    const float den = __fmaf_rn(p[1][0], fi[3], __fmaf_rn(p[2][0], fi[6], __fmul_rn(p[0][0], fi[0])));
#pragma unroll
    for (int k = 0; k < 9; ++k) fi[k] = __fdiv_rn(fi[k], den);
}

// rasterize_cuda_kernel.cu:57 / :128
__device__ __forceinline__ bool back_facing(const float f[9])
{
    return __fmul_rn(__fsub_rn(f[7], f[1]), __fsub_rn(f[3], f[0])) <
           __fmul_rn(__fsub_rn(f[4], f[1]), __fsub_rn(f[6], f[0]));
}

// rasterize_cuda_kernel.cu:139-153.  Returns true and (w, zp) when the pixel survives
// the near/far test; NaN zp never survives (it fails the later zp < depth_min).
__device__ __forceinline__ bool shade(const float f[9], const float fi[9], float xf, float yf, float near_, float far_,
                                      float w[3], float &zp)
{
#pragma unroll
    for (int k = 0; k < 3; ++k)
        w[k] = __fadd_rn(__fmaf_rn(fi[3 * k], xf, __fmul_rn(fi[3 * k + 1], yf)), fi[3 * k + 2]);
    float ws = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        w[k] = fminf(fmaxf(w[k], 0.f), 1.f);
        ws = __fadd_rn(ws, w[k]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = __fdiv_rn(w[k], ws);
    zp = __frcp_rn(__fadd_rn(__fadd_rn(__fdiv_rn(w[0], f[2]), __fdiv_rn(w[1], f[5])), __fdiv_rn(w[2], f[8])));
    return (zp > near_) && (zp < far_);
}

__device__ __forceinline__ uint32_t ordered_bits(float z)
{
    const uint32_t b = __float_as_uint(z);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// smallest i in [lo,hi] with pred(i) true (pred monotone false->true), hi+1 if none
template <class P> __device__ __forceinline__ int first_true(int lo, int hi, P pred)
{
    int l = lo, h = hi + 1;
    while (l < h) {
        const int m = (l + h) >> 1;
        if (pred(m)) h = m; else l = m + 1;
    }
    return l;
}

struct Edges {
    float ya[3], dx[3], xa[3], dy[3];
    __device__ __forceinline__ float A(int i, float yp) const { return __fmul_rn(__fsub_rn(yp, ya[i]), dx[i]); }
    __device__ __forceinline__ float B(int i, float xp) const { return __fmul_rn(__fsub_rn(xp, xa[i]), dy[i]); }
};

struct Window {
    int row_lo, row_hi, col_lo, col_hi;
    int tier;   // 0: bounded window (regular / needle), 1: exact monotone search over the band, 2: wild (verbatim tests), -1: skip
};

__device__ __forceinline__ void make_edges(const float f[9], Edges &e)
{
    e.ya[0] = f[1]; e.dx[0] = __fsub_rn(f[3], f[0]); e.xa[0] = f[0]; e.dy[0] = __fsub_rn(f[4], f[1]);
    e.ya[1] = f[4]; e.dx[1] = __fsub_rn(f[6], f[3]); e.xa[1] = f[3]; e.dy[1] = __fsub_rn(f[7], f[4]);
    e.ya[2] = f[7]; e.dx[2] = __fsub_rn(f[0], f[6]); e.xa[2] = f[6]; e.dy[2] = __fsub_rn(f[1], f[7]);
}

// Search window of a front-facing face inside the band [r0, r1].  A pixel that passes the three float edge
// tests lies within ~3e-5 NDC of each (computed) edge half-plane, hence within  3e-5 / sin(theta_min / 2)  of the
// vertex bounding box (theta_min = smallest corner angle; derivation in DESIGN.md 3.2).  The margin is computed
// per face from sin(theta_min) (0.8 px for well-shaped faces, growing for needles); faces thinner than ~0.03 deg
// or larger than 64 NDC search the whole band with the exact monotone predicate, non-finite or huge
// coordinates run the reference tests verbatim on every band pixel.
__device__ __forceinline__ Window classify(const float f[9], const Edges &e, const float *centre, int is, float isf, int r0, int r1)
{
    Window w;
    w.row_lo = r0; w.row_hi = r1; w.col_lo = 0; w.col_hi = is - 1; w.tier = 2;
    bool wild = false;
#pragma unroll
    for (int k = 0; k < 9; ++k)
        if (k % 3 != 2) wild |= !(fabsf(f[k]) <= kWild);  // catches NaN / inf / huge
    if (wild) return w;
    const float l0 = e.dx[0] * e.dx[0] + e.dy[0] * e.dy[0];
    const float l1 = e.dx[1] * e.dx[1] + e.dy[1] * e.dy[1];
    const float l2 = e.dx[2] * e.dx[2] + e.dy[2] * e.dy[2];
    const float lmax = fmaxf(l0, fmaxf(l1, l2));
    const float lmid = fmaxf(fminf(l0, l1), fminf(fmaxf(l0, l1), l2));
    const float cross = e.dx[0] * e.dy[1] - e.dy[0] * e.dx[1];      // twice the signed area
    const float sin2 = cross * cross;                                 // = sin^2(theta_min) * lmax * lmid
    const float ll = lmax * lmid;
    const float big = fmaxf(fmaxf(fabsf(f[0]), fabsf(f[3])), fmaxf(fmaxf(fabsf(f[6]), fabsf(f[1])), fmaxf(fabsf(f[4]), fabsf(f[7]))));
    // margin (pixels) = 0.5 + 2 x the bound  3e-5 * is / sin(theta_min)  on how far the float pass region can
    // extend beyond the vertex bounding box; faces thinner than ~0.03 deg fall through to the
    // exact whole-band search
    float margin = -1.f;
    if (big <= 64.f && ll > 1e-30f && cross > 0.f) {
        const float m = 0.5f + 6e-5f * isf * sqrtf(ll / sin2);
        if (m <= 0.5f + 0.06f * isf) margin = m;
    }
    if (margin > 0.f) {
        const float xmin = fminf(f[0], fminf(f[3], f[6])), xmax = fmaxf(f[0], fmaxf(f[3], f[6]));
        const float ymin = fminf(f[1], fminf(f[4], f[7])), ymax = fmaxf(f[1], fmaxf(f[4], f[7]));
        // pixel coordinate of an NDC position: 0.5 * (v * is + is - 1)
        w.col_lo = max(0, (int)floorf(0.5f * (xmin * isf + isf - 1.f) - margin));
        w.col_hi = min(is - 1, (int)ceilf(0.5f * (xmax * isf + isf - 1.f) + margin));
        w.row_lo = max(r0, (int)floorf(0.5f * (ymin * isf + isf - 1.f) - margin));
        w.row_hi = min(r1, (int)ceilf(0.5f * (ymax * isf + isf - 1.f) + margin));
        w.tier = (w.row_lo > w.row_hi || w.col_lo > w.col_hi) ? -1 : 0;
        return w;
    }
    // exact band test from the monotone predicate alone (degenerate faces)
    w.tier = 1;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float bmin = fminf(e.B(i, centre[0]), e.B(i, centre[is - 1]));
        if (!(fmaxf(e.A(i, centre[r0]), e.A(i, centre[r1])) >= bmin)) w.tier = -1;
    }
    return w;
}

// Scatter one face into the band's key buffer.  Inside the window the passing pixels are decided by the
// reference's own float predicate (short windows: evaluated directly; wide ones: per-row binary search on it).
__device__ __forceinline__ void scatter_face(const float f[9], const Edges &e, Window w, int fn, const float *centre,
                                             unsigned long long *keys, int is, float isf, int r0, int r1, float near_, float far_)
{
    if (w.tier == 1) {   // exact row range by binary search on the monotone row predicate
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float bmin = fminf(e.B(i, centre[0]), e.B(i, centre[is - 1]));
            if (e.dx[i] >= 0.f) w.row_lo = max(w.row_lo, first_true(r0, r1, [&](int y) { return e.A(i, centre[y]) >= bmin; }));
            else                w.row_hi = min(w.row_hi, first_true(r0, r1, [&](int y) { return !(e.A(i, centre[y]) >= bmin); }) - 1);
        }
        if (w.row_lo > w.row_hi) return;
    }
    float fi[9];
    face_inverse(f, isf, fi);
    const bool direct = w.tier == 2 || (w.col_hi - w.col_lo) < 8;
    for (int y = w.row_lo; y <= w.row_hi; ++y) {
        const float yp = centre[y];
        const float a0 = e.A(0, yp), a1 = e.A(1, yp), a2 = e.A(2, yp);
        int c_lo = w.col_lo, c_hi = w.col_hi;
        if (!direct) {
            const float a[3] = {a0, a1, a2};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (e.dy[i] >= 0.f)  // B non-decreasing in x: pass set {B <= a} is a prefix
                    c_hi = min(c_hi, first_true(w.col_lo, w.col_hi, [&](int x) { return a[i] < e.B(i, centre[x]); }) - 1);
                else                 // B non-increasing: pass set is a suffix
                    c_lo = max(c_lo, first_true(w.col_lo, w.col_hi, [&](int x) { return !(a[i] < e.B(i, centre[x])); }));
            }
        }
        for (int x = c_lo; x <= c_hi; ++x) {
            if (direct) {  // rasterize_cuda_kernel.cu:132-135 verbatim (NaN compares false => passes)
                const float xp = centre[x];
                if ((a0 < e.B(0, xp)) || (a1 < e.B(1, xp)) || (a2 < e.B(2, xp))) continue;
            }
            float wgt[3], zp;
            if (!shade(f, fi, (float)x, (float)y, near_, far_, wgt, zp)) continue;
            const unsigned long long key = ((unsigned long long)ordered_bits(zp) << 32) | (uint32_t)fn;
            atomicMin(&keys[(size_t)(y - r0) * is + x], key);
        }
    }
}

// Warp-cooperative scatter of one "heavy" face (needle / degenerate / wild: large search window): the window's
// pixels are spread over the 32 lanes and each lane runs the reference's edge tests verbatim on its pixels.
__device__ __forceinline__ void scatter_face_warp(const float f[9], const Edges &e, Window w, int fn, const float *centre,
                                                  unsigned long long *keys, int is, float isf, int r0, int r1, float near_, float far_,
                                                  int lane)
{
    if (w.tier == 1) {   // exact row range by binary search on the monotone row predicate (uniform across lanes)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float bmin = fminf(e.B(i, centre[0]), e.B(i, centre[is - 1]));
            if (e.dx[i] >= 0.f) w.row_lo = max(w.row_lo, first_true(r0, r1, [&](int y) { return e.A(i, centre[y]) >= bmin; }));
            else                w.row_hi = min(w.row_hi, first_true(r0, r1, [&](int y) { return !(e.A(i, centre[y]) >= bmin); }) - 1);
        }
        if (w.row_lo > w.row_hi) return;
    }
    float fi[9];
    face_inverse(f, isf, fi);
    const int cols = w.col_hi - w.col_lo + 1;
    const int npx = (w.row_hi - w.row_lo + 1) * cols;
    for (int q = lane; q < npx; q += 32) {
        const int y = w.row_lo + q / cols, x = w.col_lo + q % cols;
        const float yp = centre[y], xp = centre[x];
        // rasterize_cuda_kernel.cu:132-135 verbatim (NaN compares false => passes)
        if ((e.A(0, yp) < e.B(0, xp)) || (e.A(1, yp) < e.B(1, xp)) || (e.A(2, yp) < e.B(2, xp))) continue;
        float wgt[3], zp;
        if (!shade(f, fi, (float)x, (float)y, near_, far_, wgt, zp)) continue;
        const unsigned long long key = ((unsigned long long)ordered_bits(zp) << 32) | (uint32_t)fn;
        atomicMin(&keys[(size_t)(y - r0) * is + x], key);
    }
}

constexpr int kListCap = 6144;   // compacted in-band faces per CTA (overflow is processed in place)
constexpr int kHeavyCap = 2048;  // faces with a large search window, scattered one per warp
constexpr int kLightArea = 100;   // window pixels up to which one thread scatters the face by itself

// Pre-pass (used when a workspace is supplied): every face is culled and classified ONCE per mesh and a packed
// (band << 24 | face) entry is appended to the mesh's bin for each band its window touches, so the raster CTAs
// no longer re-visit all F faces per band.  bins: per mesh [count | entries[cap]].
__global__ void __launch_bounds__(256)
rast_bin_kernel(const float *__restrict__ faces, int F, int is, int band_rows, int n_bands, uint32_t *__restrict__ bins, int cap)
{
    __shared__ float centre_s[2048];
    for (int i = threadIdx.x; i < is; i += blockDim.x) centre_s[i] = (float)((2. * i + 1 - is) / is);
    __syncthreads();
    const int mesh = blockIdx.y;
    const int fn = blockIdx.x * blockDim.x + threadIdx.x;
    if (fn >= F) return;
    const float *fp = faces + ((size_t)mesh * F + fn) * 9;
    float f[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = __ldg(fp + k);
    if (back_facing(f)) return;
    Edges e;
    make_edges(f, e);
    const Window w = classify(f, e, centre_s, is, (float)is, 0, is - 1);
    if (w.tier < 0) return;
    uint32_t *bin = bins + (size_t)mesh * (cap + 1);
    const int b0 = w.row_lo / band_rows, b1 = w.row_hi / band_rows;   // tiers 1/2 span every band; the raster CTA re-tests exactly
    for (int b = b0; b <= b1; ++b) {
        const uint32_t slot = atomicAdd(bin, 1u);
        if (slot < (uint32_t)cap) bin[1 + slot] = ((uint32_t)b << 24) | (uint32_t)fn;
    }
}

// One CTA per (mesh, band).  Dynamic smem: keys[band_rows*is] (u64) | centre[is] (f32) | list[kListCap] (i32).
// Phase A culls / classifies every face and compacts the in-band survivors into a shared list so that phase B
// runs with (nearly) full warps; phase C decodes the keys.
__global__ void __launch_bounds__(kRastThreads)
rasterize_kernel(const float *__restrict__ faces, int F, int is, int band_rows, int n_bands, float near_, float far_,
                 int flip_y, int32_t *__restrict__ fim, float *__restrict__ wim, float *__restrict__ depth,
                 const uint32_t *__restrict__ bins, int cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);
    float *centre = reinterpret_cast<float *>(keys + (size_t)band_rows * is);
    int *list = reinterpret_cast<int *>(centre + is);
    int *heavy = list + kListCap;
    __shared__ int list_n, heavy_n;

    const int mesh = blockIdx.x / n_bands;
    const int band = blockIdx.x % n_bands;
    const int r0 = band * band_rows;
    const int r1 = min(is, r0 + band_rows) - 1;  // inclusive
    const int npix = (r1 - r0 + 1) * is;
    const float isf = (float)is;

    for (int i = threadIdx.x; i < npix; i += blockDim.x) keys[i] = kEmptyKey;
    // rasterize_cuda_kernel.cu:113-114: pixel centre in f64, rounded once
    for (int i = threadIdx.x; i < is; i += blockDim.x) centre[i] = (float)((2. * i + 1 - is) / is);
    if (threadIdx.x == 0) { list_n = 0; heavy_n = 0; }
    __syncthreads();

    const float *mf = faces + (size_t)mesh * F * 9;
    // Loops below are written warp-synchronously (no `continue`, __syncwarp at the end of every trip): with
    // independent thread scheduling an early `continue` lets lanes run ahead into their next face and the warp
    // never reconverges (measured: 4 of 32 lanes active per issued instruction).
    // ---- phase A: compact this band's faces -- from the mesh's bin when the pre-pass ran (and did not overflow) ...
    const uint32_t *bin = bins ? bins + (size_t)mesh * (cap + 1) : nullptr;
    const uint32_t n_bin = bin ? bin[0] : 0u;
    const bool use_bin = bin && n_bin <= (uint32_t)cap;
    if (use_bin) {
        for (uint32_t base = 0; base < n_bin; base += blockDim.x) {
            const uint32_t i = base + threadIdx.x;
            const uint32_t ent = i < n_bin ? __ldg(bin + 1 + i) : 0xffffffffu;
            int fn = -1;
            if (i < n_bin && (int)(ent >> 24) == band) {
                fn = (int)(ent & 0xffffffu);
                const int slot = atomicAdd(&list_n, 1);
                if (slot < kListCap) { list[slot] = fn; fn = -1; }
            }
            if (fn >= 0) {   // list full: do it now
                float f[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) f[k] = __ldg(mf + (size_t)fn * 9 + k);
                Edges e;
                make_edges(f, e);
                const Window w = classify(f, e, centre, is, isf, r0, r1);
                if (w.tier >= 0) scatter_face(f, e, w, fn, centre, keys, is, isf, r0, r1, near_, far_);
            }
            __syncwarp();
        }
    } else {
        // ... or by culling / classifying every face of the mesh here
        for (int base = 0; base < F; base += blockDim.x) {
            const int fn = base + threadIdx.x;
            float f[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) f[k] = fn < F ? __ldg(mf + (size_t)fn * 9 + k) : 0.f;
            Window w;
            w.tier = -1;
            Edges e;
            if (fn < F && !back_facing(f)) {
                make_edges(f, e);
                w = classify(f, e, centre, is, isf, r0, r1);
            }
            bool overflow = false;
            if (w.tier >= 0) {
                const int slot = atomicAdd(&list_n, 1);
                if (slot < kListCap) list[slot] = fn;
                else overflow = true;
            }
            if (overflow) scatter_face(f, e, w, fn, centre, keys, is, isf, r0, r1, near_, far_);
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- phase B: scatter the compacted faces
    const int n_list = min(list_n, kListCap);
    for (int base = 0; base < n_list; base += blockDim.x) {
        const int i = base + threadIdx.x;
        if (i < n_list) {
            const int fn = list[i];
            float f[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) f[k] = __ldg(mf + (size_t)fn * 9 + k);
            Edges e;
            make_edges(f, e);
            const Window w = classify(f, e, centre, is, isf, r0, r1);
            if (w.tier >= 0) {
                const bool light = w.tier == 0 && (w.row_hi - w.row_lo + 1) * (w.col_hi - w.col_lo + 1) <= kLightArea;
                int slot = kHeavyCap;
                if (!light) slot = atomicAdd(&heavy_n, 1);
                if (slot < kHeavyCap) heavy[slot] = fn;                 // large window: one warp per face below
                else scatter_face(f, e, w, fn, centre, keys, is, isf, r0, r1, near_, far_);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    {
        const int n_heavy = min(heavy_n, kHeavyCap);
        const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int h = wid; h < n_heavy; h += nw) {
            const int fn = heavy[h];
            float f[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) f[k] = __ldg(mf + (size_t)fn * 9 + k);
            Edges e;
            make_edges(f, e);
            const Window w = classify(f, e, centre, is, isf, r0, r1);
            if (w.tier >= 0) scatter_face_warp(f, e, w, fn, centre, keys, is, isf, r0, r1, near_, far_, lane);
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- phase C: keys -> fim / wim / depth  (rasterize_cuda_kernel.cu:174-179, rasterize.py:50-52,335-338)
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
        const int y = r0 + i / is, x = i % is;
        const unsigned long long key = keys[i];
        const int yo = flip_y ? (is - 1 - y) : y;
        const size_t o = ((size_t)mesh * is + yo) * is + x;
        int face = -1;
        float w[3] = {0.f, 0.f, 0.f};
        float zp = far_;
        if (key != kEmptyKey) {
            face = (int)(uint32_t)key;
            float f[9], fi[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) f[k] = __ldg(mf + (size_t)face * 9 + k);
            face_inverse(f, isf, fi);
            float z2;
            shade(f, fi, (float)x, (float)y, near_, far_, w, z2);
            zp = from_ordered_bits((uint32_t)(key >> 32));
        }
        fim[o] = face;
        wim[3 * o] = w[0]; wim[3 * o + 1] = w[1]; wim[3 * o + 2] = w[2];
        if (depth) depth[o] = zp;
    }
}

__global__ void face_inv_kernel(const float *__restrict__ faces, int64_t BF, int is, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BF) return;
    float f[9], fi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = faces[i * 9 + k];
    if (back_facing(f)) return;
    face_inverse(f, (float)is, fi);
#pragma unroll
    for (int k = 0; k < 9; ++k) out[i * 9 + k] = fi[k];
}

// R0: utils/nmr.py:109-140 + :506 + look_at (identity rotation for eye on -z) + vertices_to_faces.
// One thread per (mesh, face, corner).
__global__ void project_faces_kernel(const float *__restrict__ verts, const float *__restrict__ cam,
                                     const int32_t *__restrict__ fidx, int B, int V, int F, float eye_z,
                                     float *__restrict__ faces)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * F * 3) return;
    const int b = (int)(t / ((int64_t)F * 3));
    const int fc = (int)(t % ((int64_t)F * 3));
    const int v = fidx[fc];
    const float *p = verts + ((size_t)b * V + v) * 3;
    const float *c = cam + (size_t)b * 15;
    // OpenGL -> camera coords (x, -y, -z); einsum accumulations written as the
    // left-to-right sums torch performs for K=3 dot products.
    const float X = p[0], Y = -p[1], Z = -p[2];
    const float px = c[0] * X + c[1] * Y + c[2] * Z;
    const float py = c[3] * X + c[4] * Y + c[5] * Z;
    const float pz = c[6] * X + c[7] * Y + c[8] * Z;
    const float u = px / pz, w = py / pz;
    float ox = c[9] * u + c[10] * w + c[11];
    float oy = c[12] * u + c[13] * w + c[14];
    ox = ox / 255.0f * 2.f - 1.f;
    oy = oy / 255.0f * 2.f - 1.f;
    float *o = faces + (size_t)t * 3;
    o[0] = ox;
    o[1] = -oy;          // utils/nmr.py:506
    o[2] = Z - eye_z;    // look_at: vertices - eye, identity rotation
}

// R4/R5: table gathers by fim (utils/nmr.py:567-595) + one-hot seg + hand mask (trainer.py:71-72).
__global__ void condition_maps_kernel(const int32_t *__restrict__ fim, int64_t npix_total, int hw, int F,
                                      const float *__restrict__ map_fn, const float *__restrict__ sem_full,
                                      int n_hand, float *__restrict__ cond, float *__restrict__ seg,
                                      float *__restrict__ not_hand)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix_total) return;
    const int b = (int)(i / hw), p = (int)(i % hw);
    const int f = fim[i];
    const int row = f < 0 ? F : f;  // python negative index -1 -> last (background) row
    if (cond) {
#pragma unroll
        for (int c = 0; c < 3; ++c) cond[((size_t)b * 3 + c) * hw + p] = map_fn[(size_t)row * 3 + c];
    }
    if (seg) {
        const float s = sem_full[row];
#pragma unroll
        for (int c = 0; c < 15; ++c) seg[((size_t)b * 15 + c) * hw + p] = (s == (float)(c + 1)) ? 1.f : 0.f;
    }
    if (not_hand) not_hand[i] = 1.f - ((f != -1 && f < n_hand) ? 1.f : 0.f);
}

// R7: utils/nmr.py:874-925 (T only) with the y re-negation of trainer.py:67-68.
__global__ void bc_transform_kernel(const float *__restrict__ src_faces, const int32_t *__restrict__ fim,
                                    const float *__restrict__ wim, int64_t npix_total, int hw, int F,
                                    float *__restrict__ T)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix_total) return;
    const int b = (int)(i / hw);
    const int f = fim[i];
    float tx = -2.f, ty = -2.f;
    if (f != -1) {
        const float *fv = src_faces + ((size_t)b * F + f) * 9;
        const float w0 = wim[3 * i], w1 = wim[3 * i + 1], w2 = wim[3 * i + 2];
        // (f2pts * w[:, :, None]).sum(dim=1): products rounded, summed in vertex order
        tx = __fadd_rn(__fadd_rn(__fmul_rn(fv[0], w0), __fmul_rn(fv[3], w1)), __fmul_rn(fv[6], w2));
        ty = __fadd_rn(__fadd_rn(__fmul_rn(-fv[1], w0), __fmul_rn(-fv[4], w1)), __fmul_rn(-fv[7], w2));
    }
    T[2 * i] = tx;
    T[2 * i + 1] = ty;
}

// R6: utils/util.py:142-153.  Inputs are 0/1 masks; erode == all ks*ks neighbours (pad value 1) are 1.
// The reference tests conv_sum == ks*ks on floats, reproduced as an exact float sum of the window.
__global__ void erode_kernel(const float *__restrict__ in, float *__restrict__ out, int B, int H, int W, int ks)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * H * W) return;
    const int b = (int)(i / ((int64_t)H * W));
    const int y = (int)((i / W) % H), x = (int)(i % W);
    const int r = ks / 2;
    float s = 0.f;
    for (int dy = -r; dy <= r; ++dy)
        for (int dx = -r; dx <= r; ++dx) {
            const int yy = y + dy, xx = x + dx;
            s += (yy < 0 || yy >= H || xx < 0 || xx >= W) ? 1.f : in[((size_t)b * H + yy) * W + xx];
        }
    out[i] = (s == (float)(ks * ks)) ? 1.f : 0.f;
}

}  // namespace

// ------------------------------------------------------------------ stage R8: UV-texture warp
// utils/nmr.py:973-1040 (get_texture_backward_warp, first half), one thread per (sample, atlas pixel).
__global__ void uv_backward_warp_kernel(const float *__restrict__ src_faces, const int32_t *__restrict__ fim_uv,
                                        const float *__restrict__ wim_uv, const int32_t *__restrict__ src_fim, int64_t n, int hw_uv,
                                        int F, int is, float *__restrict__ T, float *__restrict__ O)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = (int)(i / hw_uv), p = (int)(i % hw_uv);
    const int f = fim_uv[p];
    float tx = -2.f, ty = -2.f, o = 0.f;
    if (f != -1) {
        const float *fv = src_faces + ((size_t)b * F + f) * 9;
        const float w0 = wim_uv[3 * p], w1 = wim_uv[3 * p + 1], w2 = wim_uv[3 * p + 2];
        // (f2pts * w[:, :, None]).sum(dim=1): products rounded, summed in vertex order; y negated as trainer.py:67-68
        tx = __fadd_rn(__fadd_rn(__fmul_rn(fv[0], w0), __fmul_rn(fv[3], w1)), __fmul_rn(fv[6], w2));
        ty = __fadd_rn(__fadd_rn(__fmul_rn(-fv[1], w0), __fmul_rn(-fv[4], w1)), __fmul_rn(-fv[7], w2));
        // ((T + 1) / 2.0 * 255.0).long().clamp(0, 255): truncation toward zero, the literal 255 of nmr.py:1014
        const float lim = (float)(is - 1);
        const int cx = min(max((int)__fmul_rn(__fdiv_rn(__fadd_rn(tx, 1.f), 2.f), lim), 0), is - 1);
        const int cy = min(max((int)__fmul_rn(__fdiv_rn(__fadd_rn(ty, 1.f), 2.f), lim), 0), is - 1);
        const int32_t *sf = src_fim + (size_t)b * is * is;
        bool vis = false;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int x = min(max(cx + dx, 0), is - 1), y = min(max(cy + dy, 0), is - 1);
                vis |= sf[y * is + x] == f;
            }
        o = vis ? 0.f : 1.f;
    }
    T[2 * i] = tx;
    T[2 * i + 1] = ty;
    O[i] = o;
}

// utils/nmr.py:1068-1100 sample_from_texture_dense: uv_coord (F,3,2) shared by the batch
__global__ void sample_texture_dense_kernel(const float *__restrict__ uv_coord, const int32_t *__restrict__ fim,
                                            const float *__restrict__ wim, int64_t n, float *__restrict__ T)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int f = fim[i];
    float tx = -2.f, ty = -2.f;
    if (f != -1) {
        const float *uv = uv_coord + (size_t)f * 6;
        const float w0 = wim[3 * i], w1 = wim[3 * i + 1], w2 = wim[3 * i + 2];
        tx = __fadd_rn(__fadd_rn(__fmul_rn(uv[0], w0), __fmul_rn(uv[2], w1)), __fmul_rn(uv[4], w2));
        ty = __fadd_rn(__fadd_rn(__fmul_rn(uv[1], w0), __fmul_rn(uv[3], w1)), __fmul_rn(uv[5], w2));
    }
    T[2 * i] = tx;
    T[2 * i + 1] = ty;
}

// F.grid_sample(im, grid, 'bilinear', 'zeros', align_corners) on NCHW f32 (aten GridSampler.cuh semantics: unnormalise,
// floor, the four corner weights as products of the distances to the opposite corner, out-of-range corners contribute 0)
__global__ void grid_sample_nchw_kernel(const float *__restrict__ im, int B, int C, int Hi, int Wi, const float *__restrict__ grid,
                                        int Ho, int Wo, int align, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Ho * Wo) return;
    const int b = (int)(i / ((int64_t)Ho * Wo));
    const int64_t p = i % ((int64_t)Ho * Wo);
    const float gx = grid[2 * i], gy = grid[2 * i + 1];
    const float ix = align ? (gx + 1.f) / 2.f * (float)(Wi - 1) : ((gx + 1.f) * (float)Wi - 1.f) / 2.f;
    const float iy = align ? (gy + 1.f) / 2.f * (float)(Hi - 1) : ((gy + 1.f) * (float)Hi - 1.f) / 2.f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float nw = (x1 - ix) * (y1 - iy), ne = (ix - x0) * (y1 - iy), sw = (x1 - ix) * (iy - y0), se = (ix - x0) * (iy - y0);
    const bool vx0 = x0 >= 0 && x0 < Wi, vx1 = x1 >= 0 && x1 < Wi, vy0 = y0 >= 0 && y0 < Hi, vy1 = y1 >= 0 && y1 < Hi;
    for (int c = 0; c < C; ++c) {
        const float *pl = im + ((size_t)b * C + c) * Hi * Wi;
        float acc = 0.f;
        if (vy0 && vx0) acc += pl[(size_t)y0 * Wi + x0] * nw;
        if (vy0 && vx1) acc += pl[(size_t)y0 * Wi + x1] * ne;
        if (vy1 && vx0) acc += pl[(size_t)y1 * Wi + x0] * sw;
        if (vy1 && vx1) acc += pl[(size_t)y1 * Wi + x1] * se;
        out[((size_t)b * C + c) * Ho * Wo + p] = acc;
    }
}

// utils/nmr.py:1049-1056: O <- 1 - erode3(1 - erode3(O)) (util.morph, pad value 1), syn = syn*(1-O) + O, then the stock object
// texture over columns >= x0
__global__ void uv_texture_compose_kernel(float *__restrict__ syn, const float *__restrict__ O, const float *__restrict__ preload,
                                          int B, int C, int Hu, int Wu, int x0)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Hu * Wu) return;
    const int b = (int)(i / ((int64_t)Hu * Wu));
    const int y = (int)((i / Wu) % Hu), x = (int)(i % Wu);
    const float *Ob = O + (size_t)b * Hu * Wu;
    float open = 0.f;   // dilate3(erode3(O)): any in-image neighbour q whose whole (pad-1) 3x3 window is 1
    for (int dy = -1; dy <= 1 && open == 0.f; ++dy)
        for (int dx = -1; dx <= 1 && open == 0.f; ++dx) {
            const int qy = y + dy, qx = x + dx;
            if (qy < 0 || qy >= Hu || qx < 0 || qx >= Wu) continue;
            float s = 0.f;
            for (int ey = -1; ey <= 1; ++ey)
                for (int ex = -1; ex <= 1; ++ex) {
                    const int ry = qy + ey, rx = qx + ex;
                    s += (ry < 0 || ry >= Hu || rx < 0 || rx >= Wu) ? 1.f : Ob[(size_t)ry * Wu + rx];
                }
            if (s == 9.f) open = 1.f;
        }
    for (int c = 0; c < C; ++c) {
        float *v = syn + (((size_t)b * C + c) * Hu + y) * Wu + x;
        if (preload && x >= x0) *v = preload[((size_t)y * (Wu - x0) + (x - x0)) * C + c];
        else *v = *v * (1.f - open) + 1.0f * open;
    }
}


// ------------------------------------------------------------------ row N1: HandRecoveryFlow tail in one pass
// models/trainer.py:66-145 for the whole batch: everything the generator consumes, straight from the two face-index maps.
// One thread per (sample, pixel).  The erosions (utils/util.py:142-153: pad value 1, ks x ks window sum == ks^2) are evaluated
// from the face-index map itself, so no intermediate mask tensor is written and read back.
struct CondSide {
    const int32_t *fim;       // (B,is,is)
    const float *render;      // (B,3,is,is) UV re-rendering at this pose (stage R8)
    float *obj_inputs, *obj_conds, *hand_inputs, *hand_conds, *mask_bg, *mask_hand;
};
struct CondArgs {
    CondSide side[2];         // 0 = src, 1 = ref/tsf
    const float *src_img;     // (B,3,is,is)
    const float *src_faces;   // (B,F,3,3)
    const float *wim_ref;     // (B,is,is,3)
    const float *map_fn, *sem_full;
    float *bg_inputs, *T;
    int B, F, is, n_hand, bg_ks;
};

// value of the erosion inputs at (y, x) of one face-index map; outside the image util.morph pads with 1
__device__ __forceinline__ void mask_values(const CondArgs &a, const int32_t *fim, int y, int x, float &bg, float &not_hand)
{
    if (y < 0 || y >= a.is || x < 0 || x >= a.is) { bg = 1.f; not_hand = 1.f; return; }
    const int f = fim[y * a.is + x];
    bg = a.map_fn[(size_t)(f < 0 ? a.F : f) * 3 + 2];
    not_hand = 1.f - ((f != -1 && f < a.n_hand) ? 1.f : 0.f);
}

constexpr int CT_W = 32, CT_H = 8, CT_R = 7;       // tile and the largest supported erosion radius (ks <= 15)

// 2-D tiles: the erosion inputs of a tile and its halo are gathered once into shared memory (face index -> table lookups are
// the expensive part), the 15 x 15 erosion is evaluated separably (row sums, then column sums).
__global__ void __launch_bounds__(CT_W * CT_H) condition_inputs_kernel(const CondArgs a)
{
    __shared__ float s_bg[2][CT_H + 2 * CT_R][CT_W + 2 * CT_R];    // [side]: background channel of cond (side 1 only needs halo 1)
    __shared__ float s_nh[2][CT_H + 2][CT_W + 2];                  // not_hand, halo 1
    __shared__ float s_row[CT_H + 2 * CT_R][CT_W];                 // horizontal sums of the source background window
    const int hw = a.is * a.is;
    const int b = blockIdx.z, x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * CT_W + tx;
    const int r = a.bg_ks / 2;
    const int32_t *fim_s = a.side[0].fim + (size_t)b * hw, *fim_r = a.side[1].fim + (size_t)b * hw;
    for (int i = tid; i < (CT_H + 2 * CT_R) * (CT_W + 2 * CT_R); i += CT_W * CT_H) {
        const int yy = i / (CT_W + 2 * CT_R), xx = i % (CT_W + 2 * CT_R);
        float bg, nh;
        mask_values(a, fim_s, y0 + yy - CT_R, x0 + xx - CT_R, bg, nh);
        s_bg[0][yy][xx] = bg;
        const int hy = yy - (CT_R - 1), hx = xx - (CT_R - 1);                       // position inside the halo-1 tiles
        if (hy >= 0 && hy < CT_H + 2 && hx >= 0 && hx < CT_W + 2) {
            s_nh[0][hy][hx] = nh;
            mask_values(a, fim_r, y0 + yy - CT_R, x0 + xx - CT_R, bg, nh);
            s_bg[1][yy][xx] = bg;
            s_nh[1][hy][hx] = nh;
        }
    }
    __syncthreads();
    for (int i = tid; i < (CT_H + 2 * CT_R) * CT_W; i += CT_W * CT_H) {              // row sums over [x - r, x + r]
        const int yy = i / CT_W, xx = i % CT_W;
        float sacc = 0.f;
        for (int dx = -r; dx <= r; ++dx) sacc += s_bg[0][yy][xx + CT_R + dx];
        s_row[yy][xx] = sacc;
    }
    __syncthreads();
    const int x = x0 + tx, y = y0 + ty;
    if (x >= a.is || y >= a.is) return;
    const int p = y * a.is + x;
    const int64_t i = (int64_t)b * hw + p;
    float bgm;
    {
        float sacc = 0.f;
        for (int dy = -r; dy <= r; ++dy) sacc += s_row[ty + CT_R + dy][tx];
        bgm = sacc == (float)(a.bg_ks * a.bg_ks) ? 1.f : 0.f;
    }
    float m_hand_ref = 0.f;
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
        const CondSide &S = a.side[sd];
        const int f = (sd == 0 ? fim_s : fim_r)[p];
        const int row = f < 0 ? a.F : f;
        const float c0 = a.map_fn[(size_t)row * 3], c1 = a.map_fn[(size_t)row * 3 + 1], c2 = a.map_fn[(size_t)row * 3 + 2];
        const float sem = a.sem_full[row];
        // erode3 of not_hand and of the background channel of cond (trainer.py:72,109-110)
        float s_hand = 0.f, s_bgsum = 0.f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                s_hand += s_nh[sd][ty + 1 + dy][tx + 1 + dx];
                s_bgsum += s_bg[sd][ty + CT_R + dy][tx + CT_R + dx];
            }
        const float m_hand = s_hand == 9.f ? 1.f : 0.f, m_bg = s_bgsum == 9.f ? 1.f : 0.f;
        if (sd == 1) m_hand_ref = m_hand;
        const float hm = c0 < 1.5f ? 1.f : 0.f, om = c0 > 1.5f ? 1.f : 0.f;           // trainer.py:112-124
        const size_t p3 = (size_t)b * 3 * hw + p, p12 = (size_t)b * 12 * hw + p;
        S.hand_conds[p3] = hm * c0; S.hand_conds[p3 + hw] = hm * c1; S.hand_conds[p3 + 2 * (size_t)hw] = c2 + 1.f - hm;
        S.obj_conds[p12] = om * c0; S.obj_conds[p12 + hw] = om * c1; S.obj_conds[p12 + 2 * (size_t)hw] = c2 + 1.f - om;
#pragma unroll
        for (int c = 0; c < 9; ++c) S.obj_conds[p12 + (size_t)(3 + c) * hw] = (sem == (float)(c + 7)) ? 1.f : 0.f;   // seg[:, 6:]
        const float *hand_rgb = sd == 0 ? a.src_img : S.render;                       // trainer.py:128 vs 132
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            S.obj_inputs[p3 + (size_t)c * hw] = S.render[p3 + (size_t)c * hw] * (m_hand - m_bg);        // trainer.py:127,131
            S.hand_inputs[p3 + (size_t)c * hw] = hand_rgb[p3 + (size_t)c * hw] * (1.f - m_hand);
        }
        S.mask_bg[i] = m_bg;
        S.mask_hand[i] = m_hand;
    }
    // bg_inputs = [src_img * erode_ks(cond_src[:, -1:]), erode_ks(...)]   (trainer.py:135-136, ks = 15)
    {
        const size_t p3 = (size_t)b * 3 * hw + p, p4 = (size_t)b * 4 * hw + p;
#pragma unroll
        for (int c = 0; c < 3; ++c) a.bg_inputs[p4 + (size_t)c * hw] = a.src_img[p3 + (size_t)c * hw] * bgm;
        a.bg_inputs[p4 + 3 * (size_t)hw] = bgm;
    }
    // T = cal_bc_transform (nmr.py:874-925), masked to the hand region of the target pose (trainer.py:80-81)
    {
        const int f = fim_r[p];
        float tx_ = -2.f, ty_ = -2.f;
        if (f != -1 && m_hand_ref != 1.f) {
            const float *fv = a.src_faces + ((size_t)b * a.F + f) * 9;
            const float w0 = a.wim_ref[3 * i], w1 = a.wim_ref[3 * i + 1], w2 = a.wim_ref[3 * i + 2];
            tx_ = __fadd_rn(__fadd_rn(__fmul_rn(fv[0], w0), __fmul_rn(fv[3], w1)), __fmul_rn(fv[6], w2));
            ty_ = __fadd_rn(__fadd_rn(__fmul_rn(-fv[1], w0), __fmul_rn(-fv[4], w1)), __fmul_rn(-fv[7], w2));
        }
        a.T[2 * i] = tx_;
        a.T[2 * i + 1] = ty_;
    }
}

}  // namespace hoig

using namespace hoig;

static int g_band_pixels = 8192;   // 32 rows at 256 px: two CTAs per SM measured best (profiles/)
// Tuning hook: pixels per band (<= 16384, the 128 KB key buffer); smaller bands -> more CTAs per SM, more face re-visits.
extern "C" void hoig_set_rasterizer_band_pixels(int n) { g_band_pixels = n < 256 ? 256 : (n > kMaxBandPixels ? kMaxBandPixels : n); }

static inline int bin_cap(int F) { return 2 * F + 1024; }
// Optional workspace for the binning pre-pass: per mesh a counter and 2F+1024 packed (band, face) entries.
// Without it (NULL / too small) every band CTA scans all faces itself.
extern "C" size_t hoig_rasterize_workspace_bytes(int B, int F, int) { return (size_t)B * (bin_cap(F) + 1) * sizeof(uint32_t); }  // z-buffer lives in shared memory

extern "C" int hoig_rasterize_fim_wim(const float *faces, int B, int F, int image_size, float near_, float far_,
                                      int flip_y, int32_t *fim, float *wim, float *depth, void *workspace, size_t workspace_bytes,
                                      hoigStream_t stream)
{
    HOIG_REQUIRE(B >= 0 && F >= 0 && image_size >= 1 && image_size <= 2048, "rasterize: bad shape B=%d F=%d is=%d", B, F, image_size);
    if (B == 0) return HOIG_OK;
    HOIG_REQUIRE((faces || F == 0) && fim && wim, "rasterize: null pointer");
    const int is = image_size;
    int band_rows = g_band_pixels / is;
    if (band_rows > is) band_rows = is;
    HOIG_REQUIRE(band_rows >= 1, "rasterize: image too wide");
    const int n_bands = ceil_div(is, band_rows);
    const size_t smem = (size_t)band_rows * is * sizeof(unsigned long long) + (size_t)is * sizeof(float) + (size_t)(kListCap + kHeavyCap) * sizeof(int);
    if (first_use_on_device(SLOT_RASTERIZE) &&
        cudaFuncSetAttribute(rasterize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return check_launch("rasterize smem attribute");
    uint32_t *bins = nullptr;
    const int cap = bin_cap(F);
    if (workspace && workspace_bytes >= hoig_rasterize_workspace_bytes(B, F, is) && F > 0 && F < (1 << 24) && n_bands <= 256 && B <= 65535) {
        bins = static_cast<uint32_t *>(workspace);
        // zero the per-mesh counters (strided): one 2-D memset
        if (cudaMemset2DAsync(bins, (size_t)(cap + 1) * sizeof(uint32_t), 0, sizeof(uint32_t), (size_t)B, as_stream(stream)) != cudaSuccess)
            return check_launch("rasterize bin memset");
        rast_bin_kernel<<<dim3((unsigned)ceil_div(F, 256), (unsigned)B), 256, 0, as_stream(stream)>>>(faces, F, is, band_rows, n_bands, bins, cap);
        const int rc = check_launch("rast_bin_kernel");
        if (rc != HOIG_OK) return rc;
    }
    rasterize_kernel<<<(unsigned)((int64_t)B * n_bands), kRastThreads, smem, as_stream(stream)>>>(
        faces, F, is, band_rows, n_bands, near_, far_, flip_y, fim, wim, depth, bins, cap);
    return check_launch("rasterize_kernel");
}

extern "C" int hoig_face_inv(const float *faces, int64_t BF, int image_size, float *faces_inv, hoigStream_t stream)
{
    HOIG_REQUIRE(faces && faces_inv && BF >= 0, "face_inv: bad argument");
    if (BF == 0) return HOIG_OK;
    face_inv_kernel<<<ceil_div(BF, 256), 256, 0, as_stream(stream)>>>(faces, BF, image_size, faces_inv);
    return check_launch("face_inv_kernel");
}

extern "C" int hoig_project_faces(const float *verts, const float *cam, const int32_t *faces_idx, int B, int V, int F,
                                  float eye_z, float *faces, hoigStream_t stream)
{
    HOIG_REQUIRE(verts && cam && faces_idx && faces, "project_faces: null pointer");
    const int64_t n = (int64_t)B * F * 3;
    if (n == 0) return HOIG_OK;
    project_faces_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(verts, cam, faces_idx, B, V, F, eye_z, faces);
    return check_launch("project_faces_kernel");
}

extern "C" int hoig_condition_maps(const int32_t *fim, int B, int F, int image_size, const float *map_fn,
                                   const float *sem_full, int n_hand_faces, float *cond, float *seg, float *not_hand,
                                   hoigStream_t stream)
{
    HOIG_REQUIRE(fim && map_fn && sem_full, "condition_maps: null pointer");
    const int hw = image_size * image_size;
    const int64_t n = (int64_t)B * hw;
    if (n == 0) return HOIG_OK;
    condition_maps_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(fim, n, hw, F, map_fn, sem_full, n_hand_faces,
                                                                        cond, seg, not_hand);
    return check_launch("condition_maps_kernel");
}

extern "C" int hoig_bc_transform(const float *src_faces, const int32_t *fim_ref, const float *wim_ref, int B, int F,
                                 int image_size, float *T, hoigStream_t stream)
{
    HOIG_REQUIRE(src_faces && fim_ref && wim_ref && T, "bc_transform: null pointer");
    const int hw = image_size * image_size;
    const int64_t n = (int64_t)B * hw;
    if (n == 0) return HOIG_OK;
    bc_transform_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(src_faces, fim_ref, wim_ref, n, hw, F, T);
    return check_launch("bc_transform_kernel");
}

extern "C" int hoig_erode(const float *in, float *out, int B, int H, int W, int ks, hoigStream_t stream)
{
    HOIG_REQUIRE(in && out && ks >= 1 && (ks & 1), "erode: bad argument");
    const int64_t n = (int64_t)B * H * W;
    if (n == 0) return HOIG_OK;
    erode_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(in, out, B, H, W, ks);
    return check_launch("erode_kernel");
}

extern "C" int hoig_uv_backward_warp(const float *src_faces, const int32_t *fim_uv, const float *wim_uv, const int32_t *src_fim,
                                     int B, int F, int Hu, int Wu, int image_size, float *T, float *O, hoigStream_t stream)
{
    HOIG_REQUIRE(src_faces && fim_uv && wim_uv && src_fim && T && O && image_size > 0, "uv_backward_warp: bad argument");
    const int64_t n = (int64_t)B * Hu * Wu;
    if (n == 0) return HOIG_OK;
    uv_backward_warp_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(src_faces, fim_uv, wim_uv, src_fim, n, Hu * Wu, F, image_size, T, O);
    return check_launch("uv_backward_warp_kernel");
}

extern "C" int hoig_sample_texture_dense(const float *uv_coord, const int32_t *fim, const float *wim, int B, int H, int W, float *T,
                                         hoigStream_t stream)
{
    HOIG_REQUIRE(uv_coord && fim && wim && T, "sample_texture_dense: null pointer");
    const int64_t n = (int64_t)B * H * W;
    if (n == 0) return HOIG_OK;
    sample_texture_dense_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(uv_coord, fim, wim, n, T);
    return check_launch("sample_texture_dense_kernel");
}

extern "C" int hoig_grid_sample_nchw(const float *im, int B, int C, int Hi, int Wi, const float *grid, int Ho, int Wo, int align_corners,
                                     float *out, hoigStream_t stream)
{
    HOIG_REQUIRE(im && grid && out && C > 0 && Hi > 0 && Wi > 0, "grid_sample_nchw: bad argument");
    const int64_t n = (int64_t)B * Ho * Wo;
    if (n == 0) return HOIG_OK;
    grid_sample_nchw_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(im, B, C, Hi, Wi, grid, Ho, Wo, align_corners, out);
    return check_launch("grid_sample_nchw_kernel");
}

extern "C" int hoig_uv_texture_compose(float *syn, const float *O, const float *preload, int B, int C, int Hu, int Wu, int x0,
                                       hoigStream_t stream)
{
    HOIG_REQUIRE(syn && O && x0 >= 0 && x0 <= Wu, "uv_texture_compose: bad argument");
    const int64_t n = (int64_t)B * Hu * Wu;
    if (n == 0) return HOIG_OK;
    uv_texture_compose_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(syn, O, preload, B, C, Hu, Wu, x0);
    return check_launch("uv_texture_compose_kernel");
}

extern "C" int hoig_condition_inputs(const hoigCondInputsDesc *d, hoigStream_t stream)
{
    HOIG_REQUIRE(d, "condition_inputs: null descriptor");
    HOIG_REQUIRE(d->fim_src && d->fim_ref && d->wim_ref && d->src_faces && d->src_img && d->render_src && d->render_ref && d->map_fn &&
                     d->sem_full, "condition_inputs: null input");
    HOIG_REQUIRE(d->bg_inputs && d->src_obj_inputs && d->src_obj_conds && d->src_hand_inputs && d->src_hand_conds && d->tsf_obj_inputs &&
                     d->tsf_obj_conds && d->tsf_hand_inputs && d->tsf_hand_conds && d->T && d->src_mask_bg && d->ref_mask_bg &&
                     d->src_mask_hand && d->ref_mask_hand, "condition_inputs: null output");
    HOIG_REQUIRE(d->B >= 0 && d->F > 0 && d->image_size > 0 && d->bg_erode_ks >= 1 && (d->bg_erode_ks & 1), "condition_inputs: bad shape");
    CondArgs a;
    a.side[0] = {d->fim_src, d->render_src, d->src_obj_inputs, d->src_obj_conds, d->src_hand_inputs, d->src_hand_conds, d->src_mask_bg, d->src_mask_hand};
    a.side[1] = {d->fim_ref, d->render_ref, d->tsf_obj_inputs, d->tsf_obj_conds, d->tsf_hand_inputs, d->tsf_hand_conds, d->ref_mask_bg, d->ref_mask_hand};
    a.src_img = d->src_img; a.src_faces = d->src_faces; a.wim_ref = d->wim_ref; a.map_fn = d->map_fn; a.sem_full = d->sem_full;
    a.bg_inputs = d->bg_inputs; a.T = d->T;
    a.B = d->B; a.F = d->F; a.is = d->image_size; a.n_hand = d->n_hand_faces; a.bg_ks = d->bg_erode_ks;
    HOIG_REQUIRE(d->bg_erode_ks <= 2 * CT_R + 1, "condition_inputs: erosion size %d not supported (<= %d)", d->bg_erode_ks, 2 * CT_R + 1);
    if (d->B == 0) return HOIG_OK;
    const dim3 grid((unsigned)ceil_div(d->image_size, CT_W), (unsigned)ceil_div(d->image_size, CT_H), (unsigned)d->B);
    condition_inputs_kernel<<<grid, dim3(CT_W, CT_H), 0, as_stream(stream)>>>(a);
    return check_launch("condition_inputs_kernel");
}
