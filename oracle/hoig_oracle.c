/*
 * hoig_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity oracle).
 *
 * Plain-C CPU restatement of the three native ops on HOGAN's generator hot
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library; the product (hoig_b200/) never
 * does.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/HOIG_HOv3/thirdparty).  Float semantics follow the op
 * sequence the reference's kernels compile to with nvcc's default
 * -fmad=true: contraction sites are written as
 * explicit fmaf(), everything else is a separately rounded IEEE op, so this
 * file must be built with -ffp-contract=off and without -ffast-math.
 *
 * Parity pinning: the reference ships no runnable golden vectors for these
 * forwards (its PNG fixtures are absent).  The restatement is pinned instead
 * by (a) the look_at / pixel-shuffle KATs the reference tests imply
 * (tests/test_oracle_kats.py) and (b) on a GPU box, bit-comparison with the
 * reference's own kernels compiled by oracle/build_ref.py (oracle/_ref).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

/* Tiny pthread parallel-for (the image has no libgomp): fn(ctx, begin, end)
 * over [0, n) split into contiguous chunks, one per online core. */
typedef void (*range_fn)(void *ctx, long begin, long end);
typedef struct { range_fn fn; void *ctx; long begin, end; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->ctx, j->begin, j->end); return 0; }

int oracle_num_threads(void)
{
    const char *e = getenv("HOIG_ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (n > 256 ? 256 : (int)n);
}

static void parallel_for(long n, range_fn fn, void *ctx)
{
    int nt = oracle_num_threads();
    if (nt > n) nt = n > 0 ? (int)n : 1;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < nt; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].begin = n * t / nt; jobs[t].end = n * (t + 1) / nt;
        if (t + 1 < nt) pthread_create(&th[t], 0, job_main, &jobs[t]);
    }
    job_main(&jobs[nt - 1]);
    for (int t = 0; t + 1 < nt; ++t) pthread_join(th[t], 0);
}

/* neural_renderer/neural_renderer/cuda/rasterize_cuda_kernel.cu:41-84
 * (forward_face_index_map_cuda_kernel_1): per-face inverse of [p;1].
 * faces: (BF,3,3) xyz per vertex.  faces_inv: (BF,9), left untouched (caller
 * zero-fills, rasterize.py:164) for back-facing faces. */
void oracle_face_inv(const float *faces, long BF, int is, float *faces_inv)
{
    const float isf = (float)is;
    for (long i = 0; i < BF; ++i) {
        const float *f = faces + i * 9;
        float *o = faces_inv + i * 9;
        /* :57 back-face cull, plain mul/sub */
        if ((f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]))
            continue;
        /* :62-66  p = 0.5 * (v * is + is - 1)  ->  fmul(fadd(fma(v,is,is),-1),0.5) */
        float p[3][2];
        for (int n = 0; n < 3; ++n)
            for (int d = 0; d < 2; ++d)
                p[n][d] = (fmaf(f[3 * n + d], isf, isf) + -1.0f) * 0.5f;
        /* :69-72 adjugate rows */
        float fi[9];
        fi[0] = p[1][1] - p[2][1];
        fi[1] = p[2][0] - p[1][0];
        fi[2] = fmaf(p[1][0], p[2][1], -(p[2][0] * p[1][1]));
        fi[3] = p[2][1] - p[0][1];
        fi[4] = p[0][0] - p[2][0];
        fi[5] = fmaf(p[2][0], p[0][1], -(p[0][0] * p[2][1]));
        fi[6] = p[0][1] - p[1][1];
        fi[7] = p[1][0] - p[0][0];
        fi[8] = fmaf(p[0][0], p[1][1], -(p[1][0] * p[0][1]));
        /* :73-76 determinant */
        float den = fmaf(p[1][0], fi[3], fmaf(p[2][0], fi[6], p[0][0] * fi[0]));
        /* :77-79 IEEE divide */
        for (int k = 0; k < 9; ++k)
            o[k] = fi[k] / den;
    }
}

/* rasterize_cuda_kernel.cu:87-186 (forward_face_index_map_cuda_kernel_2)
 * + rasterize.py:50-52 (fim=-1, wim=0, depth=far) + rasterize.py:335-338
 * (vertical flip of fim/wim; depth is flipped too when returned).
 * Brute force: every pixel visits every face in index order.
 * fim (B,is,is) int32, wim (B,is,is,3), depth (B,is,is) or NULL.
 * flip_y != 0 applies torch.flip(dims=(1,)) to the outputs. */
typedef struct {
    const float *faces, *finv; int B, F, is; float near_, far_; int flip_y;
    int32_t *fim; float *wim, *depth;
} rast_ctx;

static void rast_range(void *vc, long begin, long end)
{
    const rast_ctx *c = (const rast_ctx *)vc;
    const float *faces = c->faces, *finv = c->finv;
    const int F = c->F, is = c->is, flip_y = c->flip_y;
    const float near_ = c->near_, far_ = c->far_;
    int32_t *fim = c->fim; float *wim = c->wim, *depth = c->depth;
    for (long i = begin; i < end; ++i) {
        const int bn = (int)(i / ((long)is * is));
        const int pn = (int)(i % ((long)is * is));
        const int yi = pn / is, xi = pn % is;
        /* :113-114 pixel centre in f64, rounded once to f32 */
        const float yp = (float)((2. * yi + 1 - is) / is);
        const float xp = (float)((2. * xi + 1 - is) / is);
        const float xf = (float)xi, yf = (float)yi;
        float depth_min = far_;
        int face_min = -1;
        float wmin[3] = {0.f, 0.f, 0.f};
        for (int fn = 0; fn < F; ++fn) {
            const float *f = faces + ((size_t)bn * F + fn) * 9;
            const float *fi = finv + ((size_t)bn * F + fn) * 9;
            /* :128 cull */
            if ((f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]))
                continue;
            /* :132-135 three edge tests */
            if (((yp - f[1]) * (f[3] - f[0]) < (xp - f[0]) * (f[4] - f[1])) ||
                ((yp - f[4]) * (f[6] - f[3]) < (xp - f[3]) * (f[7] - f[4])) ||
                ((yp - f[7]) * (f[0] - f[6]) < (xp - f[6]) * (f[1] - f[7])))
                continue;
            /* :139-141  w = fadd(fma(a, x, fmul(b, y)), c) */
            float w[3];
            for (int k = 0; k < 3; ++k)
                w[k] = fmaf(fi[3 * k], xf, fi[3 * k + 1] * yf) + fi[3 * k + 2];
            /* :144-148 clamp in f64 (fmax/fmin drop NaN like CUDA's) and sum */
            float w_sum = 0.f;
            for (int k = 0; k < 3; ++k) {
                w[k] = (float)fmin(fmax((double)w[k], 0.), 1.);
                w_sum += w[k];
            }
            /* :149-151 */
            for (int k = 0; k < 3; ++k)
                w[k] /= w_sum;
            /* :153  zp = 1 / (w0/z0 + w1/z1 + w2/z2) */
            const float zp = 1.0f / ((w[0] / f[2] + w[1] / f[5]) + w[2] / f[8]);
            if (zp <= near_ || far_ <= zp)
                continue;
            /* :159 strict less: lowest face index wins ties */
            if (zp < depth_min) {
                depth_min = zp;
                face_min = fn;
                wmin[0] = w[0]; wmin[1] = w[1]; wmin[2] = w[2];
            }
        }
        const int yo = flip_y ? (is - 1 - yi) : yi;
        const size_t o = ((size_t)bn * is + yo) * is + xi;
        if (face_min >= 0) {
            fim[o] = face_min;
            wim[3 * o] = wmin[0]; wim[3 * o + 1] = wmin[1]; wim[3 * o + 2] = wmin[2];
            if (depth) depth[o] = depth_min;
        } else {
            fim[o] = -1;
            wim[3 * o] = 0.f; wim[3 * o + 1] = 0.f; wim[3 * o + 2] = 0.f;
            if (depth) depth[o] = far_;
        }
    }
}

void oracle_rasterize(const float *faces, int B, int F, int is, float near_, float far_,
                      int flip_y, int32_t *fim, float *wim, float *depth)
{
    float *finv = (float *)calloc((size_t)B * F * 9, sizeof(float));
    oracle_face_inv(faces, (long)B * F, is, finv);
    rast_ctx c = {faces, finv, B, F, is, near_, far_, flip_y, fim, wim, depth};
    /* interleave rows across threads poorly balanced?  chunks are contiguous
     * pixel ranges; good enough for a checker. */
    parallel_for((long)B * is * is, rast_range, &c);
    free(finv);
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* block_extractor/block_extractor_kernel.cu:21-85
 * (kernel_block_extractor_update_output).  src (B,C,Hs,Ws), flow (B,2,Hf,Wf),
 * out (B,C,k*Hf,k*Wf); border = index clamp; taps accumulated LT,RT,LB,RB. */
void oracle_block_extract(const float *src, const float *flow, float *out,
                          int B, int C, int Hs, int Ws, int Hf, int Wf, int k)
{
    const int Ho = k * Hf, Wo = k * Wf;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const float *s = src + ((size_t)b * C + c) * Hs * Ws;
            float *o = out + ((size_t)b * C + c) * Ho * Wo;
            for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x) {
                    const int yf = y / k, xf = x / k;                       /* :52-53 */
                    const int yo = y % k - k / 2, xo = x % k - k / 2;       /* :54-55 */
                    const float fy = flow[(((size_t)b * 2 + 1) * Hf + yf) * Wf + xf] + (float)yo;
                    const float fx = flow[(((size_t)b * 2 + 0) * Hf + yf) * Wf + xf] + (float)xo;
                    const float dy = fy + (float)yf, dx = fx + (float)xf;   /* :61-62 */
                    const float fdx = floorf(dx), fdy = floorf(dy);
                    const int xL = clampi((int)fdx, 0, Ws - 1);             /* :64-67 */
                    const int xR = clampi((int)(fdx + 1.f), 0, Ws - 1);
                    const int yT = clampi((int)fdy, 0, Hs - 1);
                    const int yB = clampi((int)(fdy + 1.f), 0, Hs - 1);
                    const float xLp = 1.f - (dx - fdx), xRp = dx - fdx;     /* :68-71 */
                    const float yTp = 1.f - (dy - fdy), yBp = dy - fdy;
                    float acc = 0.0f;                                       /* :73-79 */
                    acc = fmaf(xLp * yTp, s[yT * Ws + xL], acc);
                    acc = fmaf(xRp * yTp, s[yT * Ws + xR], acc);
                    acc = fmaf(xLp * yBp, s[yB * Ws + xL], acc);
                    acc = fmaf(xRp * yBp, s[yB * Ws + xR], acc);
                    o[(size_t)y * Wo + x] = acc;
                }
        }
}

/* local_attn_reshape/local_attn_reshape_kernel.cu:21-61
 * (kernel_local_attn_reshape_update_output): (B,k*k,H,W) -> (B,1,kH,kW),
 * out[y,x] = in[(y%k)*k + x%k, y/k, x/k]. */
void oracle_local_attn_reshape(const float *in, float *out, int B, int k, int H, int W)
{
    const int Ho = k * H, Wo = k * W;
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < Ho; ++y)
            for (int x = 0; x < Wo; ++x) {
                const int ch = (y % k) * k + (x % k);
                out[((size_t)b * Ho + y) * Wo + x] =
                    in[(((size_t)b * k * k + ch) * H + y / k) * W + x / k];
            }
}


/* block_extractor/block_extractor_kernel.cu:86-166 (kernel_block_extractor_backward), sequential: for every output element
 * its gradient is scattered to the four source taps with the forward's bilinear weights, and the flow gradient of the
 * (yf, xf) cell accumulates grad * d(sample)/d(flow) over channels and the k x k taps.  grad_src / grad_flow are ADDED to
 * (the reference's autograd Function passes zero-filled tensors, block_extractor.py:36-37). */
void oracle_block_extract_backward(const float *src, const float *flow, const float *gout, float *gsrc, float *gflow,
                                   int B, int C, int Hs, int Ws, int Hf, int Wf, int k)
{
    const int Ho = k * Hf, Wo = k * Wf;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const float *s = src + ((size_t)b * C + c) * Hs * Ws;
            float *gs = gsrc + ((size_t)b * C + c) * Hs * Ws;
            const float *go = gout + ((size_t)b * C + c) * Ho * Wo;
            for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x) {
                    const int yf = y / k, xf = x / k;
                    const int yo = y % k - k / 2, xo = x % k - k / 2;
                    const float fy = flow[(((size_t)b * 2 + 1) * Hf + yf) * Wf + xf] + (float)yo;
                    const float fx = flow[(((size_t)b * 2 + 0) * Hf + yf) * Wf + xf] + (float)xo;
                    const float dy = fy + (float)yf, dx = fx + (float)xf;
                    const float fdx = floorf(dx), fdy = floorf(dy);
                    const int xL = clampi((int)fdx, 0, Ws - 1), xR = clampi((int)(fdx + 1.f), 0, Ws - 1);
                    const int yT = clampi((int)fdy, 0, Hs - 1), yB = clampi((int)(fdy + 1.f), 0, Hs - 1);
                    const float xLp = 1.f - (dx - fdx), xRp = dx - fdx;
                    const float yTp = 1.f - (dy - fdy), yBp = dy - fdy;
                    const float vLT = s[yT * Ws + xL], vRT = s[yT * Ws + xR], vLB = s[yB * Ws + xL], vRB = s[yB * Ws + xR];
                    const float g = go[(size_t)y * Wo + x];
                    gs[yT * Ws + xL] += g * xLp * yTp;                       /* :152-155 */
                    gs[yT * Ws + xR] += g * xRp * yTp;
                    gs[yB * Ws + xL] += g * xLp * yBp;
                    gs[yB * Ws + xR] += g * xRp * yBp;
                    const float gy = g * (-xLp * vLT - xRp * vRT + xLp * vLB + xRp * vRB);     /* :157-158 */
                    const float gx = g * (-yTp * vLT - yBp * vLB + yTp * vRT + yBp * vRB);
                    gflow[(((size_t)b * 2 + 1) * Hf + yf) * Wf + xf] += gy;  /* :161-162 */
                    gflow[(((size_t)b * 2 + 0) * Hf + yf) * Wf + xf] += gx;
                }
        }
}

/* local_attn_reshape/local_attn_reshape_kernel.cu:62-104 (kernel_local_attn_reshape_backward): the forward is a permutation,
 * so grad_in[b, (y%k)*k + x%k, y/k, x/k] += grad_out[b, 0, y, x]. */
void oracle_local_attn_reshape_backward(const float *gout, float *gin, int B, int k, int H, int W)
{
    const int Ho = k * H, Wo = k * W;
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < Ho; ++y)
            for (int x = 0; x < Wo; ++x) {
                const int ch = (y % k) * k + (x % k);
                gin[(((size_t)b * k * k + ch) * H + y / k) * W + x / k] += gout[((size_t)b * Ho + y) * Wo + x];
            }
}
