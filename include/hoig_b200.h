/*
 * hoig_b200.h -- C ABI of the B200-native HOGAN generator hot path.
 *
 * One shared library (libhoig_b200.so), plain pointers and sizes, explicit
 * stream, no torch types.  Every entry point returns 0 on success or a
 * negative hoigStatus; none allocates device memory (workspaces are passed
 * in), none throws, all are asynchronous on `stream`.  hoig_last_error()
 * returns a thread-local description of the most recent failure.
 *
 * Each function cites the reference interface it replaces
 * (paths relative to /root/reference/HOIG_HOv3).
 *
 * Activation tensors inside the generator are NHWC ("pixel-major"): element
 * (n,y,x,c) of a tensor with pixel stride `ld` lives at
 * base[((n*H + y)*W + x)*ld + c].  `ld >= C` lets a tensor be a channel slice
 * of a wider buffer (the U-Net skip concatenations are never materialised
 * separately).  dtype: HOIG_F32 (SIMT fp32 parity path), HOIG_BF16 or HOIG_F16
 * (tcgen05 tensor-core path, kind::f16, fp32 accumulate; same rate, 8 vs 11 mantissa bits).
 */
#ifndef HOIG_B200_H_
#define HOIG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *hoigStream_t; /* a cudaStream_t */

typedef enum {
    HOIG_OK = 0,
    HOIG_ERR_INVALID = -1,   /* bad argument (shape, alignment, null) */
    HOIG_ERR_WORKSPACE = -2, /* workspace too small */
    HOIG_ERR_CUDA = -3,      /* a CUDA runtime/driver call or launch failed */
    HOIG_ERR_ARCH = -4       /* device is not sm_100 */
} hoigStatus;

typedef enum { HOIG_F32 = 0, HOIG_BF16 = 1, HOIG_F16 = 2 } hoigDType;
typedef enum { HOIG_ACT_NONE = 0, HOIG_ACT_RELU = 1, HOIG_ACT_LEAKY = 2, HOIG_ACT_TANH = 3, HOIG_ACT_SIGMOID = 4 } hoigAct;
typedef enum { HOIG_CONV = 0, HOIG_CONV_TRANSPOSED = 1, HOIG_CONV_LOCAL_ATTN = 2 } hoigConvMode;

const char *hoig_version(void);
const char *hoig_last_error(void);
/* 0 if the current device is compute capability 10.x, else HOIG_ERR_ARCH / HOIG_ERR_CUDA. */
int hoig_check_device(void);

/* ------------------------------------------------------------------ stage R
 * Condition rasterizer.  Replaces
 *   thirdparty/neural_renderer/neural_renderer/cuda/rasterize_cuda.cpp:70-95
 *   (forward_face_index_map: kernels _1 and _2 of rasterize_cuda_kernel.cu:41-186)
 * plus the output initialisation of rasterize.py:50-52 and the vertical flip
 * of rasterize.py:335-338, batched over B meshes in one call.
 *   faces  (B,F,3,3) f32 camera-space xyz per face vertex
 *   fim    (B,is,is) int32, -1 = background        [bit-exact with the reference]
 *   wim    (B,is,is,3) f32 barycentric weights, 0 on background
 *   depth  (B,is,is) f32 or NULL, `far` on background
 *   flip_y != 0 writes row (is-1-y), i.e. torch.flip(dims=(1,)).
 * image_size must be a multiple of 16.  Workspace: hoig_rasterize_workspace_bytes. */
size_t hoig_rasterize_workspace_bytes(int B, int F, int image_size);
int hoig_rasterize_fim_wim(const float *faces, int B, int F, int image_size, float near, float far,
                           int flip_y, int32_t *fim, float *wim, float *depth,
                           void *workspace, size_t workspace_bytes, hoigStream_t stream);

/* Per-face inverse matrices only (kernel _1, rasterize_cuda_kernel.cu:41-84);
 * faces_inv (B*F,9) must be zero-filled by the caller like rasterize.py:164. */
int hoig_face_inv(const float *faces, int64_t BF, int image_size, float *faces_inv, hoigStream_t stream);

/* R0: projection + look_at + vertices_to_faces fused (utils/nmr.py:109-140,:506;
 * neural_renderer/look_at.py:6-62 with eye (0,0,eye_z), identity rotation;
 * vertices_to_faces.py:4-22).  verts (B,V,3), cam (B,15), faces_idx (F,3) int32
 * shared by the batch -> faces (B,F,3,3). */
int hoig_project_faces(const float *verts, const float *cam, const int32_t *faces_idx,
                       int B, int V, int F, float eye_z, float *faces, hoigStream_t stream);

/* R4-R7: condition maps from fim/wim (utils/nmr.py:567-595 encode_fim/encode_sem,
 * models/trainer.py:71-72 one-hot seg + hand mask, utils/nmr.py:874-925
 * cal_bc_transform's T).  Outputs NCHW f32:
 *   cond (B,3,is,is) = map_fn[fim]; seg (B,15,is,is) = (sem_full[fim]==i), i=1..15;
 *   not_hand (B,1,is,is) = 1 - [(fim!=-1)&(fim<n_hand_faces)]   (pre-erosion);
 *   T (B,is,is,2) or NULL: sum_k src_faces[fim_ref][k].xy(y negated) * wim_ref[k], -2 where fim_ref==-1.
 * map_fn (F+1,3), sem_full (F+1) with the background row last (fim=-1). */
int hoig_condition_maps(const int32_t *fim, int B, int F, int image_size, const float *map_fn,
                        const float *sem_full, int n_hand_faces, float *cond, float *seg,
                        float *not_hand, hoigStream_t stream);
int hoig_bc_transform(const float *src_faces, const int32_t *fim_ref, const float *wim_ref,
                      int B, int F, int image_size, float *T, hoigStream_t stream);
/* R6 utils/util.py:142-153 erode (pad value 1, all-ones ks x ks window). in/out (B,1,H,W) f32. */
int hoig_erode(const float *in, float *out, int B, int H, int W, int ks, hoigStream_t stream);

/* ---- row N1: the tail of HandRecoveryFlow.forward (models/trainer.py:66-145) for the whole batch in ONE launch: condition
 * tables, one-hot segmentation, hand / background masks (3x3 erosions, the 15x15 erosion of the source background), the dense
 * correspondence T masked to the target hand region, and the assembly of every Generator.forward input.  All tensors NCHW f32
 * unless noted; the 1 + 8 erosion windows are evaluated from the face-index maps, nothing intermediate touches memory.
 *   in : fim_src, fim_ref (B,is,is) int32; wim_ref (B,is,is,3); src_faces (B,F,3,3) (hoig_project_faces); src_img (B,3,is,is);
 *        render_src / render_ref (B,3,is,is) (stage R8); map_fn (F+1,3), sem_full (F+1) with the background row last.
 *   out: bg_inputs (B,4), {src,tsf}_obj_inputs (B,3), {src,tsf}_obj_conds (B,12), {src,tsf}_hand_inputs (B,3),
 *        {src,tsf}_hand_conds (B,3), T (B,is,is,2), {src,ref}_mask_bg, {src,ref}_mask_hand (B,1). */
typedef struct hoigCondInputsDesc {
    const int32_t *fim_src, *fim_ref;
    const float *wim_ref, *src_faces, *src_img, *render_src, *render_ref, *map_fn, *sem_full;
    float *bg_inputs, *src_obj_inputs, *src_obj_conds, *src_hand_inputs, *src_hand_conds;
    float *tsf_obj_inputs, *tsf_obj_conds, *tsf_hand_inputs, *tsf_hand_conds, *T;
    float *src_mask_bg, *ref_mask_bg, *src_mask_hand, *ref_mask_hand;
    int B, F, image_size, n_hand_faces, bg_erode_ks;
} hoigCondInputsDesc;
int hoig_condition_inputs(const hoigCondInputsDesc *desc, hoigStream_t stream);

/* ---- stage R8: UV-texture warp (utils/nmr.py:973-1100, models/trainer.py:83-87)
 * hoig_uv_backward_warp: nmr.py:973-1040.  For every atlas pixel p with fim_uv[p] != -1 (fim_uv (Hu,Wu) int32, wim_uv (Hu,Wu,3),
 * shared by the batch):  T[b,p] = sum_k src_faces[b][fim_uv[p]][k].xy (y negated, trainer.py:67-68) * wim_uv[p][k], else -2;
 * O[b,p] = 1 - [one of the 3x3 neighbours (clamped) of trunc((T+1)/2*(is-1)) in src_fim[b] (is x is) equals fim_uv[p]], else 0.
 * src_faces (B,F,3,3) f32 as produced by hoig_project_faces; T (B,Hu,Wu,2), O (B,1,Hu,Wu) f32. */
int hoig_uv_backward_warp(const float *src_faces, const int32_t *fim_uv, const float *wim_uv, const int32_t *src_fim,
                          int B, int F, int Hu, int Wu, int image_size, float *T, float *O, hoigStream_t stream);
/* nmr.py:1068-1100 sample_from_texture_dense: T[b,p] = sum_k uv_coord[fim[b,p]][k] * wim[b,p][k], -2 where fim == -1;
 * uv_coord (F,3,2) f32 shared by the batch; fim (B,H,W) int32; wim (B,H,W,3); T (B,H,W,2). */
int hoig_sample_texture_dense(const float *uv_coord, const int32_t *fim, const float *wim, int B, int H, int W, float *T,
                              hoigStream_t stream);
/* F.grid_sample(im, grid, mode='bilinear', padding_mode='zeros', align_corners) on NCHW f32 images:
 * im (B,C,Hi,Wi), grid (B,Ho,Wo,2) -> out (B,C,Ho,Wo)  (nmr.py:1047 uses align_corners=False, trainer.py:85,87 True). */
int hoig_grid_sample_nchw(const float *im, int B, int C, int Hi, int Wi, const float *grid, int Ho, int Wo, int align_corners,
                          float *out, hoigStream_t stream);
/* nmr.py:1049-1056: O <- 1 - erode3(1 - erode3(O)); syn <- syn*(1-O) + O; columns >= x0 of syn are replaced by the stock object
 * texture `preload` ((Hu, Wu-x0, C) HWC f32) when it is non-NULL.  syn (B,C,Hu,Wu) in place, O (B,1,Hu,Wu). */
int hoig_uv_texture_compose(float *syn, const float *O, const float *preload, int B, int C, int Hu, int Wu, int x0,
                            hoigStream_t stream);

/* ------------------------------------------------- reference op boundary B2
 * thirdparty/block_extractor/block_extractor_cuda.cc:5-16 (forward):
 *   source (B,C,Hs,Ws), flow (B,2,Hf,Wf), out (B,C,k*Hf,k*Wf), all NCHW f32. */
int hoig_block_extract_f32(const float *source, const float *flow, float *out, int B, int C,
                           int Hs, int Ws, int Hf, int Wf, int k, hoigStream_t stream);
/* thirdparty/local_attn_reshape/local_attn_reshape_cuda.cc:5-13 (forward):
 *   in (B,k*k,H,W) -> out (B,1,k*H,k*W). */
int hoig_local_attn_reshape_f32(const float *in, float *out, int B, int k, int H, int W, hoigStream_t stream);
/* Backward of the two ops (row N3; thirdparty/block_extractor/block_extractor_cuda.cc:18-33, block_extractor_kernel.cu:86-166 and
 * thirdparty/local_attn_reshape/local_attn_reshape_cuda.cc:17-29, local_attn_reshape_kernel.cu:62-104).  Like the reference
 * kernels they ADD into grad_source / grad_flow / grad_in, which the reference autograd Functions pass zero-filled.
 * grad_out (B,C,k*Hf,k*Wf) resp. (B,1,k*H,k*W); grad_source like source; grad_flow like flow; grad_in (B,k*k,H,W). */
int hoig_block_extract_backward_f32(const float *source, const float *flow, const float *grad_out, float *grad_source,
                                    float *grad_flow, int B, int C, int Hs, int Ws, int Hf, int Wf, int k, hoigStream_t stream);
int hoig_local_attn_reshape_backward_f32(const float *grad_out, float *grad_in, int B, int k, int H, int W, hoigStream_t stream);

/* ----------------------------------------------------------------- stage G
 * Building blocks of Generator.forward (models/networks/generator.py:347-491,
 * spade.py:24-38, extract_attn.py:23-29).  The host-side schedule lives in
 * hoig_b200/generator.py (GeneratorB200, drop-in for the reference Generator). */

typedef struct {
    int dtype;          /* hoigDType of activations and packed weights */
    int mode;           /* hoigConvMode */
    int N, H, W;        /* input batch / spatial size */
    int C0, C1;         /* channels taken from src0 / src1 (C1 = 0: single source); multiples of 8 */
    int OH, OW, Cout;   /* output size; Cout = logical output channels */
    int KH, KW, stride, pad;  /* pad = vertical padding; horizontal padding is pad_w (last field) */
    const void *src0; int64_t ld0;
    const void *src1; int64_t ld1;
    const void *weight; /* packed [rows][cols]; conv / local attention: k = (r*KW + s)*(C0+C1) + c;
                         * transposed: parity-block x tap-block matrix (hoig_conv_packed_dims) */
    const float *bias;  /* [Cout] or NULL */
    int act;            /* hoigAct applied after bias (+ residual) */
    const void *residual; int64_t ldr; /* NHWC (N,OH,OW,Cout) added before `act`, or NULL */
    void *dst; int64_t ldd;            /* NHWC (N,OH,OW,>=Cout) */
    double *stats;      /* NULL or [N][Cout][2]: += per-plane sum / sum of squares of the stored values */
    const float *flow;  /* HOIG_CONV_LOCAL_ATTN only: (N,H,W,2) pixel-unit offsets (x,y) */
    const int *act_table; /* NULL or [Cout] device array of hoigAct codes overriding `act` per output channel */
    int pad_w;          /* horizontal padding (set equal to pad for square kernels) */
    /* SPADE-modulating epilogue (spade.py:33-38 fused into the gamma/beta GEMM; 16-bit dtypes only).  When spade_x is
     * non-NULL the GEMM's Cout = 2*C columns are (gamma, beta) in blocks of 8 channels [g0..g7 b0..b7] (weights and bias
     * packed that way) and the kernel writes C channels:
     *   dst[pixel][c] = relu( (x[pixel][c] - mean_c) * rstd_c * (1 + gamma_c) + beta_c ),
     * x = spade_x (N,OH,OW,C) with pixel stride ld_spade_x, mean/rstd per (image, channel) from spade_stats ([N][C][2] sum /
     * sum of squares) and spade_eps.  The gamma/beta tensor never reaches memory.  stats, residual, act_table must be NULL. */
    const void *spade_x; int64_t ld_spade_x;
    const double *spade_stats;
    float spade_eps;
} hoigConvDesc;

/* Rows / columns of the packed weight matrix for a given problem.  Conv / local attention: rows = Cout padded
 * to 16, cols = KH*KW*Cin padded to 64.  HOIG_CONV_TRANSPOSED (k3 s2 p1 op1): rows = 4*Cout (one block per output
 * parity (a,b), block index a*2+b), cols = 4*Cin padded to 64 (one block per input tap (dy,dx) of the 2x2
 * neighbourhood, index dy*2+dx); block [(a,b)][(dy,dx)] = W[:, :, a+pad-2dy, b+pad-2dx]^T or zero. */
int hoig_conv_packed_dims(int mode, int Cout, int KH, int KW, int Cin, int stride, int pad, int *rows, int *cols);
/* Implicit-GEMM convolution (nn.Conv2d / nn.ConvTranspose2d k3 s2 p1 op1 /
 * the k5 s5 conv over cat[BlockExtractor(tgt,0), BlockExtractor(src,flow)] of
 * extract_attn.py:24-26 with both extractions fused into the operand gather).
 * dtype BF16: tcgen05.mma (kind::f16, fp32 accumulators in TMEM), weights by TMA.
 * dtype F32 : SIMT fp32 FFMA. */
int hoig_conv2d(const hoigConvDesc *desc, hoigStream_t stream);
/* Test hooks (not used by the product schedule): the SIMT kernel for any dtype, and a switch that
 * forces the bf16 kernel's A operand through the cp.async gather path instead of TMA boxes. */
int hoig_conv2d_simt(const hoigConvDesc *desc, hoigStream_t stream);
void hoig_set_umma_gather_only(int on);
/* Test hook: 0 = one CTA per tile, 1 (default) = CTA pairs (cta_group::2) for long reductions, 2 = pairs wherever legal. */
void hoig_set_umma_pair_mode(int on);
/* Test hook: 1 (default) = narrow-N tensor-core convs run two MMA issue pipelines per CTA, 0 = one. */
void hoig_set_umma_dual_mode(int on);
/* Test hook: 1 (default) = weight matrices that fit stay resident in shared memory, 0 = always streamed through the ring. */
void hoig_set_umma_bres_mode(int on);
/* Test hook: 1 (default) = full-row tiles of regular stride-1 convs load one activation box per kernel row, 0 = one per tap. */
void hoig_set_umma_halo_mode(int on);
/* Test hook: 1 (default) = kh x 1 convs (the re-associated 7x7 stems / heads) run on 8 x 16 pixel tiles that load one activation box
 * per 64 channels for all kh vertical taps (weights resident), 0 = one box per tap. */
void hoig_set_umma_vhalo_mode(int on);
/* Test hook: 1 (default) = hoig_attn_combine runs its weighted source-patch sum on the tensor cores with the source window of an 8x8
 * pixel tile staged in shared memory (16-bit dtypes, k = 5, h % 8 == 0, C % 64 == 0), 0 = per-pixel gathers for every pixel. */
void hoig_set_attn_tc_mode(int on);
/* Tuning hook: pixels per rasterizer band (256..16384; the band's 64-bit key buffer lives in shared memory). */
void hoig_set_rasterizer_band_pixels(int n);

/* NCHW f32 (B,C,H,W) -> NHWC dtype with `Cpad` channels (extra channels zero). */
int hoig_nchw_to_nhwc(const float *src, int B, int C, int H, int W, void *dst, int64_t ldd, int Cpad,
                      int dtype, hoigStream_t stream);
/* NHWC dtype -> NCHW f32 (first C channels). */
int hoig_nhwc_to_nchw(const void *src, int64_t lds, int dtype, int B, int C, int H, int W, float *dst,
                      hoigStream_t stream);
/* spade.py:30 F.interpolate(segmap, size, 'nearest') fused with the layout change:
 * seg NCHW f32 (B,C,Hi,Wi) -> NHWC dtype (B,Ho,Wo,Cpad). */
int hoig_seg_resize_nearest(const float *seg, int B, int C, int Hi, int Wi, void *dst, int64_t ldd, int Cpad,
                            int Ho, int Wo, int dtype, hoigStream_t stream);
/* per-(n,c) sum / sum-of-squares of an NHWC tensor into stats [N][C][2] (+=). */
int hoig_plane_stats(const void *x, int64_t ldx, int dtype, int N, int HW, int C, double *stats, hoigStream_t stream);
/* nn.InstanceNorm2d (eps 1e-5, biased variance) finalisation:
 *   y = (x-mean)*rstd;  affine: y = y*gamma[c]+beta[c]   (generator.py:16-22 ...)
 *   SPADE (spade.py:36): y = y*(1+gb[...,c]) + gb[...,C+c]   (gb NHWC with 2C channels)
 *   y += residual (optional, generator.py:31);  relu optional. */
int hoig_instnorm_apply(const void *x, int64_t ldx, const double *stats, const float *gamma, const float *beta,
                        const void *gb, int64_t ldgb, const void *residual, int64_t ldr, int relu,
                        void *dst, int64_t ldd, int dtype, int N, int HW, int C, float eps, hoigStream_t stream);
/* generator.py:466-473 resize_trans (bilinear, align_corners=True, size=(h,h)) fused with
 * generator.py:484-488  flow = T_scale - idt   ('ij' identity grid, quirks Q1-Q3).
 * T (B,Hi,Wi,2) f32 -> flow (B,h,h,2) f32.  subtract_identity = 0 returns T_scale itself
 * (the grid generator.py:475-478 hands to grid_sample for the non-attention variants). */
int hoig_resize_flow(const float *T, int B, int Hi, int Wi, int h, int subtract_identity, float *flow,
                     hoigStream_t stream);
/* extract_attn.py:24-28 tail: logits = W2 . hidden + b2 (conv1x1 128->k*k), softmax over k*k,
 * out = tgt + (1/k^2) * sum_t softmax_t * BlockExtractor(src,flow)[tap t]   (avg_pool of attn*block)
 * hidden (N,h,h,Chid) dtype, w2 [k*k][Chid] f32, b2 [k*k] f32; src/tgt/dst NHWC (N,h,h,C). */
int hoig_attn_finish(const void *hidden, int64_t ldh, int Chid, const float *w2, const float *b2,
                     const void *src, int64_t lds, const float *flow, const void *tgt, int64_t ldt,
                     void *dst, int64_t ldd, int dtype, int N, int h, int C, int k,
                     const void *unfold, int64_t ldu, hoigStream_t stream);
/* extract_attn.py:24-25 for the tensor-core path: out (N,h,h,2*k*k*C), channel t*2C + c = BlockExtractor(tgt,0)
 * tap t (c < C) | BlockExtractor(src,flow) tap t (c >= C); the k5s5 conv over cat[block_target, block_source]
 * is then hoig_conv2d with a 1x1 kernel over this tensor, and hoig_attn_finish can read the source taps from
 * it (`unfold`) instead of re-sampling. */
int hoig_attn_unfold(const void *src, int64_t lds, const void *tgt, int64_t ldt, const float *flow, void *out,
                     int64_t ldo, int dtype, int N, int h, int C, int k, hoigStream_t stream);
/* ---- local attention, tensor-core formulation (extract_attn.py:19-28 re-associated) -----------------------
 * The k x k stride-k conv over cat[BlockExtractor(tgt,0), BlockExtractor(src,flow)] is linear in the extracted
 * taps and every tap of one pixel shares the same bilinear fractions (block_extractor_kernel.cu:57-82), so
 *     conv(cat[...])(p) = Gt(p) + sum_{q in 2x2} w_q(p) * Gs(p0 + q),
 * with Gt = k x k conv of the replicate-padded target, Gs = k x k conv of the replicate-padded source evaluated on
 * the (h+k-1)^2 extended grid (each tap clamps on its own, so a corner up to k/2 pixels outside the image still
 * sees distinct taps), p0 = floor(p + flow(p)).  Same sum, different association: results agree with the
 * reference to rounding (tests gate the tensor-core path at the north-star tolerance); the fp32 parity path keeps
 * the reference's order of operations (HOIG_CONV_LOCAL_ATTN).
 *
 * hoig_replicate_pad: dst[n, y, x, :] = src[n, clamp(y-pad), clamp(x-pad), :], dst (N, h+2pad, h+2pad, C). */
int hoig_replicate_pad(const void *src, int64_t lds, void *dst, int64_t ldd, int dtype, int N, int h, int C, int pad,
                       hoigStream_t stream);
/* Dense KH x KW stride-1 convolution over padded rasters (N, Hp, Wp, C) -> (N, Hp, Wp, Cout), bf16 / fp16 tensor
 * cores, raw accumulators (no bias / activation).  out[m] = sum_{r,s,c} W[n][(r*KW+s)*C + c] * in[m + (r-KH/2)*Wp +
 * (s-KW/2)][c] over the flattened pixel index m (zero outside the tensor): pixels at least KH/2 rows and KW/2
 * columns inside their image get the exact convolution, the others hold finite don't-care values.  `weight` is the
 * packed (Cout, KH*KW*C) matrix of hoig_conv_packed_dims.  Up to two problems (segments) share one launch. */
typedef struct hoigHaloConvSeg {
    const void *src; int64_t ld;
    int N, Hp, Wp, C;
    const void *weight;
    void *dst; int64_t ldd;
} hoigHaloConvSeg;
int hoig_conv2d_halo(int dtype, int KH, int KW, int Cout, const hoigHaloConvSeg *segs, int nsegs, hoigStream_t stream);
/* Test hook: A-operand strategy of hoig_conv2d_halo (0 = one box per kernel row, taps by shifted descriptors;
 * 2 = one box per tap). */
void hoig_set_halo_variant(int v);
/* Tail of the attention block for the formulation above: hidden = LeakyReLU(Gt + bilinear(Gs) + b1); 1x1 conv to
 * k*k logits; softmax; dst = tgt + (1/k^2) * sum_t a_t * BlockExtractor(src,flow)_t, the 25 x 4 bilinear taps
 * folded into one (k+1)^2 patch of coefficients.  gt: (N, h+k-1, h+k-1, Chid) with the image at offset k/2;
 * gs: (N, h+2(k-1), h+2(k-1), Chid) with extended-grid cell (0,0) at offset k/2; src/tgt/dst (N,h,h,C). */
int hoig_attn_combine(const void *gt, int64_t ldgt, const void *gs, int64_t ldgs, int Chid, const float *b1, const float *w2,
                      const float *b2, const void *src, int64_t lds, const float *flow, const void *tgt, int64_t ldt,
                      void *dst, int64_t ldd, int dtype, int N, int h, int C, int k, hoigStream_t stream);
/* generator.py:475-478 stn: F.grid_sample(x, grid) bilinear / zeros / align_corners=False;
 * x, dst NHWC (N,h,h,C); grid (N,h,h,2) f32; dst = tgt + sample when tgt != NULL. */
int hoig_grid_sample(const void *x, int64_t ldx, const float *grid, const void *tgt, int64_t ldt,
                     void *dst, int64_t ldd, int dtype, int N, int h, int C, hoigStream_t stream);
/* spade.py:30-31 for the tensor-core path: nearest resize of the segmentation map fused with the 3x3 im2col of
 * mlp_shared's input.  seg NCHW f32 (B,C,Hi,Wi) -> NHWC dtype (B,Ho,Wo,Kpad), channel t*C + c = resized seg channel c at
 * (y + t/3 - 1, x + t%3 - 1), zero outside the image and for channels >= 9*C.  mlp_shared is then a 1x1 hoig_conv2d over
 * this tensor with the weight columns in the same (t, c) order. */
int hoig_seg_unfold3(const float *seg, int B, int C, int Hi, int Wi, void *dst, int64_t ldd, int Kpad, int Ho, int Wo,
                     int dtype, hoigStream_t stream);
/* 7x7 convs with few input or output channels (generator.py:99,125,151,223-241) are computed as a 7x1
 * "vertical taps" implicit GEMM plus a horizontal (un)fold, which cuts the operand traffic 7x:
 *  - stems:  hoig_hunfold_nchw builds x7[b,y,x, s*C + c] = x[b,c,y,x+s-k/2] (zero outside, channels padded to Cpad)
 *            straight from the NCHW f32 network input; the stem is then a kx1 conv over x7;
 *  - heads:  a kx1 conv produces Z[b,y,x, s*G + g] for the G head outputs and hoig_hfold_nchw sums
 *            y[b,g,y,x] = act_g( sum_s Z[b,y,x+s-k/2, s*G + g] ) into NCHW f32 (exact re-association of the sum).
 * hoig_hfold_nchw writes up to 4 output tensors: segment i takes channels [seg_c0[i], seg_c0[i]+seg_n[i]). */
int hoig_hunfold_nchw(const float *src, int B, int C, int H, int W, int k, void *dst, int64_t ldd, int Cpad, int dtype,
                      hoigStream_t stream);
int hoig_hfold_nchw(const void *z, int64_t ldz, int dtype, int B, int H, int W, int G, int k, const int *act_table,
                    int nseg, float *const *outs, const int *seg_c0, const int *seg_n, hoigStream_t stream);
/* models/trainer.py:400-401: img = m_bg*bg + (1-m_bg)*(obj*m_hand + hand*(1-m_hand)); NCHW f32, 3 channels. */
int hoig_composite(const float *img_bg, const float *obj, const float *hand, const float *mask_bg,
                   const float *mask_hand, float *out, int B, int HW, hoigStream_t stream);


/* ----------------------------------------------------------------- training (row N3, fp32)
 * The reference trains through torch.autograd of nn.Conv2d / nn.ConvTranspose2d / nn.InstanceNorm2d (models/networks/generator.py,
 * discriminator.py; step: models/trainer.py:417-434).  Here the data gradient of a convolution is itself a convolution and runs on
 * hoig_conv2d (hoig_b200/autograd.py); the two entry points below are the pieces that have no forward equivalent.
 *
 * hoig_conv2d_wgrad_f32: dW[co][r][s][ci] += sum_{n,oy,ox} g[n,oy,ox,co] * x[n, oy*stride + r - pad_h, ox*stride + s - pad_w, ci]
 *   x (N,H,W,Cin) / g (N,OH,OW,Cout) NHWC f32 with pixel strides ldx / ldg (multiples of 4); dW dense [Cout][KH][KW][Cin] f32 that
 *   the caller zero-fills (the kernel accumulates with atomics, like the reference's float-atomic backward kernels).
 *   nn.ConvTranspose2d: call it with the roles swapped (x := grad_out, g := input, the transposed conv's stride / padding).
 * hoig_instnorm_backward_f32: backward of y = gamma * (x - mean) * rstd + beta per (n, c) plane (biased variance, eps as forward):
 *   stats = the forward's double [N][C][2] (sum, sum of squares) of x; scratch = double [N][C][2] zero-filled by the caller;
 *   dx (N,HW,C); dgamma / dbeta (C) f32 accumulated (both NULL for affine=False, gamma NULL = 1). */
int hoig_conv2d_wgrad_f32(const float *x, int64_t ldx, const float *g, int64_t ldg, float *dw, int N, int H, int W, int Cin,
                          int OH, int OW, int Cout, int KH, int KW, int stride, int pad_h, int pad_w, hoigStream_t stream);
int hoig_instnorm_backward_f32(const float *x, int64_t ldx, const float *gy, int64_t ldg, const double *stats, const float *gamma,
                               float *dx, int64_t lddx, double *scratch, float *dgamma, float *dbeta, int N, int HW, int C,
                               float eps, hoigStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HOIG_B200_H_ */
