// conv_plan.cu -- host-side expansion of a hoigConvDesc into launch descriptions
// (tap tables, strided input views, output mapping, weight column ranges).
#include <string.h>

#include "conv_common.cuh"

namespace hoig {

static inline int ceil_to(int x, int m) { return (x + m - 1) / m * m; }
static inline int mod2(int q) { return ((q % 2) + 2) % 2; }

// taps of transposed-conv output parity `a` along one axis: kernel index r' contributes
// iff (a + pad - r') is even; the input offset is (a + pad - r') / 2.

void packed_layout(int mode, int Cout, int KH, int KW, int Cin, int stride, int pad, int *rows, int *cols, int col_off[4],
                   int col_len[4])
{
    *rows = ceil_to(Cout, 16);
    for (int i = 0; i < 4; ++i) col_off[i] = col_len[i] = 0;
    if (mode != HOIG_CONV_TRANSPOSED) {
        col_len[0] = ceil_to(KH * KW * Cin, 64);
        *cols = col_len[0];
        return;
    }
    (void)stride; (void)pad;
    // transposed (stride 2): one GEMM, rows = 4 parity blocks of Cout, cols = 2x2 input taps x Cin
    *rows = ceil_to(4 * Cout, 16);
    col_len[0] = ceil_to(4 * Cin, 64);
    *cols = col_len[0];
}

static InputView full_view(const void *base, int H, int W, int64_t ld)
{
    InputView v;
    v.base = base; v.H = H; v.W = W; v.sx = ld; v.sy = (int64_t)W * ld; v.sn = (int64_t)H * W * ld;
    return v;
}

int plan_conv(const hoigConvDesc *d, int bm, ConvPlan *plan)
{
    HOIG_REQUIRE(d && d->src0 && d->weight && d->dst, "conv2d: null pointer");
    HOIG_REQUIRE(d->dtype == HOIG_F32 || d->dtype == HOIG_BF16 || d->dtype == HOIG_F16, "conv2d: bad dtype %d", d->dtype);
    HOIG_REQUIRE(d->mode >= HOIG_CONV && d->mode <= HOIG_CONV_LOCAL_ATTN, "conv2d: bad mode %d", d->mode);
    HOIG_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->OH > 0 && d->OW > 0 && d->Cout > 0, "conv2d: bad shape");
    HOIG_REQUIRE(d->C0 > 0 && d->C0 % 8 == 0 && d->C1 >= 0 && d->C1 % 8 == 0, "conv2d: C0/C1 must be multiples of 8 (got %d,%d)", d->C0, d->C1);
    HOIG_REQUIRE(d->C1 == 0 || d->src1, "conv2d: C1 > 0 needs src1");
    HOIG_REQUIRE(d->KH > 0 && d->KW > 0 && d->KH * d->KW <= kMaxTaps && d->stride > 0 && d->pad >= 0 && d->pad_w >= 0,
                 "conv2d: bad kernel geometry");
    HOIG_REQUIRE(d->ld0 >= d->C0 && d->ld0 % 8 == 0 && (d->C1 == 0 || (d->ld1 >= d->C1 && d->ld1 % 8 == 0)),
                 "conv2d: source pixel stride must be a multiple of 8 and >= channels");
    HOIG_REQUIRE(d->ldd >= (d->spade_x ? d->Cout / 2 : d->Cout), "conv2d: ldd < Cout");
    if (d->spade_x) {
        HOIG_REQUIRE(d->dtype != HOIG_F32 && d->mode == HOIG_CONV && d->Cout % 16 == 0 && d->Cout <= 2048 && d->spade_stats && !d->stats &&
                         !d->residual && !d->act_table && d->ld_spade_x >= d->Cout / 2 && d->ld_spade_x % 8 == 0 && d->ldd % 8 == 0 &&
                         ((uintptr_t)d->spade_x % 16) == 0,
                     "conv2d(spade epilogue): needs a 16-bit dtype, plain conv mode, Cout = 2*C <= 2048 in 8-channel (gamma, beta) blocks, "
                     "spade_stats, no stats / residual / act_table, 16-byte aligned x");
    }
    HOIG_REQUIRE(d->mode != HOIG_CONV_TRANSPOSED || (d->KH == 3 && d->KW == 3 && d->pad == 1),
                 "conv2d(transposed): only the k3 s2 p1 op1 geometry of generator.py:118,201 is supported");
    HOIG_REQUIRE(!d->residual || d->ldr >= d->Cout, "conv2d: ldr < Cout");
    HOIG_REQUIRE(((uintptr_t)d->src0 % 16) == 0 && (!d->src1 || ((uintptr_t)d->src1 % 16) == 0) && ((uintptr_t)d->weight % 16) == 0,
                 "conv2d: sources and weights must be 16-byte aligned");
    const int esz = d->dtype == HOIG_F32 ? 4 : 2;
    const int Cin = d->C0 + d->C1;
    int rows, cols, col_off[4], col_len[4];
    packed_layout(d->mode, d->Cout, d->KH, d->KW, Cin, d->stride, d->pad, &rows, &cols, col_off, col_len);

    ConvParams base;
    memset(&base, 0, sizeof(base));
    base.mode = d->mode == HOIG_CONV_LOCAL_ATTN ? HOIG_CONV_LOCAL_ATTN : HOIG_CONV;
    base.N = d->N; base.C0 = d->C0; base.C1 = d->C1; base.Cin = Cin;
    base.Cout = d->Cout; base.Npad = rows; base.ldw = cols;
    base.bias = d->bias; base.act = d->act; base.act_table = d->act_table;
    base.residual = d->residual; base.ldr = d->ldr; base.dst = d->dst; base.ldd = d->ldd;
    base.stats = d->stats; base.flow = d->flow; base.KH = d->KH;
    base.spade_x = d->spade_x; base.ld_spade_x = d->ld_spade_x; base.spade_stats = d->spade_stats; base.spade_eps = d->spade_eps;
    base.nviews = 1;
    base.view[0] = full_view(d->src0, d->H, d->W, d->ld0);
    if (d->C1) base.view1 = full_view(d->src1, d->H, d->W, d->ld1);
    base.stride = 1;
    base.os = 1; base.OHf = d->OH; base.OWf = d->OW;

    if (d->mode == HOIG_CONV) {
        HOIG_REQUIRE(d->OH == (d->H + 2 * d->pad - d->KH) / d->stride + 1 && d->OW == (d->W + 2 * d->pad_w - d->KW) / d->stride + 1,
                     "conv2d: output size does not match geometry");
        ConvParams &p = plan->launch[0];
        p = base;
        plan->n = 1;
        p.GH = d->OH; p.GW = d->OW;
        p.ntaps = d->KH * d->KW;
        const bool phases = d->stride == 2 && d->C1 == 0;
        if (phases) {  // four parity sub-images of the input as strided views
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    InputView &v = p.view[a * 2 + b];
                    v.base = static_cast<const char *>(d->src0) + ((int64_t)a * d->W + b) * d->ld0 * esz;
                    v.H = (d->H - a + 1) / 2; v.W = (d->W - b + 1) / 2;
                    v.sx = 2 * d->ld0; v.sy = 2 * (int64_t)d->W * d->ld0; v.sn = (int64_t)d->H * d->W * d->ld0;
                }
            p.nviews = 4;
        } else {
            p.stride = d->stride;
            if (d->stride == 1) { p.kh = d->KH; p.kw = d->KW; p.pad_h = d->pad; p.pad_w = d->pad_w; }
        }
        for (int r = 0; r < d->KH; ++r)
            for (int s = 0; s < d->KW; ++s) {
                const int t = r * d->KW + s, qy = r - d->pad, qx = s - d->pad_w;
                if (phases) {
                    const int a = mod2(qy), b = mod2(qx);
                    p.tap_dy[t] = (int8_t)((qy - a) / 2); p.tap_dx[t] = (int8_t)((qx - b) / 2); p.tap_map[t] = (int8_t)(a * 2 + b);
                } else {
                    p.tap_dy[t] = (int8_t)qy; p.tap_dx[t] = (int8_t)qx; p.tap_map[t] = 0;
                }
            }
        p.K = p.ntaps * Cin; p.Kpad = col_len[0]; p.weight = d->weight;
    } else if (d->mode == HOIG_CONV_TRANSPOSED) {
        HOIG_REQUIRE(d->stride == 2 && d->OH == 2 * d->H && d->OW == 2 * d->W && d->C1 == 0,
                     "conv2d(transposed): only stride 2 with output = 2x input (k3 p1 op1 style) is supported");
        HOIG_REQUIRE(d->Cout % 16 == 0, "conv2d(transposed): Cout must be a multiple of 16");
        ConvParams &p = plan->launch[0];
        p = base;
        plan->n = 1;
        p.GH = d->H; p.GW = d->W;
        p.os = 2; p.ooy = 0; p.oox = 0;
        p.phase_cout = d->Cout;
        p.Cout = 4 * d->Cout;
        p.ntaps = 4;
        for (int t = 0; t < 4; ++t) { p.tap_dy[t] = (int8_t)(t >> 1); p.tap_dx[t] = (int8_t)(t & 1); p.tap_map[t] = 0; }
        p.K = 4 * Cin; p.Kpad = col_len[0]; p.weight = d->weight;
    } else {
        HOIG_REQUIRE(d->flow && d->src1 && d->C0 == d->C1 && d->OH == d->H && d->OW == d->W && d->KH == d->KW,
                     "conv2d(local_attn): needs flow, src1, C0 == C1, OH == H");
        ConvParams &p = plan->launch[0];
        p = base;
        plan->n = 1;
        p.GH = d->OH; p.GW = d->OW;
        p.ntaps = d->KH * d->KW;
        p.K = p.ntaps * Cin; p.Kpad = col_len[0]; p.weight = d->weight;
    }
    for (int i = 0; i < plan->n; ++i)
        plan->launch[i].tiles_per_image = ceil_div((int64_t)plan->launch[i].GH * plan->launch[i].GW, bm);
    return HOIG_OK;
}

}  // namespace hoig

extern "C" int hoig_conv_packed_dims(int mode, int Cout, int KH, int KW, int Cin, int stride, int pad, int *rows, int *cols)
{
    int r, c, off[4], len[4];
    hoig::packed_layout(mode, Cout, KH, KW, Cin, stride, pad, &r, &c, off, len);
    if (rows) *rows = r;
    if (cols) *cols = c;
    return HOIG_OK;
}
