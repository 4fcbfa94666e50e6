// api.cu -- error plumbing and dtype dispatch of the C ABI (include/hoig_b200.h).
#include <stdarg.h>

#include "conv_common.cuh"

namespace hoig {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what)
{
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return HOIG_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return HOIG_ERR_CUDA;
}

static int g_sm_count[kMaxDevices];
static bool g_slot_used[kMaxDevices][SLOT_COUNT];

static int current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
    return dev;
}

int device_sm_count()
{
    const int dev = current_device();
    if (!g_sm_count[dev]) cudaDeviceGetAttribute(&g_sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    return g_sm_count[dev];
}

bool first_use_on_device(int slot)
{
    bool &used = g_slot_used[current_device()][slot];
    const bool first = !used;
    used = true;
    return first;
}

int conv2d_simt(const hoigConvDesc *d, cudaStream_t stream);
int conv2d_umma(const hoigConvDesc *d, cudaStream_t stream);
int conv2d_halo(int dtype, int KH, int KW, int Cout, const hoigHaloConvSeg *segs, int nsegs, cudaStream_t stream);

}  // namespace hoig

extern "C" const char *hoig_version(void) { return "hoig_b200 0.1 (sm_100a)"; }
extern "C" const char *hoig_last_error(void) { return hoig::g_err; }

extern "C" int hoig_check_device(void)
{
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        hoig::set_error("no usable CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        return HOIG_ERR_CUDA;
    }
    if (major != 10) {
        hoig::set_error("device compute capability %d.x is not sm_100", major);
        return HOIG_ERR_ARCH;
    }
    return HOIG_OK;
}

extern "C" int hoig_conv2d(const hoigConvDesc *desc, hoigStream_t stream)
{
    if (!desc) { hoig::set_error("conv2d: null descriptor"); return HOIG_ERR_INVALID; }
    if (desc->dtype == HOIG_BF16 || desc->dtype == HOIG_F16) return hoig::conv2d_umma(desc, hoig::as_stream(stream));
    return hoig::conv2d_simt(desc, hoig::as_stream(stream));
}

// Test hook: the SIMT kernel with bf16 storage, to cross-check the tensor-core kernel.
extern "C" int hoig_conv2d_simt(const hoigConvDesc *desc, hoigStream_t stream)
{
    if (!desc) { hoig::set_error("conv2d: null descriptor"); return HOIG_ERR_INVALID; }
    return hoig::conv2d_simt(desc, hoig::as_stream(stream));
}

extern "C" int hoig_conv2d_halo(int dtype, int KH, int KW, int Cout, const hoigHaloConvSeg *segs, int nsegs, hoigStream_t stream)
{
    return hoig::conv2d_halo(dtype, KH, KW, Cout, segs, nsegs, hoig::as_stream(stream));
}
