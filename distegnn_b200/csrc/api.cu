// Host-side plumbing of the C ABI: error string, parameter-block layout, device query.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <set>
#include <utility>

#include <cuda.h>

#include "common.cuh"

namespace degnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_dims(int A, int C, int Na) {
    if (A < 0 || A > DISTEGNN_MAX_EDGE_ATTR) {
        set_error("edge_attr_nf=%d outside [0,%d]", A, DISTEGNN_MAX_EDGE_ATTR);
        return DISTEGNN_EINVAL;
    }
    if (C < 1 || C > DISTEGNN_MAX_CHANNELS) {
        set_error("virtual_channels=%d outside [1,%d]", C, DISTEGNN_MAX_CHANNELS);
        return DISTEGNN_EINVAL;
    }
    if (Na < 0 || Na > DISTEGNN_MAX_NODE_ATTR) {
        set_error("node_attr_nf=%d outside [0,%d]", Na, DISTEGNN_MAX_NODE_ATTR);
        return DISTEGNN_EINVAL;
    }
    return DISTEGNN_OK;
}

Layout make_layout(int A, int C, int Na) {
    Layout L;
    int64_t o = 0;
    auto put = [&](int field, int64_t n) {
        L.off[field] = o;
        o += (n + 3) / 4 * 4;
    };
    const int64_t HH = (int64_t)H * H;
    put(DISTEGNN_P_E_W1A, HH);
    put(DISTEGNN_P_E_W1B, HH);
    put(DISTEGNN_P_E_W1R, H);
    put(DISTEGNN_P_E_W1E, (int64_t)A * H);
    put(DISTEGNN_P_E_B1, H);
    put(DISTEGNN_P_E_W2, HH);
    put(DISTEGNN_P_E_B2, H);
    put(DISTEGNN_P_E_WC, HH);
    put(DISTEGNN_P_E_BC, H);
    put(DISTEGNN_P_E_W3, H);
    put(DISTEGNN_P_V_W1H, HH);
    put(DISTEGNN_P_V_W1V, HH);
    put(DISTEGNN_P_V_W1R, H);
    put(DISTEGNN_P_V_W1M, (int64_t)C * H);
    put(DISTEGNN_P_V_B1, H);
    put(DISTEGNN_P_V_W2, HH);
    put(DISTEGNN_P_V_B2, H);
    put(DISTEGNN_P_V_WXV, HH);
    put(DISTEGNN_P_V_BXV, H);
    put(DISTEGNN_P_V_W3XV, H);
    put(DISTEGNN_P_V_WX, HH);
    put(DISTEGNN_P_V_BX, H);
    put(DISTEGNN_P_V_W3X, H);
    put(DISTEGNN_P_L_W, HH);
    put(DISTEGNN_P_L_B, H);
    put(DISTEGNN_P_L_W3, H);
    put(DISTEGNN_P_L_B3, 4);
    put(DISTEGNN_P_N_W1, (int64_t)(3 * H + Na) * H);
    put(DISTEGNN_P_N_B1, H);
    put(DISTEGNN_P_N_W2, HH);
    put(DISTEGNN_P_N_B2, H);
    put(DISTEGNN_P_M_W1, 2 * HH);
    put(DISTEGNN_P_M_B1, H);
    put(DISTEGNN_P_M_W2, HH);
    put(DISTEGNN_P_M_B2, H);
    L.total = o;
    return L;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

void ensure_dynamic_smem(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<int, const void*>> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({dev, kernel})) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    done.insert({dev, kernel});
}

// Tensor map over a row-major fp32 matrix [n_rows][row_floats] for TMA row gathers: box = {box_floats, box_rows}.  A box
// wider than the row (box_floats > row_floats) is legal — the out-of-bounds tail is zero-filled — and gives the rows a
// padded pitch of box_floats in shared memory.  cuTensorMapEncodeTiled is a host-only driver call (no stream, no sync);
// its entry point is resolved through the runtime, so the library has no link-time dependency on libcuda.
int make_rows_tmap(void* out_map, const float* base, int64_t n_rows, int row_floats, int box_floats, int box_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!fn) {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
            if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
                (void)cudaGetLastError();
                set_error("cuTensorMapEncodeTiled is not available from the driver (%s)", cudaGetErrorString(e));
                return DISTEGNN_ECUDA;
            }
            fn = (EncodeFn)p;
        }
    }
    const cuuint64_t dims[2] = {(cuuint64_t)row_floats, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)row_floats * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_floats, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn((CUtensorMap*)out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (base %p, %lld rows x %d floats, box %d x %d)", (int)r,
                  (const void*)base, (long long)n_rows, row_floats, box_floats, box_rows);
        return DISTEGNN_ECUDA;
    }
    return DISTEGNN_OK;
}

}  // namespace degnn

extern "C" {

int distegnn_abi_version(void) { return DISTEGNN_ABI_VERSION; }

const char* distegnn_last_error(void) { return degnn::g_err; }

int distegnn_param_layout(int A, int C, int Na, int64_t* offsets_host, int64_t* total_floats_host) {
    if (int rc = degnn::check_dims(A, C, Na)) return rc;
    DEGNN_CHECK_ARG(offsets_host && total_floats_host, "null output pointer");
    degnn::Layout L = degnn::make_layout(A, C, Na);
    memcpy(offsets_host, L.off, sizeof(L.off));
    *total_floats_host = L.total;
    return DISTEGNN_OK;
}

}  // extern "C"
