// Host-side plumbing of the C ABI: error string, parameter-block layout, device query.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <set>
#include <utility>

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
