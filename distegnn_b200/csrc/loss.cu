// Loss side of the training step: node-count weighted MSE + MMD regulariser on the virtual coordinates,
// with the per-step scalar collectives folded into ONE packed SUM all-reduce.
//
// Reference: utils/train.py:98-147 — `loss(loc_pred, loc_target)` (MSE), two scalar all-reduces (total node count :104,
// logged loss :109) and an all_gather consistency check of `loc_mean` (:52-61) per step, then a Python loop over the
// graphs of the batch with `randperm` sampling, two `cdist` kernels (:11-14) per graph and ~20 small ATen launches each.
//
// Here (fp32; sums of at most a few thousand positive terms):
//   distegnn_loss_partials   one launch: Σ(pred−target)² over the rank's nodes folded straight into the packed vector
//                            [n_r, n_r·MSE_r, loc_mean of this rank in its slot], and per graph l_vv = Σ k(V_c,V_c'),
//                            l_rv = Σ k(R_s,V_c), k(x,y) = exp(−‖x−y‖/(2σ²)) (distance NOT squared, :12-13), together with
//                            the un-weighted gradient of (l_vv/B/C² − 2·l_rv/B/S/C) w.r.t. the virtual coordinates
//   (caller)                 ONE all-reduce (SUM) of the packed vector: total node count, logged loss and every rank's
//                            loc_mean (each rank fills only its own slot) arrive together
//   distegnn_loss_finalize   one launch: coef = world·n_r/Σn (:110); loss = coef·(MSE + weight·MMD) / accumulation_steps;
//                            the gradients d loss/d pred [N,3] and d loss/d Xv [B,3,C] (the forward of a fused loss already
//                            knows them); logged loss; max deviation of the ranks' loc_mean from rank 0's
#include "common.cuh"

namespace degnn {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_NODES_PER_CTA = 2048;

struct LossArgs {
    int64_t N;
    int B, C, S, world, rank;
    float sigma, weight, inv_accum;
    const float* pred;        // [N,3]
    const float* target;      // [N,3]
    const float* Xv;          // [B,3,C]
    const float* loc_mean;    // [B,3] or null
    const int64_t* graph_ptr; // [B+1] first node of every graph (data_batch is sorted)
    const int32_t* samples;   // [B,S] local node indices drawn for the MMD (−1 = none), train.py:128-129
    float* acc;               // [3]: Σ_b l_vv, Σ_b l_rv, n_r·MSE_r      (zeroed by the caller)
    float* packed;            // [2 + world·3B]                          (zeroed by the caller)
    float* gV_raw;            // [B,3,C] gradient of (l_vv/B/C² − 2 l_rv/B/S/C) w.r.t. Xv
    float* g_pred;            // [N,3]
    float* g_Xv;              // [B,3,C]
    float* out;               // [4]: loss, logged loss, MMD term, max |loc_mean_r − loc_mean_0|
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x < LOSS_THREADS / 32) s = sh[threadIdx.x];
    if (w == 0) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    }
    return s;   // valid in thread 0
}

// blocks [0, node_blocks): squared error; blocks [node_blocks, node_blocks + B): MMD of graph b
__global__ void __launch_bounds__(LOSS_THREADS) loss_partials_kernel(const LossArgs a, int node_blocks) {
    __shared__ float sh[LOSS_THREADS / 32];
    __shared__ float sV[3 * DISTEGNN_MAX_CHANNELS];
    __shared__ float sG[3 * DISTEGNN_MAX_CHANNELS];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < node_blocks) {
        const int64_t e0 = (int64_t)blockIdx.x * LOSS_NODES_PER_CTA * 3;
        const int64_t e1 = min(e0 + (int64_t)LOSS_NODES_PER_CTA * 3, a.N * 3);
        float s = 0.f;
        for (int64_t i = e0 + tid; i < e1; i += LOSS_THREADS) {
            const float d = __ldg(a.pred + i) - __ldg(a.target + i);
            s = fmaf(d, d, s);
        }
        s = block_sum(s, sh);
        if (tid == 0) {
            // n_r · MSE_r = n_r · sse / (3 n_r) = sse / 3: the rank's term of the logged loss before the division by Σn
            atomicAdd(a.packed + 1, s * (1.0f / 3.0f));
            atomicAdd(a.acc + 2, s * (1.0f / 3.0f));
            if (blockIdx.x == 0) {
                a.packed[0] = (float)a.N;
                if (a.loc_mean)
                    for (int i = 0; i < 3 * a.B; ++i) a.packed[2 + (size_t)a.rank * 3 * a.B + i] = a.loc_mean[i];
            }
        }
        return;
    }
    const int b = blockIdx.x - node_blocks, C = a.C, S = a.S;
    const float inv2s2 = 1.0f / (2.0f * a.sigma * a.sigma);
    if (tid < 3 * C) {
        sV[tid] = a.Xv[(size_t)b * 3 * C + tid];     // [3][C]
        sG[tid] = 0.f;
    }
    __syncthreads();
    const float wvv = 1.0f / ((float)a.B * C * C), wrv = 2.0f / ((float)a.B * S * C);
    float lvv = 0.f, lrv = 0.f;
    // virtual-virtual: ordered pairs (c, c'); the gradient w.r.t. V_c collects both orders: 2·∂k(V_c,V_c')/∂V_c
    for (int p = tid; p < C * C; p += LOSS_THREADS) {
        const int c = p / C, d = p - c * C;
        const float dx = sV[c] - sV[d], dy = sV[C + c] - sV[C + d], dz = sV[2 * C + c] - sV[2 * C + d];
        const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
        const float k = expf(-dist * inv2s2);
        lvv += k;
        if (dist > 0.f) {                            // cdist's backward is 0 at coincident points
            const float g = -2.0f * wvv * k * inv2s2 / dist;
            atomicAdd(sG + c, g * dx);
            atomicAdd(sG + C + c, g * dy);
            atomicAdd(sG + 2 * C + c, g * dz);
        }
    }
    // sampled real nodes vs virtual: pairs (s, c)
    const int64_t n0 = a.graph_ptr[b];
    for (int p = tid; p < S * C; p += LOSS_THREADS) {
        const int s = p / C, c = p - s * C;
        const int li = a.samples[(size_t)b * S + s];
        if (li < 0) continue;
        const float* r = a.target + (size_t)(n0 + li) * 3;
        const float dx = sV[c] - __ldg(r), dy = sV[C + c] - __ldg(r + 1), dz = sV[2 * C + c] - __ldg(r + 2);
        const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
        const float k = expf(-dist * inv2s2);
        lrv += k;
        if (dist > 0.f) {
            const float g = wrv * k * inv2s2 / dist;     // −wrv · ∂k/∂V_c
            atomicAdd(sG + c, g * dx);
            atomicAdd(sG + C + c, g * dy);
            atomicAdd(sG + 2 * C + c, g * dz);
        }
    }
    lvv = block_sum(lvv, sh);
    lrv = block_sum(lrv, sh);
    __syncthreads();
    if (tid == 0) {
        atomicAdd(a.acc + 0, lvv);
        atomicAdd(a.acc + 1, lrv);
    }
    if (tid < 3 * C) a.gV_raw[(size_t)b * 3 * C + tid] = sG[tid];
}

// after the all-reduce of `packed`: scalars (block 0) and the gradients (all blocks)
__global__ void __launch_bounds__(LOSS_THREADS) loss_finalize_kernel(const LossArgs a) {
    const int tid = threadIdx.x, C = a.C;
    const float n_r = (float)a.N, n_tot = a.packed[0];
    const float share = n_r / n_tot;                                    // node_cnt / total_node_cnt, train.py:105
    const float coef = (float)a.world * share * a.inv_accum;            // :110 (DDP averages, the reference wants the sum), :150
    const float mse = n_r > 0.f ? a.acc[2] / n_r : 0.f;
    const float mmd = a.acc[0] / ((float)a.B * C * C) - 2.0f * a.acc[1] / ((float)a.B * a.S * C);     // :142-145
    if (blockIdx.x == 0) {
        if (tid == 0) {
            a.out[0] = coef * (mse + a.weight * mmd);
            a.out[1] = a.packed[1] / n_tot;                             // Σ_r n_r/Σn · MSE_r  (:106-108)
            a.out[2] = mmd;
            float dev = 0.f;
            if (a.loc_mean)
                for (int r = 1; r < a.world; ++r)
                    for (int i = 0; i < 3 * a.B; ++i)
                        dev = fmaxf(dev, fabsf(a.packed[2 + (size_t)r * 3 * a.B + i] - a.packed[2 + i]));
            a.out[3] = dev;
        }
        const float cV = coef * a.weight;
        for (int i = tid; i < a.B * 3 * C; i += LOSS_THREADS) a.g_Xv[i] = cV * a.gV_raw[i];
    }
    const float cp = n_r > 0.f ? coef * 2.0f / (3.0f * n_r) : 0.f;      // d MSE / d pred = 2 (pred − target) / (3 n_r)
    const int64_t e0 = (int64_t)blockIdx.x * LOSS_NODES_PER_CTA * 3;
    const int64_t e1 = min(e0 + (int64_t)LOSS_NODES_PER_CTA * 3, a.N * 3);
    for (int64_t i = e0 + tid; i < e1; i += LOSS_THREADS) a.g_pred[i] = cp * (__ldg(a.pred + i) - __ldg(a.target + i));
}

}  // namespace degnn

static int loss_fill(degnn::LossArgs& a, int64_t n_nodes, int n_graphs, int C, int S, int world, int rank, float sigma,
                     float weight, int accumulation_steps, const float* pred, const float* target, const float* Xv,
                     const float* loc_mean, const int64_t* graph_ptr, const int32_t* samples, float* acc, float* packed,
                     float* gV_raw) {
    using namespace degnn;
    DEGNN_CHECK_ARG(n_nodes >= 0 && n_graphs > 0, "bad size");
    DEGNN_CHECK_ARG(C >= 1 && C <= DISTEGNN_MAX_CHANNELS, "virtual_channels out of range");
    DEGNN_CHECK_ARG(S >= 1 && world >= 1 && rank >= 0 && rank < world && accumulation_steps >= 1, "bad argument");
    DEGNN_CHECK_ARG(sigma > 0.f, "sigma must be positive");
    DEGNN_CHECK_ARG((n_nodes == 0 || (pred && target)) && Xv && graph_ptr && samples && acc && packed && gV_raw,
                    "null pointer");
    a.N = n_nodes; a.B = n_graphs; a.C = C; a.S = S; a.world = world; a.rank = rank;
    a.sigma = sigma; a.weight = weight; a.inv_accum = 1.0f / (float)accumulation_steps;
    a.pred = pred; a.target = target; a.Xv = Xv; a.loc_mean = loc_mean; a.graph_ptr = graph_ptr; a.samples = samples;
    a.acc = acc; a.packed = packed; a.gV_raw = gV_raw; a.g_pred = nullptr; a.g_Xv = nullptr; a.out = nullptr;
    return DISTEGNN_OK;
}

extern "C" int distegnn_loss_packed_floats(int n_graphs, int world) { return 2 + world * 3 * n_graphs; }

extern "C" int distegnn_loss_partials(int64_t n_nodes, int n_graphs, int C, int S, int world, int rank, float sigma,
                                      const float* pred, const float* target, const float* Xv, const float* loc_mean,
                                      const int64_t* graph_ptr, const int32_t* samples, float* acc, float* packed,
                                      float* gV_raw, void* stream) {
    using namespace degnn;
    LossArgs a;
    if (int rc = loss_fill(a, n_nodes, n_graphs, C, S, world, rank, sigma, 0.f, 1, pred, target, Xv, loc_mean, graph_ptr,
                           samples, acc, packed, gV_raw))
        return rc;
    int node_blocks = (int)((n_nodes + LOSS_NODES_PER_CTA - 1) / LOSS_NODES_PER_CTA);
    if (node_blocks < 1) node_blocks = 1;                               // block 0 also writes n_r and the loc_mean slot
    loss_partials_kernel<<<(unsigned)(node_blocks + n_graphs), LOSS_THREADS, 0, (cudaStream_t)stream>>>(a, node_blocks);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}

extern "C" int distegnn_loss_finalize(int64_t n_nodes, int n_graphs, int C, int S, int world, int rank, float sigma,
                                      float weight, int accumulation_steps, const float* pred, const float* target,
                                      const float* loc_mean, const float* acc, const float* packed, const float* gV_raw,
                                      float* g_pred, float* g_Xv, float* out, void* stream) {
    using namespace degnn;
    LossArgs a;
    static const int64_t dummy_ptr = 0;
    static const int32_t dummy_smp = 0;
    if (int rc = loss_fill(a, n_nodes, n_graphs, C, S, world, rank, sigma, weight, accumulation_steps, pred, target,
                           gV_raw /*unused Xv slot*/, loc_mean, &dummy_ptr, &dummy_smp, const_cast<float*>(acc),
                           const_cast<float*>(packed), const_cast<float*>(gV_raw)))
        return rc;
    DEGNN_CHECK_ARG((n_nodes == 0 || g_pred) && g_Xv && out, "null output pointer");
    a.g_pred = g_pred; a.g_Xv = g_Xv; a.out = out;
    int node_blocks = (int)((n_nodes + LOSS_NODES_PER_CTA - 1) / LOSS_NODES_PER_CTA);
    if (node_blocks < 1) node_blocks = 1;
    loss_finalize_kernel<<<(unsigned)node_blocks, LOSS_THREADS, 0, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
