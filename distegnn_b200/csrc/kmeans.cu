// Lloyd iterations of the k-means node partitioner on the device.
//
// Reference: datasets/distribute_graphs.py:118-143, 188-198 — `KMeans(n_clusters=P, random_state=0, n_init="auto")
// .fit_predict(pos)` (sklearn, host) assigns every node of the big graph to one of P partitions.  The O(N·P·iterations)
// part runs here; the k-means++ seeding stays sklearn's own (`kmeans_plusplus`, which reproduces `random_state=0`) on the
// caller's side.  One call enqueues `iters` Lloyd iterations with sklearn's stopping rules evaluated ON THE DEVICE
// (`_kmeans_single_lloyd`): stop when no label changed (strict convergence) or when the squared centre shift falls to
// `tol` — then one more assignment pass so that labels match the final centres; iterations enqueued after convergence
// are no-ops.  state[0]: 0 running, 1 final assignment pending, 2 done; state[1]: iterations done; state[2]: labels
// changed in the last pass.  Cluster sums are accumulated in float64 (order-independent to ~1e-16).
#include "common.cuh"

namespace degnn {

constexpr int KM_MAXK = 64;
constexpr int KM_THREADS = 256;

struct KmArgs {
    int64_t N;
    int K;
    float tol;
    const float* pos;       // [N,3]
    float* centers;         // [K,3]
    int32_t* labels;        // [N] (in: previous labels, −1 initially)
    double* sums;           // [K,4] Σx, Σy, Σz, count   (zero on entry to every pass)
    int32_t* state;         // [4]
};

__global__ void __launch_bounds__(KM_THREADS) kmeans_assign_kernel(const KmArgs a) {
    __shared__ float sc[KM_MAXK * 3];
    __shared__ double ssum[KM_MAXK * 4];
    __shared__ int schanged;
    const int st = a.state[0];
    if (st == 2) return;
    const int tid = threadIdx.x, K = a.K;
    for (int i = tid; i < K * 3; i += KM_THREADS) sc[i] = a.centers[i];
    for (int i = tid; i < K * 4; i += KM_THREADS) ssum[i] = 0.0;
    if (tid == 0) schanged = 0;
    __syncthreads();
    int changed = 0;
    for (int64_t i = (int64_t)blockIdx.x * KM_THREADS + tid; i < a.N; i += (int64_t)gridDim.x * KM_THREADS) {
        const float x = __ldg(a.pos + i * 3), y = __ldg(a.pos + i * 3 + 1), z = __ldg(a.pos + i * 3 + 2);
        float best = INFINITY;
        int bk = 0;
        for (int k = 0; k < K; ++k) {
            const float dx = x - sc[3 * k], dy = y - sc[3 * k + 1], dz = z - sc[3 * k + 2];
            const float d = dx * dx + dy * dy + dz * dz;
            if (d < best) { best = d; bk = k; }          // first minimum wins, as argmin
        }
        if (a.labels[i] != bk) {
            ++changed;
            a.labels[i] = bk;
        }
        if (st == 0) {
            atomicAdd(ssum + 4 * bk, (double)x);
            atomicAdd(ssum + 4 * bk + 1, (double)y);
            atomicAdd(ssum + 4 * bk + 2, (double)z);
            atomicAdd(ssum + 4 * bk + 3, 1.0);
        }
    }
    if (changed) atomicAdd(&schanged, changed);
    __syncthreads();
    if (st == 0)
        for (int i = tid; i < K * 4; i += KM_THREADS)
            if (ssum[i] != 0.0) atomicAdd(a.sums + i, ssum[i]);
    if (tid == 0 && schanged) atomicAdd(a.state + 2, schanged);
}

__global__ void kmeans_update_kernel(const KmArgs a) {
    __shared__ float shift[KM_MAXK];
    const int st = a.state[0];
    if (st == 2) return;
    const int k = threadIdx.x;
    if (st == 1) {                                        // the final assignment has run
        if (k == 0) a.state[0] = 2;
        return;
    }
    float s = 0.f;
    if (k < a.K) {
        const double n = a.sums[4 * k + 3];
        if (n > 0.0) {                                    // an empty cluster keeps its centre
            const float cx = (float)(a.sums[4 * k] / n), cy = (float)(a.sums[4 * k + 1] / n), cz = (float)(a.sums[4 * k + 2] / n);
            const float dx = cx - a.centers[3 * k], dy = cy - a.centers[3 * k + 1], dz = cz - a.centers[3 * k + 2];
            s = dx * dx + dy * dy + dz * dz;
            a.centers[3 * k] = cx; a.centers[3 * k + 1] = cy; a.centers[3 * k + 2] = cz;
        }
        a.sums[4 * k] = a.sums[4 * k + 1] = a.sums[4 * k + 2] = a.sums[4 * k + 3] = 0.0;
    }
    if (k < KM_MAXK) shift[k] = s;
    __syncthreads();
    if (k == 0) {
        float tot = 0.f;
        for (int i = 0; i < a.K; ++i) tot += shift[i];
        a.state[1] += 1;
        if (a.state[2] == 0) a.state[0] = 2;              // strict convergence: labels already match the centres
        else if (tot <= a.tol) a.state[0] = 1;            // converged by tolerance: one more assignment pass
        a.state[2] = 0;
    }
}

}  // namespace degnn

extern "C" int distegnn_kmeans_lloyd(int64_t n_nodes, int n_clusters, const float* pos, float* centers, int32_t* labels,
                                     double* sums, int32_t* state, float tol, int iters, void* stream) {
    using namespace degnn;
    DEGNN_CHECK_ARG(n_nodes > 0 && pos && centers && labels && sums && state, "null pointer / bad size");
    DEGNN_CHECK_ARG(n_clusters >= 1 && n_clusters <= KM_MAXK, "n_clusters outside [1,64]");
    DEGNN_CHECK_ARG(iters >= 1 && tol >= 0.f, "bad iteration count / tolerance");
    KmArgs a;
    a.N = n_nodes; a.K = n_clusters; a.tol = tol; a.pos = pos; a.centers = centers; a.labels = labels; a.sums = sums;
    a.state = state;
    int64_t blocks = (n_nodes + KM_THREADS * 4 - 1) / (KM_THREADS * 4);
    if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
    for (int it = 0; it < iters; ++it) {
        kmeans_assign_kernel<<<(unsigned)blocks, KM_THREADS, 0, (cudaStream_t)stream>>>(a);
        kmeans_update_kernel<<<1, KM_MAXK, 0, (cudaStream_t)stream>>>(a);
    }
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
