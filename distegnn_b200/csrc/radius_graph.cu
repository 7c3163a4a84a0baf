// On-device radius graph: replaces the CPU `radius_graph(pos_i, r=radius, max_num_neighbors=N)` +
// `edge_attr = ‖Δx‖ duplicated` of the reference's partitioners (datasets/distribute_graphs.py:43-44, PyG /
// torch_cluster on the host) with a uniform-grid cell list on the GPU that emits the edges already grouped by
// destination row — i.e. the CSR the edge kernels want, without the radix sort of distegnn_build_csr.
//
// Two phases, because the edge count is only known after the first (the library never allocates):
//   distegnn_radius_count   per node: number of neighbours within r (same graph id, j != i unless `loop`)
//   (caller: exclusive prefix sum of the counts -> rowptr, allocates E = rowptr[N] entries)
//   distegnn_radius_fill    per node: writes col[rowptr[i] ..], row (= i repeated) and the edge length
// Cell list: the caller passes the nodes' cell keys SORTED (`order` = node ids in key order, `keys` = sorted keys,
// key = graph·ncell + (ix·ny + iy)·nz + iz, cell size >= r) — sorting is a library call on the caller's side (torch) —
// and a dense table `cell_start[B·ncell + 1]` (first position of every key in the sorted order).  A thread handles
// one node and scans the 27 neighbouring cells; nodes are processed in key order so that a warp touches neighbouring
// cells together (L1/L2 locality); positions are read as float4 (x4 layout, w ignored).
#include "common.cuh"

namespace degnn {

struct RadiusArgs {
    int64_t N;
    const float* x4;            // [N,4]
    const int32_t* batch;       // [N] graph id (may be null: single graph)
    const int32_t* order;       // [N] node ids sorted by cell key
    const int64_t* cell_start;  // [B*ncell + 1]
    float ox, oy, oz, inv_cell; // grid origin, 1 / cell size
    int nx, ny, nz;
    float r2;
    int loop;
    int32_t* deg;               // count phase: [N]
    const int64_t* rowptr;      // fill phase: [N+1]
    int32_t* row;               // fill phase: [E]
    int32_t* col;               // fill phase: [E]
    float* dist;                // fill phase: [E] or null
};

template <bool FILL>
__global__ void __launch_bounds__(256) radius_kernel(const RadiusArgs a) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.N) return;
    const int i = __ldg(a.order + k);
    const float4 p = ldg4(a.x4 + (size_t)i * 4);
    const int g = a.batch ? __ldg(a.batch + i) : 0;
    const int ix = min(max((int)((p.x - a.ox) * a.inv_cell), 0), a.nx - 1);
    const int iy = min(max((int)((p.y - a.oy) * a.inv_cell), 0), a.ny - 1);
    const int iz = min(max((int)((p.z - a.oz) * a.inv_cell), 0), a.nz - 1);
    const int64_t gbase = (int64_t)g * a.nx * a.ny * a.nz;
    int cnt = 0;
    int64_t w = FILL ? __ldg(a.rowptr + i) : 0;
    for (int dx = -1; dx <= 1; ++dx) {
        const int cx = ix + dx;
        if (cx < 0 || cx >= a.nx) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            const int cy = iy + dy;
            if (cy < 0 || cy >= a.ny) continue;
            // the three z-neighbours are consecutive keys: one contiguous range of the sorted order
            const int z0 = max(iz - 1, 0), z1 = min(iz + 1, a.nz - 1);
            const int64_t key0 = gbase + ((int64_t)cx * a.ny + cy) * a.nz + z0;
            const int64_t s = __ldg(a.cell_start + key0), e = __ldg(a.cell_start + key0 + (z1 - z0) + 1);
            for (int64_t q = s; q < e; ++q) {
                const int j = __ldg(a.order + q);
                if (j == i && !a.loop) continue;
                const float4 pj = ldg4(a.x4 + (size_t)j * 4);
                const float ddx = p.x - pj.x, ddy = p.y - pj.y, ddz = p.z - pj.z;
                const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                if (d2 < a.r2) {          // strict, as torch_cluster
                    if (FILL) {
                        a.row[w] = i;
                        a.col[w] = j;
                        if (a.dist) a.dist[w] = sqrtf(d2);
                        ++w;
                    } else {
                        ++cnt;
                    }
                }
            }
        }
    }
    if (!FILL) a.deg[i] = cnt;
}

static int fill_args(RadiusArgs& a, int64_t n_nodes, const float* x4, const int32_t* batch32, const int32_t* order,
                     const int64_t* cell_start, const float* origin_host, float cell_size, const int32_t* dims_host, float radius,
                     int loop) {
    a.N = n_nodes; a.x4 = x4; a.batch = batch32; a.order = order; a.cell_start = cell_start;
    a.ox = origin_host[0]; a.oy = origin_host[1]; a.oz = origin_host[2];
    a.inv_cell = 1.0f / cell_size;
    a.nx = dims_host[0]; a.ny = dims_host[1]; a.nz = dims_host[2];
    a.r2 = radius * radius;
    a.loop = loop;
    a.deg = nullptr; a.rowptr = nullptr; a.row = nullptr; a.col = nullptr; a.dist = nullptr;
    return 0;
}

}  // namespace degnn

extern "C" int distegnn_radius_count(int64_t n_nodes, const float* x4, const int32_t* batch32, const int32_t* order,
                                     const int64_t* cell_start, const float* origin_host, float cell_size,
                                     const int32_t* dims_host, float radius, int loop, int32_t* deg, void* stream) {
    using namespace degnn;
    if (n_nodes == 0) return DISTEGNN_OK;
    DEGNN_CHECK_ARG(n_nodes > 0 && x4 && order && cell_start && origin_host && dims_host && deg, "null pointer / bad size");
    DEGNN_CHECK_ARG(cell_size >= radius && radius > 0.f, "cell size must be >= radius > 0");
    DEGNN_CHECK_ARG(dims_host[0] > 0 && dims_host[1] > 0 && dims_host[2] > 0, "bad grid dims");
    RadiusArgs a;
    fill_args(a, n_nodes, x4, batch32, order, cell_start, origin_host, cell_size, dims_host, radius, loop);
    a.deg = deg;
    radius_kernel<false><<<(unsigned)((n_nodes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}

extern "C" int distegnn_radius_fill(int64_t n_nodes, const float* x4, const int32_t* batch32, const int32_t* order,
                                    const int64_t* cell_start, const float* origin_host, float cell_size,
                                    const int32_t* dims_host, float radius, int loop, const int64_t* rowptr, int32_t* row,
                                    int32_t* col, float* dist, void* stream) {
    using namespace degnn;
    if (n_nodes == 0) return DISTEGNN_OK;
    DEGNN_CHECK_ARG(n_nodes > 0 && x4 && order && cell_start && origin_host && dims_host && rowptr && row && col,
                    "null pointer / bad size");
    DEGNN_CHECK_ARG(cell_size >= radius && radius > 0.f, "cell size must be >= radius > 0");
    RadiusArgs a;
    fill_args(a, n_nodes, x4, batch32, order, cell_start, origin_host, cell_size, dims_host, radius, loop);
    a.rowptr = rowptr; a.row = row; a.col = col; a.dist = dist;
    radius_kernel<true><<<(unsigned)((n_nodes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    DEGNN_CHECK_LAUNCH();
    return DISTEGNN_OK;
}
