/*
 * lv_measure.cu — K1..K4 fused: one h-evaluation of the measurement model on the GPU.
 *
 * Replaces, per input point, the chain (reference paths relative to the LIMO-Velo tree)
 *   Mapper::match                 src/Modules/Mapper.cpp:40-56       world transform
 *   KD_TREE::Nearest_Search       include/ikd-Tree/ikd_Tree/ikd_Tree.cpp:426-461   exact 5-NN
 *   Plane::Plane / estimate_plane src/Objects/Plane.cpp:19-55, src/Utils/Utils.cpp:32-66
 *   Match::Match                  src/Objects/Match.cpp:18-22
 *   Localizator::calculate_H      src/Modules/Localizator.cpp:29-57
 * and the reduction IKFoM performs on its output,
 *   HTH = h_x^T h_x, h_x^T h      esekfom.hpp:1723,1727
 * H (Nm x 12 fp64) is never materialised: every thread produces its row in registers, rows are
 * staged once in shared memory and folded into the 78 + 12 unique sums per block, in a fixed
 * order (deterministic).  One block = kMeasureThreads queries per tile, grid-stride over tiles.
 *
 * Bound: HBM/L2 gather latency: algorithmic traffic is 72 B per point (12 B query +
 * 5 x 12 B neighbours), no tensor-core-shaped work.
 */
#include "lv_internal.h"

namespace lv {

/* the 90 (a, b) products each block accumulates: 78 upper-triangle entries of HTH, then 12 of HTh
 * (b = 12 selects h) */
__constant__ uint8_t c_pair_a[90];
__constant__ uint8_t c_pair_b[90];
static bool g_pairs_ready = false;

static void init_pairs() {
    if (g_pairs_ready) return;
    uint8_t a[90], b[90];
    int e = 0;
    for (int i = 0; i < 12; ++i)
        for (int j = i; j < 12; ++j) { a[e] = (uint8_t)i; b[e] = (uint8_t)j; ++e; }
    for (int i = 0; i < 12; ++i) { a[e] = (uint8_t)i; b[e] = 12; ++e; }
    cudaMemcpyToSymbol(c_pair_a, a, sizeof(a));
    cudaMemcpyToSymbol(c_pair_b, b, sizeof(b));
    g_pairs_ready = true;
}

#define LV_ROW_STRIDE (kMeasureThreads + 1)   /* +1 double: 13 row-columns land in distinct banks */

__global__ void __launch_bounds__(kMeasureThreads) lv_measure_kernel(const MeasureArgs a) {
    if (a.ctrl->done) return;   /* update already finished (uniform over the grid) */

    __shared__ Frame s_frame;
    __shared__ double s_rows[13 * LV_ROW_STRIDE];

    {   /* the frame of the current iterate: written by the step kernel, broadcast via smem */
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&a.ctrl->frame);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_frame);
        for (int i = threadIdx.x; i < (int)(sizeof(Frame) / 4); i += kMeasureThreads) dst[i] = src[i];
    }
    __syncthreads();

    const int tid = threadIdx.x;
    double acc = 0.0;
    int count = 0;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int i = tile * kMeasureThreads + tid;
        bool chosen = false;
        double row[12], hval = 0.0;
        if (i < a.n) {
            const float px = a.xyz[3 * i], py = a.xyz[3 * i + 1], pz = a.xyz[3 * i + 2];
            float g[3];
            rt_apply(s_frame.lidar_to_world, px, py, pz, g);                  /* Mapper.cpp:51 */
            Top5 t;
            top5_init(t, a.max_d2);
            const bool finite = (fabsf(g[0]) < 1e9f) && (fabsf(g[1]) < 1e9f) && (fabsf(g[2]) < 1e9f);
            if (finite) knn5(a.map, g[0], g[1], g[2], a.max_d2, a.max_ring, t);
            float abcd[4] = {0.f, 0.f, 0.f, 0.f};
            float dist = 0.f;
            /* Plane.cpp:36-43: 5 neighbours and the farthest closer than MAX_DIST_PLANE */
            if (t.i4 >= 0 && (double)t.d4 < a.gate_d2) {
                float q[5][3];
                const float4 q0 = load_point(a.map.pts + t.i0), q1 = load_point(a.map.pts + t.i1),
                             q2 = load_point(a.map.pts + t.i2), q3 = load_point(a.map.pts + t.i3),
                             q4 = load_point(a.map.pts + t.i4);
                q[0][0] = q0.x; q[0][1] = q0.y; q[0][2] = q0.z;
                q[1][0] = q1.x; q[1][1] = q1.y; q[1][2] = q1.z;
                q[2][0] = q2.x; q[2][1] = q2.y; q[2][2] = q2.z;
                q[3][0] = q3.x; q[3][1] = q3.y; q[3][2] = q3.z;
                q[4][0] = q4.x; q[4][1] = q4.y; q[4][2] = q4.z;
                chosen = plane_fit(q, a.planes_threshold, abcd);               /* Plane.cpp:45-55 */
                if (chosen) {
                    dist = plane_dist(abcd, g);                                /* Match.cpp:21 */
                    jacobian_row(s_frame, g, abcd, dist, a.estimate_extrinsics != 0, row, &hval);
                } else {
                    abcd[0] = abcd[1] = abcd[2] = abcd[3] = 0.f;
                }
            }
            if (a.valid) a.valid[i] = chosen ? 1 : 0;
            if (a.g_world) { a.g_world[3 * i] = g[0]; a.g_world[3 * i + 1] = g[1]; a.g_world[3 * i + 2] = g[2]; }
            if (a.nn_idx || a.nn_sqd) {
                const int ids[5] = {t.i0, t.i1, t.i2, t.i3, t.i4};
                const float ds[5] = {t.d0, t.d1, t.d2, t.d3, t.d4};
                const bool full = t.i4 >= 0;
                for (int k = 0; k < 5; ++k) {
                    if (a.nn_idx)
                        a.nn_idx[5 * i + k] = full ? __float_as_int(load_point(a.map.pts + ids[k]).w) : -1;
                    if (a.nn_sqd) a.nn_sqd[5 * i + k] = full ? ds[k] : INFINITY;
                }
            }
            if (a.plane) { for (int k = 0; k < 4; ++k) a.plane[4 * i + k] = abcd[k]; }
            if (a.dist) a.dist[i] = dist;
            if (a.rows) {
                for (int k = 0; k < 12; ++k) a.rows[13 * (size_t)i + k] = chosen ? row[k] : 0.0;
                a.rows[13 * (size_t)i + 12] = chosen ? hval : 0.0;
            }
        }
        /* stage the row (zeros when rejected) and fold the tile into the block's 90 sums */
#pragma unroll
        for (int k = 0; k < 12; ++k) s_rows[k * LV_ROW_STRIDE + tid] = chosen ? row[k] : 0.0;
        s_rows[12 * LV_ROW_STRIDE + tid] = chosen ? hval : 0.0;
        count += __syncthreads_count(chosen ? 1 : 0);
        if (tid < 90) {
            const double* ra = s_rows + c_pair_a[tid] * LV_ROW_STRIDE;
            const double* rb = s_rows + c_pair_b[tid] * LV_ROW_STRIDE;
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 8
            for (int k = 0; k < kMeasureThreads; k += 4) {
                s0 += ra[k] * rb[k];
                s1 += ra[k + 1] * rb[k + 1];
                s2 += ra[k + 2] * rb[k + 2];
                s3 += ra[k + 3] * rb[k + 3];
            }
            acc += (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
    }
    double* out = a.partials + (size_t)blockIdx.x * kPartialStride;
    if (tid < 90) out[tid] = acc;
    if (tid == 90) out[90] = (double)count;
}

/* ---- fixed-order reduction of the per-block partials ---------------------------------------- */
/* Each of the 16 warps of the block owns blocks w, w+16, ...; lanes own elements lane, lane+32,
 * lane+64.  Then the 16 warp sums are added in warp order.  Deterministic for a given grid.    */
__device__ void reduce_partials_block(const double* partials, int n_partials, double* s_tmp /*16*96*/,
                                      double* HTH, double* HTh, int64_t* nm) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    double a0 = 0, a1 = 0, a2 = 0;
    for (int b = warp; b < n_partials; b += nwarps) {
        const double* p = partials + (size_t)b * kPartialStride;
        a0 += p[lane];
        a1 += p[lane + 32];
        a2 += p[lane + 64];
    }
    s_tmp[warp * 96 + lane] = a0;
    s_tmp[warp * 96 + lane + 32] = a1;
    s_tmp[warp * 96 + lane + 64] = a2;
    __syncthreads();
    if (threadIdx.x < 91) {
        double s = 0;
        for (int w = 0; w < nwarps; ++w) s += s_tmp[w * 96 + threadIdx.x];
        const int e = threadIdx.x;
        if (e < 78) {
            const int i = c_pair_a[e], j = c_pair_b[e];
            HTH[i * 12 + j] = s;
            HTH[j * 12 + i] = s;
        } else if (e < 90) {
            HTh[e - 78] = s;
        } else {
            *nm = (int64_t)(s + 0.5);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kStepThreads) lv_ieskf_step_kernel(UpdateCtrl* c, const IeskfParams prm,
                                                                     const double* partials, int n_partials) {
    if (c->done) return;
    __shared__ IeskfWork w;
    __shared__ double s_tmp[(kStepThreads / 32) * 96];
    reduce_partials_block(partials, n_partials, s_tmp, w.HTH, w.HTh, &w.n_matches);
    ExecBlock ex;
    ieskf_step(ex, prm, c, &w);
}

__global__ void __launch_bounds__(kStepThreads) lv_reduce_partials_kernel(const double* partials, int n_partials,
                                                                          double* out) {
    __shared__ double s_tmp[(kStepThreads / 32) * 96];
    __shared__ double HTH[144], HTh[12];
    __shared__ int64_t nm;
    reduce_partials_block(partials, n_partials, s_tmp, HTH, HTh, &nm);
    for (int i = threadIdx.x; i < 144; i += blockDim.x) out[i] = HTH[i];
    if (threadIdx.x < 12) out[144 + threadIdx.x] = HTh[threadIdx.x];
    if (threadIdx.x == 0) out[156] = (double)nm;
}

__global__ void __launch_bounds__(256) lv_ieskf_begin_kernel(UpdateCtrl* c) {
    ExecBlock ex;
    ieskf_begin(ex, c);
}

__global__ void lv_set_frame_kernel(UpdateCtrl* c) {
    if (threadIdx.x == 0) {
        make_frame(c->x, &c->frame);
        c->done = 0;
    }
}

__global__ void lv_l2_flush_kernel(uint4* buf, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        buf[i] = make_uint4((unsigned)i, 1u, 2u, 3u);
}

/* ---- launchers ------------------------------------------------------------------------------- */
int measure_grid(int n) {
    const int tiles = (n + kMeasureThreads - 1) / kMeasureThreads;
    const int cap = 148 * 4;   /* one B200: 148 SMs, 4 blocks of 128 threads resident per SM is ample */
    return tiles < cap ? (tiles > 0 ? tiles : 1) : cap;
}

cudaError_t launch_measure(const MeasureArgs& a, int grid, cudaStream_t st) {
    init_pairs();
    lv_measure_kernel<<<grid, kMeasureThreads, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_ieskf_begin(UpdateCtrl* c, cudaStream_t st) {
    lv_ieskf_begin_kernel<<<1, 256, 0, st>>>(c);
    return cudaGetLastError();
}
cudaError_t launch_ieskf_step(UpdateCtrl* c, const IeskfParams& prm, const double* partials, int n_partials,
                              cudaStream_t st) {
    init_pairs();
    lv_ieskf_step_kernel<<<1, kStepThreads, 0, st>>>(c, prm, partials, n_partials);
    return cudaGetLastError();
}
cudaError_t launch_reduce_partials(const double* partials, int n_partials, double* out, cudaStream_t st) {
    init_pairs();
    lv_reduce_partials_kernel<<<1, kStepThreads, 0, st>>>(partials, n_partials, out);
    return cudaGetLastError();
}
cudaError_t launch_set_frame(UpdateCtrl* c, cudaStream_t st) {
    lv_set_frame_kernel<<<1, 32, 0, st>>>(c);
    return cudaGetLastError();
}
cudaError_t launch_l2_flush(void* buf, size_t bytes, cudaStream_t st) {
    lv_l2_flush_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<uint4*>(buf), bytes / 16);
    return cudaGetLastError();
}

}  // namespace lv
