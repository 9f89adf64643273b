/*
 * lv_map_add.cu — Mapper::add on an existing map, on the device.
 *
 * Replaces KD_TREE::Add_Points(points, downsample) (include/ikd-Tree/ikd_Tree/ikd_Tree.cpp:478-573)
 * as called from Mapper::add_points (src/Modules/Mapper.cpp:73-76).  The reference processes the
 * new points one by one: box = the 0.2 m voxel of the point (:491-499), storage = map points in
 * the box, keep the candidate closest to the voxel centre (:502-512) and, if the box held more
 * than one point or the winner is the new point, replace the box content by the winner (:515-521).
 * Folded over a whole batch this is order-independent up to exact ties: a voxel touched by at
 * least one new point ends up holding the single point of (old content + new points) closest to
 * its centre; untouched voxels keep everything.  That is what is computed here, for all voxels at
 * once: fine-voxel key -> stable radix sort -> one thread per voxel run picks the survivor(s) ->
 * exclusive scan -> compaction into the second map buffer.
 *
 * Tie rule (measure zero on float data, kept for determinism): a new point beats an old one at
 * equal distance (strict '<' at :507), the later of two equal new points wins.
 * Voxel membership of OLD points follows the half-open fp32 box test of Search_by_range (:1262), see
 * fine_coord_old.
 */
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/limovelo_b200.h"
#include "lv_host.h"
#include "lv_internal.h"

namespace lv {

LV_HD int fine_coord(float v, float ds) { return (int)floorf(fdiv(v, ds)); }   /* ikd_Tree.cpp:493 */
LV_HD uint64_t fine_key(int ix, int iy, int iz) {   /* any injective key of the downsample voxel */
    return ((uint64_t)(uint32_t)(iz + LV_KEY_BIAS) << 42) | ((uint64_t)(uint32_t)(iy + LV_KEY_BIAS) << 21) |
           (uint64_t)(uint32_t)(ix + LV_KEY_BIAS);
}

/* Voxel of an OLD map point as Search_by_range sees it (ikd_Tree.cpp:1262): the box of voxel k is
 * [fl(k * ds), fl(k * ds) + ds) in fp32, so a point within an ulp of a face can belong to the voxel next
 * to floor(q / ds). */
LV_HD int fine_coord_old(float v, float ds) {
    int k = fine_coord(v, ds);
    const float lo = fmul((float)k, ds);
    if (v < lo) --k;
    else if (!(v < fadd(lo, ds))) ++k;
    return k;
}

__global__ void __launch_bounds__(256) lv_add_keys_kernel(const float* __restrict__ xyz, int64_t total, int64_t n_old, float ds,
                                                           uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    keys[i] = i < n_old ? fine_key(fine_coord_old(x, ds), fine_coord_old(y, ds), fine_coord_old(z, ds))
                        : fine_key(fine_coord(x, ds), fine_coord(y, ds), fine_coord(z, ds));   /* ikd_Tree.cpp:493-498 */
    vals[i] = (uint32_t)i;
}

/* distance of p to the centre of its voxel, evaluated like ikd_Tree.cpp:493-503 */
__device__ __forceinline__ float centre_dist(float x, float y, float z, float ds) {
    float c[3];
    const float p[3] = {x, y, z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float bmin = fmul(floorf(fdiv(p[a], ds)), ds);
        const float bmax = fadd(bmin, ds);
        c[a] = (float)((double)bmin + (double)fsub(bmax, bmin) / 2.0);
    }
    return sq_dist(x, y, z, c[0], c[1], c[2]);
}

/* one thread per voxel run: mark the survivors */
__global__ void __launch_bounds__(256) lv_add_select_kernel(const float* __restrict__ xyz, int64_t total, int64_t n_old,
                                                             float ds, const uint64_t* __restrict__ keys_sorted,
                                                             const uint32_t* __restrict__ vals_sorted,
                                                             uint32_t* __restrict__ keep) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    const uint64_t key = keys_sorted[j];
    if (j > 0 && keys_sorted[j - 1] == key) return;
    int64_t end = j + 1;
    while (end < total && keys_sorted[end] == key) ++end;
    /* stable sort: old points (index < n_old) come first, new points follow in input order */
    const bool touched = vals_sorted[end - 1] >= (uint32_t)n_old;
    if (!touched) {
        for (int64_t k = j; k < end; ++k) keep[k] = 1u;
        return;
    }
    int64_t best = -1;
    float best_d = INFINITY;
    bool best_new = false;
    for (int64_t k = j; k < end; ++k) {
        const uint32_t src = vals_sorted[k];
        const bool is_new = src >= (uint32_t)n_old;
        const float d = centre_dist(xyz[3 * (size_t)src], xyz[3 * (size_t)src + 1], xyz[3 * (size_t)src + 2], ds);
        bool take;
        if (best < 0) take = true;
        else if (is_new) take = !(best_d < d);          /* new point replaces unless the kept one is strictly closer */
        else take = !best_new && d < best_d;            /* among old points: first strict minimum */
        if (take) { best = k; best_d = d; best_new = is_new; }
        keep[k] = 0u;
    }
    keep[best] = 1u;
}

__global__ void __launch_bounds__(256) lv_add_compact_kernel(const float* __restrict__ xyz, int64_t total,
                                                              const uint32_t* __restrict__ vals_sorted,
                                                              const uint32_t* __restrict__ keep,
                                                              const uint32_t* __restrict__ pos, float* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total || !keep[j]) return;
    const size_t src = vals_sorted[j], dst = pos[j];
    out[3 * dst] = xyz[3 * src];
    out[3 * dst + 1] = xyz[3 * src + 1];
    out[3 * dst + 2] = xyz[3 * src + 2];
}

}  // namespace lv

using namespace lv;

size_t lvh_map_add_tmp_bytes(int64_t cap) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)cap, 0, 63);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)cap);
    return a > b ? a : b;
}

int lvh_map_add_points(MapBuffers& b, const float* xyz_host, int64_t n, int downsample, float ds, cudaStream_t st,
                       int* launches) {
    *launches = 0;
    const int64_t n_old = b.n, total = b.n + n;
    if (total > b.cap) return LV_ERR_CAPACITY;
    if (cudaMemcpyAsync(b.xyz + 3 * n_old, xyz_host, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return LV_ERR_CUDA;
    if (!downsample) { b.n = total; return LV_OK; }                      /* ikd_Tree.cpp:549-571 */
    const unsigned blocks = (unsigned)((total + 255) / 256);
    /* scratch: keys/vals (sort input), keys_sorted/vals_sorted (sort output), b.pts reused as two
     * uint32 arrays (keep flags, scan output) — the search structure is rebuilt right after */
    uint32_t* keep = reinterpret_cast<uint32_t*>(b.pts);
    uint32_t* pos = keep + b.cap;
    lv_add_keys_kernel<<<blocks, 256, 0, st>>>(b.xyz, total, n_old, ds, b.keys, b.vals);
    size_t tmp = b.sort_tmp_bytes;
    if (cub::DeviceRadixSort::SortPairs(b.sort_tmp, tmp, b.keys, b.keys_sorted, b.vals, b.vals_sorted, (int)total, 0, 63, st) != cudaSuccess)
        return LV_ERR_CUDA;
    lv_add_select_kernel<<<blocks, 256, 0, st>>>(b.xyz, total, n_old, ds, b.keys_sorted, b.vals_sorted, keep);
    tmp = b.sort_tmp_bytes;
    if (cub::DeviceScan::ExclusiveSum(b.sort_tmp, tmp, keep, pos, (int)total, st) != cudaSuccess) return LV_ERR_CUDA;
    lv_add_compact_kernel<<<blocks, 256, 0, st>>>(b.xyz, total, b.vals_sorted, keep, pos, b.xyz_alt);
    uint32_t last_keep = 0, last_pos = 0;
    if (cudaMemcpyAsync(&last_keep, keep + total - 1, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return LV_ERR_CUDA;
    if (cudaMemcpyAsync(&last_pos, pos + total - 1, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return LV_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess) return LV_ERR_CUDA;
    if (cudaGetLastError() != cudaSuccess) return LV_ERR_CUDA;
    float* t = b.xyz; b.xyz = b.xyz_alt; b.xyz_alt = t;
    b.n = (int64_t)last_pos + (int64_t)last_keep;
    *launches = 3 + 9 + 2;
    return LV_OK;
}
