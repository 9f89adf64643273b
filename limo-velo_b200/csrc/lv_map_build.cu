/*
 * lv_map_build.cu — K0: (re)build of the device map's search structure, once per sweep.
 *
 * Replaces KD_TREE::Build / BuildTree (include/ikd-Tree/ikd_Tree/ikd_Tree.cpp:409-423,679-733),
 * i.e. what Mapper::add needs before the next Mapper::match can run (src/Modules/Mapper.cpp:
 * 22-30,64-76).  Instead of a median-split pointer tree: voxel key per point -> radix sort by key
 * -> gather into a float4 array (each voxel one contiguous run) -> hash table voxel -> (start,
 * count).  HBM-bound streaming work: 24 B algorithmic per map point (read 12 B, write sorted 12 B).
 *
 * The key sort is cub::DeviceRadixSort (CCCL ships with the CUDA toolkit); everything else is
 * hand-written.
 */
#include <cub/device/device_radix_sort.cuh>

#include "lv_internal.h"

namespace lv {

__global__ void __launch_bounds__(256) lv_map_keys_kernel(const float* __restrict__ xyz, int64_t n, float inv_cell,
                                                           uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    keys[i] = voxel_key(voxel_coord(x, inv_cell), voxel_coord(y, inv_cell), voxel_coord(z, inv_cell));
    vals[i] = (uint32_t)i;
}

/* gather into sorted float4 order and count voxel heads */
__global__ void __launch_bounds__(256) lv_map_gather_kernel(const float* __restrict__ xyz, int64_t n,
                                                             const uint64_t* __restrict__ keys_sorted,
                                                             const uint32_t* __restrict__ vals_sorted,
                                                             float4* __restrict__ pts, uint32_t* __restrict__ n_heads) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int head = 0;
    if (j < n) {
        const uint32_t src = vals_sorted[j];
        pts[j] = make_float4(xyz[3 * (size_t)src], xyz[3 * (size_t)src + 1], xyz[3 * (size_t)src + 2],
                             __int_as_float((int)src));
        head = (j == 0 || keys_sorted[j - 1] != keys_sorted[j]) ? 1 : 0;
    }
    const int c = __syncthreads_count(head);
    if (threadIdx.x == 0 && c) atomicAdd(n_heads, (uint32_t)c);
}

__global__ void __launch_bounds__(256) lv_map_table_clear_kernel(uint4* table, uint32_t slots) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < slots) table[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
}

/* every voxel head inserts (key -> start, count) */
__global__ void __launch_bounds__(256) lv_map_insert_kernel(const uint64_t* __restrict__ keys_sorted, int64_t n,
                                                             uint4* table, uint32_t mask) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t key = keys_sorted[j];
    if (j > 0 && keys_sorted[j - 1] == key) return;
    /* run length: short linear scan, then binary search for the rare long run */
    int64_t end = j + 1;
    int steps = 0;
    while (end < n && keys_sorted[end] == key && steps < 32) { ++end; ++steps; }
    if (end < n && keys_sorted[end] == key) {
        int64_t lo = end, hi = n;   /* first index in (lo, hi] whose key differs */
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (keys_sorted[mid] == key) lo = mid + 1; else hi = mid;
        }
        end = lo;
    }
    uint32_t slot = voxel_hash(key) & mask;
    unsigned long long* tab64 = reinterpret_cast<unsigned long long*>(table);
    for (;;) {
        const unsigned long long prev = atomicCAS(tab64 + 2 * (size_t)slot, LV_EMPTY_KEY, (unsigned long long)key);
        if (prev == LV_EMPTY_KEY) {
            uint32_t* e = reinterpret_cast<uint32_t*>(table + slot);
            e[2] = (uint32_t)j;
            e[3] = (uint32_t)(end - j);
            return;
        }
        slot = (slot + 1) & mask;
    }
}

size_t map_sort_tmp_bytes(int64_t cap) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)cap, 0, 63);
    return bytes;
}

VoxelMapView map_view(const MapBuffers& b) {
    VoxelMapView v;
    v.pts = b.pts;
    v.table = b.table;
    v.mask = b.table_mask;
    v.n_points = (uint32_t)b.n;
    v.cell = b.cell;
    v.inv_cell = b.inv_cell;
    return v;
}

cudaError_t map_rebuild(MapBuffers& b, cudaStream_t st, int* launches) {
    const int64_t n = b.n;
    int l = 0;
    if (n <= 0) { b.table_mask = 0; *launches = 0; return cudaSuccess; }
    const unsigned blocks = (unsigned)((n + 255) / 256);
    cudaError_t err;
    lv_map_keys_kernel<<<blocks, 256, 0, st>>>(b.xyz, n, b.inv_cell, b.keys, b.vals); ++l;
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    size_t tmp = b.sort_tmp_bytes;
    err = cub::DeviceRadixSort::SortPairs(b.sort_tmp, tmp, b.keys, b.keys_sorted, b.vals, b.vals_sorted, (int)n, 0, 63, st);
    l += 9;   /* histogram + onesweep passes (approximate; counted for gpu_launches) */
    if (err != cudaSuccess) return err;
    if ((err = cudaMemsetAsync(b.counter, 0, sizeof(uint32_t), st)) != cudaSuccess) return err;
    lv_map_gather_kernel<<<blocks, 256, 0, st>>>(b.xyz, n, b.keys_sorted, b.vals_sorted, b.pts, b.counter); ++l;
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    uint32_t heads = 0;
    if ((err = cudaMemcpyAsync(&heads, b.counter, sizeof(uint32_t), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return err;
    if ((err = cudaStreamSynchronize(st)) != cudaSuccess) return err;
    uint32_t slots = 1024;
    while (slots < 2u * heads && slots < b.table_cap) slots <<= 1;
    if (slots > b.table_cap) slots = b.table_cap;
    b.table_mask = slots - 1;
    lv_map_table_clear_kernel<<<(slots + 255) / 256, 256, 0, st>>>(b.table, slots); ++l;
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    lv_map_insert_kernel<<<blocks, 256, 0, st>>>(b.keys_sorted, n, b.table, b.table_mask); ++l;
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    *launches = l;
    return cudaSuccess;
}

}  // namespace lv
