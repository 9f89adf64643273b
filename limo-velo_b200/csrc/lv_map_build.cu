/*
 * lv_map_build.cu — K0: (re)build of the device map's search structure, once per sweep.
 *
 * Replaces KD_TREE::Build / BuildTree (include/ikd-Tree/ikd_Tree/ikd_Tree.cpp:409-423,679-733),
 * i.e. what Mapper::add needs before the next Mapper::match can run (src/Modules/Mapper.cpp:
 * 22-30,64-76).  Instead of a median-split pointer tree:
 *   Morton key of the finest voxel per point -> radix sort -> gather into a float4 array (every voxel
 *   of every pyramid level is one contiguous run) -> per level: hash table voxel -> (start, count)
 *   and halo buckets (each voxel's points + its neighbours' points in one run); lv_voxel_search.h.
 * HBM-bound streaming work: 24 B algorithmic per map point (read 12 B, write sorted 12 B) plus the
 * halo copies (~10 x 16 B per point and level for surface-like maps).
 *
 * The key sort and the offset scan are cub::DeviceRadixSort / cub::DeviceScan (CCCL ships with the
 * CUDA toolkit); everything else is hand-written.
 */
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "lv_internal.h"

namespace lv {

__global__ void __launch_bounds__(256) lv_map_keys_kernel(const float* __restrict__ xyz, int64_t n, float inv_cell0,
                                                           uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = morton3(voxel_coord(xyz[3 * i], inv_cell0), voxel_coord(xyz[3 * i + 1], inv_cell0),
                      voxel_coord(xyz[3 * i + 2], inv_cell0));
    vals[i] = (uint32_t)i;
}

/* gather into sorted float4 order and count the voxel heads of every level */
__global__ void __launch_bounds__(256) lv_map_gather_kernel(const float* __restrict__ xyz, int64_t n, int n_levels,
                                                             const uint64_t* __restrict__ keys_sorted,
                                                             const uint32_t* __restrict__ vals_sorted,
                                                             float4* __restrict__ pts, uint32_t* __restrict__ n_heads) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t diff = 0;       /* bits in which this key differs from its predecessor */
    bool valid = j < n;
    if (valid) {
        const uint32_t src = vals_sorted[j];
        pts[j] = make_float4(xyz[3 * (size_t)src], xyz[3 * (size_t)src + 1], xyz[3 * (size_t)src + 2],
                             __int_as_float((int)src));
        diff = j == 0 ? ~0ull : (keys_sorted[j - 1] ^ keys_sorted[j]);
    }
    for (int l = 0; l < n_levels; ++l) {
        const int c = __syncthreads_count(valid && (diff >> (3 * l)) != 0);
        if (threadIdx.x == 0 && c) atomicAdd(n_heads + l, (uint32_t)c);
    }
}

__global__ void __launch_bounds__(256) lv_map_table_clear_kernel(uint4* table, uint32_t slots) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;   /* one uint4 half-slot per thread */
    if (i < 2 * slots) table[i] = (i & 1u) ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
}

/* every voxel head of level `shift / 3` inserts (key -> start, count) */
__global__ void __launch_bounds__(256) lv_map_insert_kernel(const uint64_t* __restrict__ keys_sorted, int64_t n, int shift,
                                                             uint4* table, uint32_t mask) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t key = keys_sorted[j] >> shift;
    if (j > 0 && (keys_sorted[j - 1] >> shift) == key) return;
    /* run length: short linear scan, then binary search for the long runs of coarse levels */
    int64_t end = j + 1;
    int steps = 0;
    while (end < n && (keys_sorted[end] >> shift) == key && steps < 16) { ++end; ++steps; }
    if (end < n && (keys_sorted[end] >> shift) == key) {
        int64_t lo = end, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((keys_sorted[mid] >> shift) == key) lo = mid + 1; else hi = mid;
        }
        end = lo;
    }
    const uint64_t tkey = voxel_key_from_morton(keys_sorted[j], shift / 3);
    uint32_t slot = voxel_hash(tkey) & mask;
    unsigned long long* tab64 = reinterpret_cast<unsigned long long*>(table);
    for (;;) {
        const unsigned long long prev = atomicCAS(tab64 + 4 * (size_t)slot, LV_EMPTY_KEY, (unsigned long long)tkey);
        if (prev == LV_EMPTY_KEY) {
            uint32_t* e = reinterpret_cast<uint32_t*>(table + 2 * (size_t)slot);
            e[2] = (uint32_t)j;
            e[3] = (uint32_t)(end - j);
            return;
        }
        slot = (slot + 1) & mask;
    }
}

/* level 0 only: every occupied voxel makes sure its 26 neighbours have a slot too (count 0 if they
 * hold no points), so that a query landing in an empty voxel next to the map still finds a halo
 * bucket with one probe.  One thread per (voxel head, neighbour). */
__global__ void __launch_bounds__(256) lv_map_dilate_kernel(const uint64_t* __restrict__ keys_sorted, int64_t n,
                                                             uint4* table, uint32_t mask, uint32_t* __restrict__ n_added) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t j = t / 27;
    const int nb = (int)(t - j * 27);
    if (j >= n || nb == 13) return;
    const uint64_t key = keys_sorted[j];
    if (j > 0 && keys_sorted[j - 1] == key) return;
    const int cx = (int)compact21(key) + nb % 3 - 1, cy = (int)compact21(key >> 1) + (nb / 3) % 3 - 1,
              cz = (int)compact21(key >> 2) + nb / 9 - 1;
    if (cx < 0 || cy < 0 || cz < 0 || cx > 0x1FFFFF || cy > 0x1FFFFF || cz > 0x1FFFFF) return;
    const uint64_t nkey = voxel_key((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
    uint32_t slot = voxel_hash(nkey) & mask;
    unsigned long long* tab64 = reinterpret_cast<unsigned long long*>(table);
    for (;;) {
        const unsigned long long prev = atomicCAS(tab64 + 4 * (size_t)slot, LV_EMPTY_KEY, (unsigned long long)nkey);
        if (prev == LV_EMPTY_KEY) { atomicAdd(n_added, 1u); return; }   /* claimed: start = count = 0 from the clear */
        if (prev == (unsigned long long)nkey) return;                    /* occupied voxel or already added */
        slot = (slot + 1) & mask;
    }
}

/* ---- halo buckets ----------------------------------------------------------------------------
 * One warp per hash slot.  Lane l < 27 looks at neighbour voxel nb(l) (lane 0 = the voxel itself, so
 * its points come first in the bucket); the warp knows the neighbours' counts and their prefix
 * sums.  Pass 1 writes the bucket size, a device-wide exclusive scan turns sizes into offsets,
 * pass 2 copies the points and completes the slot: second half = {halo_start, halo_count, 0, 0}. */
__device__ __forceinline__ int halo_lane_to_nb(int lane) { return lane == 0 ? 13 : (lane <= 13 ? lane - 1 : lane); }

template <bool FILL>
__global__ void __launch_bounds__(256) lv_map_halo_kernel(uint4* table, uint32_t slots, VoxelLevel L,
                                                          const float4* __restrict__ pts, uint32_t* __restrict__ bsize,
                                                          const uint32_t* __restrict__ bstart, float4* __restrict__ halo) {
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (slot >= slots) return;
    const uint4 e = table[2 * (size_t)slot];
    if ((e.x & e.y) == 0xFFFFFFFFu) {
        if (!FILL && lane == 0) bsize[slot] = 0u;
        return;
    }
    const uint64_t key = ((uint64_t)e.y << 32) | e.x;
    const int bx = (int)(key & 0x1FFFFFu), by = (int)((key >> 21) & 0x1FFFFFu), bz = (int)((key >> 42) & 0x1FFFFFu);
    uint32_t s = 0, cnt = 0;
    if (lane < 27) {
        const int nb = halo_lane_to_nb(lane);
        const int cx = bx + nb % 3 - 1, cy = by + (nb / 3) % 3 - 1, cz = bz + nb / 9 - 1;
        if (lane == 0) { s = e.z; cnt = e.w; }
        else if (cx < 0 || cy < 0 || cz < 0 ||
                 voxel_find(L, voxel_key((uint32_t)cx, (uint32_t)cy, (uint32_t)cz), &s, &cnt) < 0) cnt = 0;
    }
    uint32_t incl = cnt;   /* inclusive warp scan of the counts */
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (!FILL) {
        if (lane == 0) bsize[slot] = total;
        return;
    }
    const uint32_t base = bstart[slot];
    const uint32_t off = base + incl - cnt;
    for (uint32_t k = 0; k < cnt; ++k) halo[off + k] = pts[s + k];
    if (lane == 0) table[2 * (size_t)slot + 1] = make_uint4(base, total, 0u, 0u);
}

size_t map_sort_tmp_bytes(int64_t cap) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)cap, 0, 63);
    size_t scan_bytes = 0;   /* exclusive scan over at most 4 * cap hash slots */
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(4 * cap + 2048));
    return bytes > scan_bytes ? bytes : scan_bytes;
}

static VoxelLevel level_view(const MapBuffers& b, int l) {
    VoxelLevel L;
    L.table = b.level[l].table;
    L.mask = b.level[l].mask;
    L.cell = b.cell * (float)(1 << l);
    return L;
}

VoxelMapView map_view(const MapBuffers& b) {
    VoxelMapView v;
    v.pts = b.pts;
    v.halo = b.halo;
    v.n_levels = b.n_levels;
    for (int l = 0; l < kMaxLevels; ++l) v.lv[l] = level_view(b, l < b.n_levels ? l : 0);
    v.n_points = (uint32_t)b.n;
    v.cell0 = b.cell;
    v.inv_cell0 = b.inv_cell;
    return v;
}

template <class T>
static cudaError_t grow(T** p, uint64_t* cap, uint64_t need) {   /* grow-only device buffer */
    if (need <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const uint64_t want = need + need / 4 + 1024;
    cudaError_t err = cudaMalloc(p, sizeof(T) * want);
    if (err == cudaSuccess) *cap = want;
    return err;
}

cudaError_t map_rebuild(MapBuffers& b, cudaStream_t st, int* launches) {
    const int64_t n = b.n;
    int l = 0;
    *launches = 0;
    if (n <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    cudaError_t err;
    lv_map_keys_kernel<<<blocks, 256, 0, st>>>(b.xyz, n, b.inv_cell, b.keys, b.vals); ++l;
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    size_t tmp = b.sort_tmp_bytes;
    err = cub::DeviceRadixSort::SortPairs(b.sort_tmp, tmp, b.keys, b.keys_sorted, b.vals, b.vals_sorted, (int)n, 0, 63, st);
    l += 9;   /* histogram + onesweep passes (approximate; counted for gpu_launches) */
    if (err != cudaSuccess) return err;
    if ((err = cudaMemsetAsync(b.counter, 0, sizeof(uint32_t) * 8, st)) != cudaSuccess) return err;
    lv_map_gather_kernel<<<blocks, 256, 0, st>>>(b.xyz, n, b.n_levels, b.keys_sorted, b.vals_sorted, b.pts, b.counter); ++l;
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    uint32_t heads[kMaxLevels] = {0, 0, 0, 0};
    if ((err = cudaMemcpyAsync(heads, b.counter, sizeof(uint32_t) * kMaxLevels, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return err;
    if ((err = cudaStreamSynchronize(st)) != cudaSuccess) return err;
    /* hash tables: level 0 sized for its dilation (typically 3-4x the occupied voxels) */
    for (int lev = 0; lev < b.n_levels; ++lev) {
        MapLevel& M = b.level[lev];
        uint64_t slots = 1024;
        const uint64_t want = lev == 0 ? 10ull * heads[0] : 2ull * heads[lev];
        while (slots < want && slots < (1ull << 30)) slots <<= 1;
        for (;;) {
            if ((err = grow(&M.table, &M.table_cap, 2 * slots)) != cudaSuccess) return err;
            M.mask = (uint32_t)(slots - 1);
            lv_map_table_clear_kernel<<<(unsigned)((2 * slots + 255) / 256), 256, 0, st>>>(M.table, (uint32_t)slots); ++l;
            lv_map_insert_kernel<<<blocks, 256, 0, st>>>(b.keys_sorted, n, 3 * lev, M.table, M.mask); ++l;
            if (lev != 0) break;
            const unsigned dblocks = (unsigned)(((uint64_t)n * 27 + 255) / 256);
            if ((err = cudaMemsetAsync(b.counter + 4, 0, sizeof(uint32_t), st)) != cudaSuccess) return err;
            lv_map_dilate_kernel<<<dblocks, 256, 0, st>>>(b.keys_sorted, n, M.table, M.mask, b.counter + 4); ++l;
            uint32_t added = 0;
            if ((err = cudaMemcpyAsync(&added, b.counter + 4, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return err;
            if ((err = cudaStreamSynchronize(st)) != cudaSuccess) return err;
            if ((uint64_t)heads[0] + added <= (slots * 6) / 10 || slots >= (1ull << 30)) break;
            slots <<= 1;   /* load factor above 0.6 (scattered map): redo with a larger table */
        }
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    /* level-0 halo buckets: sizes -> offsets -> (re)allocate -> fill */
    {
        MapLevel& M = b.level[0];
        const uint32_t slots = M.mask + 1;
        if ((err = grow(&b.bsize, &b.bs_cap, slots)) != cudaSuccess) return err;
        if ((err = grow(&b.bstart, &b.bs_cap2, slots)) != cudaSuccess) return err;
        const unsigned wblocks = (unsigned)(((uint64_t)slots * 32 + 255) / 256);
        lv_map_halo_kernel<false><<<wblocks, 256, 0, st>>>(M.table, slots, level_view(b, 0), b.pts, b.bsize, nullptr, nullptr); ++l;
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, b.bsize, b.bstart, (int)slots);
        if (scan_bytes > b.sort_tmp_bytes) {
            cudaFree(b.sort_tmp);
            b.sort_tmp = nullptr;
            if ((err = cudaMalloc(&b.sort_tmp, scan_bytes)) != cudaSuccess) return err;
            b.sort_tmp_bytes = scan_bytes;
        }
        tmp = b.sort_tmp_bytes;
        if ((err = cub::DeviceScan::ExclusiveSum(b.sort_tmp, tmp, b.bsize, b.bstart, (int)slots, st)) != cudaSuccess) return err;
        l += 2;
        uint32_t last[2] = {0, 0};
        if ((err = cudaMemcpyAsync(&last[0], b.bsize + slots - 1, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return err;
        if ((err = cudaMemcpyAsync(&last[1], b.bstart + slots - 1, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return err;
        if ((err = cudaStreamSynchronize(st)) != cudaSuccess) return err;
        const uint64_t halo_n = (uint64_t)last[0] + last[1];
        if (halo_n > 0x7FFFFFF0ull) return cudaErrorMemoryAllocation;   /* ids are int32 */
        if ((err = grow(&b.halo, &b.halo_cap, halo_n)) != cudaSuccess) return err;
        b.halo_n = halo_n;
        lv_map_halo_kernel<true><<<wblocks, 256, 0, st>>>(M.table, slots, level_view(b, 0), b.pts, nullptr, b.bstart, b.halo); ++l;
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    *launches = l;
    return cudaSuccess;
}

}  // namespace lv
