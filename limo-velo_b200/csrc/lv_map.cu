/*
 * lv_map.cu — Mapper::add / KD_TREE::Build / Add_Points on the device map (layout and per-item logic: lv_voxel_map.h).
 *
 * Replaces KD_TREE::Build (include/ikd-Tree/ikd_Tree/ikd_Tree.cpp:409-423) and Add_Points with the voxel-downsample
 * rule (:478-573) as called from Mapper::add (src/Modules/Mapper.cpp:22-30,64-76).  One add is six kernels plus a
 * radix sort of the NEW points only, all on the handle's stream, with every count on the device:
 *
 *   begin    reset the touched / dirty lists
 *   keys     per new point: cell (floor(x / ds), ikd_Tree.cpp:493), voxel, find-or-insert the voxel's slot,
 *            32-bit sort key (slot, cell inside the voxel)
 *   sort     cub::DeviceRadixSort over the n new (key, index) pairs            [library: CCCL ships with the toolkit]
 *   merge    one thread per touched voxel: the reference's rule per cell (map_merge_run)
 *   dilate   touched voxels create their neighbours' slots and mark the 27 neighbourhoods dirty
 *   halo     one warp per dirty voxel: 27 probes, warp scan, coalesced copy of the neighbours' points into the
 *            voxel's halo bucket (in place when it fits its capacity, else a fresh extent)
 *
 * HBM traffic is proportional to the sweep: ~20 B per new point for keys and sort, ~0.6 KB per dirty voxel for its
 * bucket (27 x 32 B probes + the copy); the 1 M points already in the map are not touched.  Nothing synchronises the
 * stream: errors (table / arena exhausted) are flagged on the device and read with the next result that is read
 * anyway (lv_map_size, lv_correct, lv_synchronize).
 */
#include <cub/device/device_radix_sort.cuh>

#include "lv_internal.h"

namespace lv {

__global__ void __launch_bounds__(256) lv_map_clear_kernel(uint4* table, uint32_t slots, uint4* btable, uint32_t bslots,
                                                            uint32_t* counters) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = i0; i < 2 * (size_t)slots; i += stride)
        table[i] = (i & 1u) ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
    for (size_t i = i0; i < (size_t)bslots; i += stride) btable[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
    if (i0 < (size_t)kMapCounters) counters[i0] = i0 == (size_t)kCtrArenaTop ? 16u : 0u;   /* position 0 is never handed out */
}

__global__ void lv_map_begin_kernel(uint32_t* counters) {
    if (threadIdx.x == 0) { counters[kCtrTouched] = 0u; counters[kCtrDirty] = 0u; }
}

__global__ void __launch_bounds__(256) lv_map_keys_kernel(const VoxelMapRW m, const float* __restrict__ xyz, uint32_t n,
                                                           uint32_t* __restrict__ skeys, uint32_t* __restrict__ svals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    skeys[i] = map_point_key(m, xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]);
    svals[i] = i;
}

__global__ void __launch_bounds__(128) lv_map_merge_kernel(const VoxelMapRW m, const uint32_t* __restrict__ skeys,
                                                            const uint32_t* __restrict__ svals, uint32_t n,
                                                            const float* __restrict__ xyz, uint32_t id_base, int downsample) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t key = skeys[j];
    if (key == 0xFFFFFFFFu) return;                                          /* not storable (non-finite / table full) */
    if (j > 0 && (skeys[j - 1] >> kCellBits) == (key >> kCellBits)) return;  /* not the head of its voxel's run */
    map_merge_run(m, skeys, svals, j, n, xyz, id_base, downsample);
}

__global__ void __launch_bounds__(256) lv_map_dilate_kernel(const VoxelMapRW m) {
    const uint32_t n = min(m.counters[kCtrTouched], m.list_cap) * 27u;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) map_dilate_item(m, m.touched[t / 27u], (int)(t % 27u));
}

/* one warp per dirty voxel (the warp form of map_halo_voxel_serial) */
__global__ void __launch_bounds__(256) lv_map_halo_kernel(const VoxelMapRW m) {
    const uint32_t n = min(m.counters[kCtrDirty], m.list_cap);
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t d = warp; d < n; d += n_warps) {
        const uint32_t slot = m.dirty[d];
        uint32_t* s32 = reinterpret_cast<uint32_t*>(m.table + 2 * (size_t)slot);
        const uint64_t key = (uint64_t)s32[0] | ((uint64_t)s32[1] << 32);
        const int bx = (int)((uint32_t)key & 0x1FFFFFu), by = (int)((uint32_t)(key >> 21) & 0x1FFFFFu), bz = (int)((uint32_t)(key >> 42) & 0x1FFFFFu);
        uint32_t s = 0, cnt = 0;
        if (lane < 27) {
            const int nb = halo_lane_to_nb(lane);
            const int cx = bx + nb % 3 - 1, cy = by + (nb / 3) % 3 - 1, cz = bz + nb / 9 - 1;
            if (cx >= 0 && cy >= 0 && cz >= 0 && cx <= 0x1FFFFF && cy <= 0x1FFFFF && cz <= 0x1FFFFF) {
                const int ns = lane == 0 ? (int)slot : voxel_find_rw(m, voxel_key((uint32_t)cx, (uint32_t)cy, (uint32_t)cz));
                if (ns >= 0) {
                    const uint32_t* n32 = reinterpret_cast<const uint32_t*>(m.table + 2 * (size_t)ns);
                    s = n32[2];
                    cnt = n32[3];
                }
            }
        }
        uint32_t incl = cnt;   /* inclusive warp scan of the counts */
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, k);
            if (lane >= k) incl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t excl = incl - cnt;
        uint32_t hstart = 0, caps = 0, hcap = 0;
        int ok = 1;
        if (lane == 0) {
            hstart = s32[4];
            caps = s32[6];
            hcap = caps_halo(caps);
            if (total > hcap) {
                const uint32_t ncap = halo_new_cap(total, hcap);
                if (total > ncap) { atomicOr(m.counters + kCtrError, (uint32_t)kErrExtentTooLarge); ok = 0; }
                else {
                    const uint32_t at = arena_alloc(m, ncap);
                    if (at == 0xFFFFFFFFu) ok = 0;
                    else { hstart = at; hcap = ncap; }
                }
            }
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        hstart = __shfl_sync(0xffffffffu, hstart, 0);
        if (ok) {
            for (uint32_t j0 = 0; j0 < total; j0 += 32) {                     /* coalesced 16-byte stores, gathers from <= 27 runs */
                const uint32_t j = j0 + (uint32_t)lane;
                int lo = 0;                                                  /* first neighbour whose inclusive prefix exceeds j */
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const uint32_t v = __shfl_sync(0xffffffffu, incl, lo + step - 1);
                    if (v <= j) lo += step;
                }
                lo = lo > 31 ? 31 : lo;
                const uint32_t sv = __shfl_sync(0xffffffffu, s, lo), ev = __shfl_sync(0xffffffffu, excl, lo);
                if (j < total) m.arena[hstart + j] = m.arena[sv + (j - ev)];
            }
        }
        if (lane == 0) {
            if (ok) {
                s32[4] = hstart;
                s32[5] = total;
                s32[6] = caps_pack(caps_own(caps), hcap);
            }
            s32[7] = 0u;
        }
        __syncwarp();
    }
}

/* lv_map_points: all own points -> out[], unordered; cursor = counters[kCtrGather] */
__global__ void __launch_bounds__(256) lv_map_gather_kernel(const VoxelMapRW m, float4* __restrict__ out, uint32_t out_cap) {
    const uint32_t slots = m.mask + 1u;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < slots; slot += stride) {
        const uint4 e = m.table[2 * (size_t)slot];
        if ((e.x & e.y) == 0xFFFFFFFFu || e.w == 0u) continue;
        const uint32_t at = atomicAdd(m.counters + kCtrGather, e.w);
        for (uint32_t t = 0; t < e.w && at + t < out_cap; ++t) out[at + t] = m.arena[e.z + t];
    }
}
__global__ void __launch_bounds__(256) lv_map_ids_kernel(const float4* __restrict__ pts, uint32_t n, uint32_t* __restrict__ ids,
                                                          uint32_t* __restrict__ idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ids[i] = (uint32_t)__float_as_int(pts[i].w);
    idx[i] = i;
}
__global__ void __launch_bounds__(256) lv_map_unpack_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ idx, uint32_t n,
                                                             float* __restrict__ xyz) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[idx[i]];
    xyz[3 * (size_t)i] = p.x; xyz[3 * (size_t)i + 1] = p.y; xyz[3 * (size_t)i + 2] = p.z;
}

/* main.cpp:99-105 on the device: the sweep (LiDAR frame) in world coordinates with the state the update just produced,
 * Xt2 * Xt2.I_Rt_L() * p in fp32 (State.cpp:79-89, RotTransl.cpp:36-48) — the very transform Mapper::match applies */
__global__ void __launch_bounds__(256) lv_sweep_to_world_kernel(const UpdateCtrl* __restrict__ c, const float* __restrict__ xyz, uint32_t n,
                                                                 float* __restrict__ out) {
    __shared__ Frame s_frame;
    if (threadIdx.x == 0) make_frame(c->x, &s_frame);
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g[3];
    rt_apply(s_frame.lidar_to_world, xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], g);
    out[3 * (size_t)i] = g[0]; out[3 * (size_t)i + 1] = g[1]; out[3 * (size_t)i + 2] = g[2];
}

/* ---- host side ------------------------------------------------------------------------------------------- */
static uint32_t pow2_at_least(uint64_t v, uint32_t lo, uint32_t hi) {
    uint64_t p = lo;
    while (p < v && p < hi) p <<= 1;
    return (uint32_t)p;
}

cudaError_t map_alloc(MapBuffers& b, int64_t max_map_points, int64_t max_points, float voxel_size, float ds) {
    memset(&b, 0, sizeof(b));
    b.cap = max_map_points;
    b.grid.ds = ds;
    int k = (int)floor((double)voxel_size / (double)ds + 0.5);
    k = k < 1 ? 1 : (k > kMaxCellsPerVoxel ? kMaxCellsPerVoxel : k);
    b.grid.k = k;
    b.grid.cell0 = (float)k * ds;
    /* occupied + dilated voxels of a surface-like map: ~1 slot per map point; 4x keeps the load below 0.3, scattered
     * maps (27 slots per isolated point) run out and report LV_ERR_CAPACITY instead of hanging */
    b.slots = pow2_at_least(4ull * (uint64_t)max_map_points, 1u << 16, 1u << 26);
    b.bslots = pow2_at_least((uint64_t)b.slots / 4u, 1u << 14, 1u << 27);
    /* own extents (~1.5x the map) + halo buckets (27 copies of every point, up to 50 % slack) + the extents that
     * buckets outgrow (geometric growth: at most twice the live ones): 1.5 KB per map point of capacity */
    uint64_t arena = 96ull * (uint64_t)max_map_points + (4ull << 20);
    if (arena > 0x7FFFFFF0ull) arena = 0x7FFFFFF0ull;                       /* ids are int32 */
    b.arena_cap = (uint32_t)arena;
    b.list_cap = b.slots;
    b.add_cap = max_map_points > max_points ? max_map_points : max_points;
    cudaError_t e;
    if ((e = cudaMalloc(&b.table, sizeof(uint4) * 2 * (size_t)b.slots)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&b.btable, sizeof(uint4) * (size_t)b.bslots)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&b.arena, sizeof(float4) * (size_t)b.arena_cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&b.counters, sizeof(uint32_t) * kMapCounters)) != cudaSuccess) return e;
    if ((e = cudaMallocHost(&b.h_counters, sizeof(uint32_t) * kMapCounters)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&b.touched, sizeof(uint32_t) * (size_t)b.list_cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&b.dirty, sizeof(uint32_t) * (size_t)b.list_cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&b.stage_xyz, sizeof(float) * 3 * (size_t)b.add_cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&b.skeys, sizeof(uint32_t) * (size_t)b.add_cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&b.skeys_alt, sizeof(uint32_t) * (size_t)b.add_cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&b.svals, sizeof(uint32_t) * (size_t)b.add_cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&b.svals_alt, sizeof(uint32_t) * (size_t)b.add_cap)) != cudaSuccess) return e;
    b.sort_bits = kCellBits;
    for (uint32_t s = b.slots; s > 1; s >>= 1) b.sort_bits++;
    b.sort_tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b.sort_tmp_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)b.add_cap, 0, 32);
    if ((e = cudaMalloc(&b.sort_tmp, b.sort_tmp_bytes)) != cudaSuccess) return e;
    b.empty = true;
    return cudaSuccess;
}

void map_free(MapBuffers& b) {
    cudaFree(b.table); cudaFree(b.btable); cudaFree(b.arena); cudaFree(b.counters); cudaFreeHost(b.h_counters);
    cudaFree(b.touched); cudaFree(b.dirty); cudaFree(b.stage_xyz); cudaFree(b.skeys); cudaFree(b.skeys_alt);
    cudaFree(b.svals); cudaFree(b.svals_alt); cudaFree(b.sort_tmp);
    memset(&b, 0, sizeof(b));
}

static VoxelMapRW map_rw(const MapBuffers& b) {
    VoxelMapRW m;
    m.table = b.table; m.mask = b.slots - 1u; m.btable = b.btable; m.bmask = b.bslots - 1u;
    m.arena = b.arena; m.arena_cap = b.arena_cap; m.counters = b.counters; m.touched = b.touched; m.dirty = b.dirty;
    m.list_cap = b.list_cap; m.grid = b.grid;
    return m;
}
VoxelMapView map_view(const MapBuffers& b) { return map_view_of(map_rw(b)); }

cudaError_t map_clear(MapBuffers& b, cudaStream_t st, int* launches) {
    lv_map_clear_kernel<<<148 * 8, 256, 0, st>>>(b.table, b.slots, b.btable, b.bslots, b.counters);
    b.n_inserted = 0;
    b.empty = true;
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t map_add(MapBuffers& b, const float* d_xyz, int64_t n, int downsample, cudaStream_t st, int* launches) {
    if (n <= 0) return cudaSuccess;
    if (n > b.add_cap) return cudaErrorInvalidValue;
    const VoxelMapRW m = map_rw(b);
    const uint32_t n32 = (uint32_t)n;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    lv_map_begin_kernel<<<1, 32, 0, st>>>(b.counters);
    lv_map_keys_kernel<<<blocks, 256, 0, st>>>(m, d_xyz, n32, b.skeys, b.svals);
    size_t tmp = b.sort_tmp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(b.sort_tmp, tmp, b.skeys, b.skeys_alt, b.svals, b.svals_alt, (int)n, 0, b.sort_bits, st);
    if (e != cudaSuccess) return e;
    lv_map_merge_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(m, b.skeys_alt, b.svals_alt, n32, d_xyz, (uint32_t)b.n_inserted, downsample);
    /* touched <= n voxels, dirty <= 27 n: fixed grids striding over the device-side counts */
    uint64_t dil = ((uint64_t)n * 27 + 255) / 256;
    dil = dil < 1 ? 1 : (dil > 148u * 8u ? 148u * 8u : dil);
    lv_map_dilate_kernel<<<(unsigned)dil, 256, 0, st>>>(m);
    uint64_t hal = ((uint64_t)n * 27 + 7) / 8;
    hal = hal < 1 ? 1 : (hal > 148u * 64u ? 148u * 64u : hal);
    lv_map_halo_kernel<<<(unsigned)hal, 256, 0, st>>>(m);
    b.n_inserted += n;
    b.empty = false;
    if (launches) *launches += 5 + 4;   /* + histogram and onesweep passes of the sort (approximate) */
    return cudaGetLastError();
}

cudaError_t map_add_sweep(MapBuffers& b, const UpdateCtrl* d_ctrl, const float* d_xyz_lidar, int64_t n, int downsample, cudaStream_t st,
                          int* launches) {
    if (n <= 0) return cudaSuccess;
    if (n > b.add_cap) return cudaErrorInvalidValue;
    lv_sweep_to_world_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_ctrl, d_xyz_lidar, (uint32_t)n, b.stage_xyz);
    if (launches) *launches += 1;
    return map_add(b, b.stage_xyz, n, b.empty ? 0 : downsample, st, launches);
}

/* device counters -> pinned mirror; the caller synchronises the stream before reading b.h_counters */
cudaError_t map_fetch_counters(MapBuffers& b, cudaStream_t st) {
    return cudaMemcpyAsync(b.h_counters, b.counters, sizeof(uint32_t) * kMapCounters, cudaMemcpyDeviceToHost, st);
}

/* all map points in insertion order (ascending id) -> d_out (n x 3 floats); returns the count through *n_out.
 * Synchronises the stream; allocates scratch on the spot (an inspection call, not on the hot path). */
cudaError_t map_points_sorted(MapBuffers& b, float* host_out, int64_t cap, int64_t* n_out, cudaStream_t st) {
    cudaError_t e;
    if ((e = map_fetch_counters(b, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    const int64_t n = (int64_t)(int32_t)b.h_counters[kCtrPoints];
    *n_out = n;
    if (n <= 0 || !host_out || cap <= 0) return cudaSuccess;
    float4* pts = nullptr;
    uint32_t *ids = nullptr, *idx = nullptr, *ids2 = nullptr, *idx2 = nullptr;
    float* xyz = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ids, ids2, idx, idx2, (int)n, 0, 32);
    e = cudaMalloc(&pts, sizeof(float4) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&ids, sizeof(uint32_t) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&idx, sizeof(uint32_t) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&ids2, sizeof(uint32_t) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&idx2, sizeof(uint32_t) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&xyz, sizeof(float) * 3 * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes);
    if (e == cudaSuccess) {
        const VoxelMapRW m = map_rw(b);
        const unsigned blocks = (unsigned)((n + 255) / 256);
        cudaMemsetAsync(b.counters + kCtrGather, 0, sizeof(uint32_t), st);
        lv_map_gather_kernel<<<148 * 8, 256, 0, st>>>(m, pts, (uint32_t)n);
        lv_map_ids_kernel<<<blocks, 256, 0, st>>>(pts, (uint32_t)n, ids, idx);
        e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, ids, ids2, idx, idx2, (int)n, 0, 32, st);
        lv_map_unpack_kernel<<<blocks, 256, 0, st>>>(pts, idx2, (uint32_t)n, xyz);
        const int64_t m_out = n < cap ? n : cap;
        if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, xyz, sizeof(float) * 3 * (size_t)m_out, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    cudaFree(pts); cudaFree(ids); cudaFree(idx); cudaFree(ids2); cudaFree(idx2); cudaFree(xyz); cudaFree(tmp);
    return e;
}

}  // namespace lv
