/*
 * lv_downsample.cu — the two point-cloud downsamplers in front of the localization path, on the device.
 *
 *   temporal   PointCloudProcessor::temporal_downsample (src/Utils/PointCloudProcessor.cpp:101-112): every
 *              `downsample_rate`-th point of the raw message, and only if farther than `min_dist`.
 *              flag -> select (order kept) -> gather.
 *   voxel grid Compensator::voxelgrid_downsample (src/Modules/Compensator.cpp:148-163) = pcl::VoxelGrid with a
 *              cubic leaf: bounding box -> PCL cell index per point -> stable radix sort -> one centroid per run,
 *              runs in ascending cell index (pcl/filters/impl/voxel_grid.hpp, applyFilter with
 *              downsample_all_data_ = true: CentroidPoint / AccumulatorXYZ: fp32 sum, divided by the count).
 *              PCL's std::sort leaves the order inside a leaf unspecified; here it is input order.
 * Everything stays on the device (the bounding box is consumed by the key kernel, not by the host); the host
 * reads back the output count and one status word.  Streaming integer/byte work: 12 B in per point, 12 B out per
 * leaf, a 32-bit key-value sort between.
 */
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "lv_internal.h"

namespace lv {

struct GridBox {          /* written by the bounding-box kernel, read by the key kernel */
    int mn[3], mx[3];     /* floats in an order-preserving int encoding */
    int overflow;         /* PCL refuses when the cell count exceeds INT32_MAX */
};

__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

__global__ void lv_ds_box_init_kernel(GridBox* b) {
    if (threadIdx.x < 3) { b->mn[threadIdx.x] = 0x7fffffff; b->mx[threadIdx.x] = (int)0x80000000; }
    if (threadIdx.x == 3) b->overflow = 0;
}

__global__ void __launch_bounds__(256) lv_ds_box_kernel(const float* __restrict__ xyz, int64_t n, GridBox* b) {   /* getMinMax3D */
    int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int v = f2ord(xyz[3 * i + a]);
            mn[a] = min(mn[a], v);
            mx[a] = max(mx[a], v);
        }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], s));
            mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], s));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&b->mn[a], mn[a]); atomicMax(&b->mx[a], mx[a]); }
    }
}

/* voxel_grid.hpp: min_b = floor(min_p * inv), div_b = floor(max_p * inv) - min_b + 1, idx = ijk . (1, div0, div0 div1) */
__global__ void __launch_bounds__(256) lv_ds_keys_kernel(const float* __restrict__ xyz, int64_t n, float inv, GridBox* b,
                                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int minb[3], divb[3];
    long long cells = 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = ord2f(b->mn[a]), hi = ord2f(b->mx[a]);
        minb[a] = (int)floorf(fmul(lo, inv));
        divb[a] = (int)floorf(fmul(hi, inv)) - minb[a] + 1;
        cells *= (long long)fmul(fsub(hi, lo), inv) + 1;
    }
    if (i == 0 && cells > 2147483647LL) b->overflow = 1;
    if (i >= n) return;
    const int mul[3] = {1, divb[0], divb[0] * divb[1]};
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) idx += (int)fsub(floorf(fmul(xyz[3 * i + a], inv)), (float)minb[a]) * mul[a];
    keys[i] = (uint32_t)idx;
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) lv_ds_heads_kernel(const uint32_t* __restrict__ keys, int64_t n, uint8_t* __restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

/* one thread per leaf: AccumulatorXYZ (xyz += p in sorted order), centroid = xyz / n */
__global__ void __launch_bounds__(256) lv_ds_centroid_kernel(const float* __restrict__ xyz, const uint32_t* __restrict__ vals,
                                                             const uint32_t* __restrict__ starts, const int* __restrict__ n_runs,
                                                             int64_t n, float* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int runs = *n_runs;
    if (r >= runs) return;
    const uint32_t s = starts[r], e = r + 1 < runs ? starts[r + 1] : (uint32_t)n;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (uint32_t j = s; j < e; ++j) {
        const uint32_t p = vals[j];
        sx = fadd(sx, xyz[3 * (size_t)p]); sy = fadd(sy, xyz[3 * (size_t)p + 1]); sz = fadd(sz, xyz[3 * (size_t)p + 2]);
    }
    const float cnt = (float)(e - s);
    out[3 * (size_t)r] = fdiv(sx, cnt); out[3 * (size_t)r + 1] = fdiv(sy, cnt); out[3 * (size_t)r + 2] = fdiv(sz, cnt);
}

/* temporal_downsample: keep = (rate <= 1 or (i + 1) % rate == 0) and min_dist < |p| */
__global__ void __launch_bounds__(256) lv_ds_temporal_flag_kernel(const float* __restrict__ xyz, int64_t n, int rate, double min_dist,
                                                                  uint8_t* __restrict__ keep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const float nrm = fsqrt(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z)));
    const bool k = (rate <= 1 || (i + 1) % rate == 0) && min_dist < (double)nrm;
    keep[i] = k ? 1 : 0;
}
__global__ void __launch_bounds__(256) lv_ds_gather_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                                                           const int* __restrict__ n_sel, float* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *n_sel) return;
    const size_t p = (size_t)idx[r];
    out[3 * (size_t)r] = xyz[3 * p]; out[3 * (size_t)r + 1] = xyz[3 * p + 1]; out[3 * (size_t)r + 2] = xyz[3 * p + 2];
}

/* ---- scratch + launchers ------------------------------------------------------------------------ */
cudaError_t ds_reserve(DownsampleScratch& s, int64_t n) {
    if (n <= s.cap) return cudaSuccess;
    cudaFree(s.keys); cudaFree(s.keys_sorted); cudaFree(s.vals); cudaFree(s.vals_sorted); cudaFree(s.flags); cudaFree(s.sel);
    cudaFree(s.tmp);
    if (!s.box) {
        cudaError_t e = cudaMalloc(&s.box, sizeof(GridBox));
        if (e != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.count, sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMallocHost(&s.h_count, 2 * sizeof(int))) != cudaSuccess) return e;
    }
    const int64_t cap = n + n / 4 + 1024;
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)cap);
    cub::DeviceSelect::Flagged(nullptr, b, cub::CountingInputIterator<int32_t>(0), (const uint8_t*)nullptr, (int32_t*)nullptr,
                               (int*)nullptr, (int)cap);
    s.tmp_bytes = (a > b ? a : b) + 256;
    cudaError_t e;
    if ((e = cudaMalloc(&s.keys, sizeof(uint32_t) * cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&s.keys_sorted, sizeof(uint32_t) * cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&s.vals, sizeof(uint32_t) * cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&s.vals_sorted, sizeof(uint32_t) * cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&s.flags, sizeof(uint8_t) * cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&s.sel, sizeof(int32_t) * cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&s.tmp, s.tmp_bytes)) != cudaSuccess) return e;
    s.cap = cap;
    return cudaSuccess;
}
void ds_free(DownsampleScratch& s) {
    cudaFree(s.keys); cudaFree(s.keys_sorted); cudaFree(s.vals); cudaFree(s.vals_sorted); cudaFree(s.flags); cudaFree(s.sel);
    cudaFree(s.tmp); cudaFree(s.box); cudaFree(s.count); cudaFreeHost(s.h_count);
    s = DownsampleScratch();
}

/* enqueues everything; afterwards s.h_count[0] = number of leaves, s.h_count[1] = overflow flag (after a stream sync) */
cudaError_t launch_voxelgrid(DownsampleScratch& s, const float* d_xyz, int64_t n, float leaf, float* d_out, cudaStream_t st,
                             int* launches) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    GridBox* box = static_cast<GridBox*>(s.box);
    const float inv = fdiv(1.0f, leaf);
    lv_ds_box_init_kernel<<<1, 32, 0, st>>>(box);
    lv_ds_box_kernel<<<grid < 592u ? grid : 592u, 256, 0, st>>>(d_xyz, n, box);
    lv_ds_keys_kernel<<<grid, 256, 0, st>>>(d_xyz, n, inv, box, s.keys, s.vals);
    size_t tmp = s.tmp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(s.tmp, tmp, s.keys, s.keys_sorted, s.vals, s.vals_sorted, (int)n, 0, 32, st);
    if (e != cudaSuccess) return e;
    lv_ds_heads_kernel<<<grid, 256, 0, st>>>(s.keys_sorted, n, s.flags);
    tmp = s.tmp_bytes;
    e = cub::DeviceSelect::Flagged(s.tmp, tmp, cub::CountingInputIterator<int32_t>(0), s.flags, s.sel, s.count, (int)n, st);
    if (e != cudaSuccess) return e;
    lv_ds_centroid_kernel<<<grid, 256, 0, st>>>(d_xyz, s.vals_sorted, reinterpret_cast<const uint32_t*>(s.sel), s.count, n, d_out);
    cudaMemcpyAsync(&s.h_count[0], s.count, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&s.h_count[1], &box->overflow, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (launches) *launches += 10;
    return cudaGetLastError();
}

cudaError_t launch_temporal(DownsampleScratch& s, const float* d_xyz, int64_t n, int rate, double min_dist, float* d_out,
                            int32_t* d_idx_out, cudaStream_t st, int* launches) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    lv_ds_temporal_flag_kernel<<<grid, 256, 0, st>>>(d_xyz, n, rate, min_dist, s.flags);
    size_t tmp = s.tmp_bytes;
    int32_t* idx = d_idx_out ? d_idx_out : s.sel;
    cudaError_t e = cub::DeviceSelect::Flagged(s.tmp, tmp, cub::CountingInputIterator<int32_t>(0), s.flags, idx, s.count, (int)n, st);
    if (e != cudaSuccess) return e;
    lv_ds_gather_kernel<<<grid, 256, 0, st>>>(d_xyz, idx, s.count, d_out);
    cudaMemcpyAsync(&s.h_count[0], s.count, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (launches) *launches += 4;
    return cudaGetLastError();
}

}  // namespace lv
