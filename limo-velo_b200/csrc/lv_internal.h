/*
 * lv_internal.h — launcher declarations shared by the translation units of liblimovelo_b200.so.
 */
#ifndef LV_INTERNAL_H_
#define LV_INTERNAL_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/limovelo_b200.h"
#include "lv_ieskf.h"
#include "lv_voxel_search.h"

namespace lv {

enum { kMeasureThreads = 128, kPartialStride = 96, kStepThreads = 512 };
/* The search kernel appends its uncertified queries to one of 32 lists picked by block index: thousands of atomics on
 * ONE counter cost it a 9 us tail on the first evaluation of an update (tools/timeline.py). */
enum { kHardBuckets = 32, kCounters = 4 + kHardBuckets };
/* a bucket receives the queries of every 32nd block: at most n / 32 + (queries per block <= 128) of them */
inline uint32_t hard_segment(int64_t n) { return (uint32_t)((n + kHardBuckets - 1) / kHardBuckets + 160); }

/* the sweep of an update replayed from a CUDA graph: written by the begin kernel (whose arguments are the
 * only thing patched per launch), read by the measurement kernels instead of MeasureArgs::xyz / n / n_tiles */
struct MeasureJob {
    const float* xyz;
    int32_t n;
    int32_t n_tiles;
};

/* per-launch constants of the measurement kernels (reuse, search, search-upper, fit) */
struct MeasureArgs {
    const float* xyz;          /* n x 3 packed, LiDAR frame                               */
    int32_t n;                 /* with `job`: the capacity the grids were sized for       */
    int32_t n_tiles;           /* ceil(n / kMeasureThreads)                               */
    const MeasureJob* job;     /* non-NULL: xyz / n / n_tiles come from device memory      */
    VoxelMapView map;
    const UpdateCtrl* ctrl;    /* frame + done flag                                       */
    UpdateCtrl* prep;          /* non-NULL: block 0 of the fit kernel runs ieskf_prepare() */
    float max_d2;              /* smallest float >= MAX_DIST_PLANE^2 (search radius^2)    */
    double gate_d2;            /* MAX_DIST_PLANE^2 in double (Plane.cpp:42)               */
    int32_t max_ring;
    float planes_threshold;
    int32_t estimate_extrinsics;
    double* partials;          /* [grid][kPartialStride]: 78 + 12 sums, count             */
    /* optional per-point outputs (NULL on the hot path) */
    uint8_t* valid;
    int32_t* nn_idx;
    float* nn_sqd;
    float* plane;
    float* dist;
    float* g_world;
    double* rows;              /* n x 13 (row[12], h)                                     */
    int4* nn_a;                /* n: search result handed from K1 to K2 (neighbours 0..3)      */
    int2* nn_b;                /* n: (neighbour 4, bits of the 5th squared distance)         */
    uint32_t* hard_list;       /* queries level 0 could not certify (K1 -> K1b): kHardBuckets segments of hard_seg entries */
    uint32_t hard_seg;         /* capacity of one segment                                      */
    uint32_t* hard_count;      /* [2] length of redo_list, [4 .. 4 + kHardBuckets) lengths of the segments; kCounters words */
    /* reuse of neighbours across the evaluations of one update (NULL: off) */
    float4* ref;               /* n: world position the stored neighbours were searched from + outsider bound */
    uint32_t* redo_list;       /* n: queries whose neighbours could not be reused (Kv -> K1)  */
};

/* one level of the voxel pyramid */
struct MapLevel {
    uint4* table;              /* hash slots (2 x uint4 each), grown on demand             */
    uint64_t table_cap;        /* allocated uint4 elements                                 */
    uint32_t mask;             /* slots in use - 1                                         */
};

/* device map storage + scratch for the per-sweep rebuild */
struct MapBuffers {
    float* xyz;                /* map points, insertion order (flatten order), cap x 3     */
    float* xyz_alt;            /* second buffer: lv_map_add compacts into it, then swaps    */
    int64_t n, cap;
    uint64_t* keys;            /* cap */
    uint64_t* keys_sorted;     /* cap */
    uint32_t* vals;
    uint32_t* vals_sorted;
    float4* pts;               /* cap, Morton-sorted                                       */
    MapLevel level[kMaxLevels];
    int32_t n_levels;
    float4* halo;              /* level-0 halo buckets, grown on demand                    */
    uint64_t halo_cap, halo_n;
    uint32_t* bsize;           /* per level-0 slot: halo bucket size / offset              */
    uint32_t* bstart;
    uint64_t bs_cap, bs_cap2;
    uint32_t* counter;         /* device scalars                                           */
    void* sort_tmp;
    size_t sort_tmp_bytes;
    float cell, inv_cell;      /* finest level                                             */
};

size_t map_sort_tmp_bytes(int64_t cap);
/* K0: rebuild the hashed-voxel structure from b.xyz[0..n).  Returns launches issued; synchronises
 * the stream once (to size the hash table).                                                    */
cudaError_t map_rebuild(MapBuffers& b, cudaStream_t st, int* launches);
VoxelMapView map_view(const MapBuffers& b);

int measure_grid(int n);
/* `probe` (optional) is called on the launching thread before the search (stage 0), after it (1), after the
 * upper-level search (2) and after the fit (3): the profiler records its events there */
struct MeasureProbe { void (*at)(void* ctx, int stage); void* ctx; };
/* reuse != 0 (needs a.ref, evaluations after the first of an update): lv_reuse_kernel first, then the search only
 * over the queries it could not vouch for; the probe then sees stage 4 before the reuse kernel */
cudaError_t launch_measure(const MeasureArgs& a, int grid, cudaStream_t st, const MeasureProbe* probe = nullptr,
                           int reuse = 0, int pdl = 0);
cudaError_t launch_ieskf_begin(UpdateCtrl* c, MeasureJob* job, const float* xyz, int n, uint32_t* counters, cudaStream_t st);
const void* ieskf_begin_kernel_ptr();
void measure_init();                          /* constant tables; call once before any capture           */
/* the three kernels of launch_measure() (search instance, search-upper, fit) with their launch shapes, for
 * patching the nodes of a captured update when the map view changes */
struct MeasureKernelShape { const void* func; unsigned grid, block; };
enum { kMeasureKernels = 5 };   /* search, search-upper, fit, search over the redo list, reuse */
void measure_kernel_shapes(const MeasureArgs& a, int grid, MeasureKernelShape out[kMeasureKernels]);
cudaError_t launch_ieskf_step(UpdateCtrl* c, const IeskfParams& prm, const double* partials, int n_partials,
                              cudaStream_t st, int pdl = 0);
/* stand-alone reduction of the partials (operator-boundary calls): out[0:144) HTH, [144:156) HTh, [156] Nm */
cudaError_t launch_reduce_partials(const double* partials, int n_partials, double* out, cudaStream_t st);
cudaError_t launch_set_frame(UpdateCtrl* c, cudaStream_t st);   /* frame from c->x, done = 0 */
cudaError_t launch_l2_flush(void* buf, size_t bytes, cudaStream_t st);

/* lv_downsample.cu: PointCloudProcessor::temporal_downsample and Compensator::voxelgrid_downsample */
struct DownsampleScratch {
    int64_t cap = 0;
    uint32_t *keys = nullptr, *keys_sorted = nullptr, *vals = nullptr, *vals_sorted = nullptr;
    uint8_t* flags = nullptr;
    int32_t* sel = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    void* box = nullptr;           /* GridBox */
    int* count = nullptr;
    int* h_count = nullptr;        /* pinned: [0] output count, [1] overflow flag */
};
cudaError_t ds_reserve(DownsampleScratch& s, int64_t n);
void ds_free(DownsampleScratch& s);
cudaError_t launch_voxelgrid(DownsampleScratch& s, const float* d_xyz, int64_t n, float leaf, float* d_out, cudaStream_t st,
                             int* launches);
cudaError_t launch_temporal(DownsampleScratch& s, const float* d_xyz, int64_t n, int rate, double min_dist, float* d_out,
                            int32_t* d_idx_out, cudaStream_t st, int* launches);

/* lv_deskew.cu: Compensator::compensate.  d_bad (one int) is set when a timestamp is out of range / out of order */
int deskew_max_states();
cudaError_t launch_deskew(const lv_state32* d_path, int ns, const lv_state32& Xt2, double t_lo, double t_hi,
                          const float* d_xyz, const double* d_t, int64_t n, float* d_out, int* d_bad, cudaStream_t st);

}  // namespace lv
#endif
