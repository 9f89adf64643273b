/*
 * lv_voxel_search.h — exact 5-NN on the device map, host+device.
 *
 * Replaces KD_TREE<Point>::Nearest_Search / Search (include/ikd-Tree/ikd_Tree/ikd_Tree.cpp:
 * 426-461, 1062-1243) as LIMO-Velo uses it: k = 5, max_dist = infinity, result ascending,
 * followed by Plane's gate "5th squared distance < MAX_DIST_PLANE^2" (src/Objects/Plane.cpp:
 * 36-43).  Because every query whose 5th neighbour is not inside that radius is rejected by the
 * gate, the device search is an EXACT 5-NN restricted to the open ball of radius MAX_DIST_PLANE:
 * identical neighbour sets for every point the reference can accept, "fewer than 5" otherwise.
 *
 * Map layout in HBM (built by lv_map_build.cu once per sweep, replacing the pointer-linked
 * kd-tree of ikd_Tree.h:66-89): a pyramid of hashed voxel grids over ONE Morton-sorted point array.
 *   pts[]      float4 (x, y, z, bits(map index)) sorted by the 63-bit Morton code of the finest
 *              voxel coordinates; the voxel of level l (edge c * 2^l) is the code >> 3l, so every
 *              voxel of every level is one contiguous run of pts[]
 *   level l    table_l[]  open-addressing hash, one 32-byte slot (= one DRAM/L2 sector) per
 *                         voxel: {key_lo, key_hi, start, count | halo_start, halo_count, -, -}
 *   level 0    additionally holds a slot (count 0) for every EMPTY voxel adjacent to an occupied
 *              one, and halo[]: "halo buckets" — for every level-0 slot one contiguous run with
 *              the voxel's own points followed by those of its (up to 26) occupied neighbours
 * The coarsest level has an edge >= MAX_DIST_PLANE.
 *
 * Search: at level 0, one thread per query (knn5_level0): ONE probe finds the home voxel and ONE
 * contiguous scan of its halo bucket sees every map point within Chebyshev ring 1, i.e. within the
 * certified radius (edge + distance to the nearest face) of the query.  If the 5th best lies
 * inside that radius the answer is exact and final (the bulk of a sweep).  Otherwise — sparse spot,
 * or a query more than a voxel away from the map — a whole warp searches ring 1 of the coarser
 * levels (knn5_upper), the lanes sharing probes and striding over the voxel runs.  At the coarsest
 * level the certified radius covers the whole search ball, so the answer is always exact.  No tree
 * descent, no data-dependent ring loops: the pointer chasing of the kd-tree becomes a streaming
 * read, paid for with ~30x the map in HBM for the level-0 halo of a surface-like map.
 */
#ifndef LV_VOXEL_SEARCH_H_
#define LV_VOXEL_SEARCH_H_

#include "lv_point_math.h"

#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };
struct uint4 { unsigned int x, y, z, w; };
#endif

#if defined(__CUDACC__)
#define LV_UNROLL_N(n) _Pragma("unroll")
#else
#define LV_UNROLL_N(n)
#endif

#ifndef LV_PROBE_COUNT
#define LV_PROBE_COUNT()      /* tuning build: counts hash probes */
#endif
/* diagnosis build (-DLV_WATCHDOG, tools/diag_hang.sh): every data-dependent loop of the search counts its trips and,
 * past a bound no legitimate input reaches, records (loop id, two values) once and leaves the loop */
#if defined(LV_WATCHDOG) && defined(LV_WATCHDOG_TU) && defined(__CUDA_ARCH__)   /* g_lv_wd: defined by the including .cu */
#define LV_WD_INIT() uint32_t lv_wd_n_ = 0
#define LV_WD(id, a, b)                                                                                   \
    if (++lv_wd_n_ > 4000000u) {                                                                          \
        if (atomicCAS(&g_lv_wd[0], 0ull, (unsigned long long)(id)) == 0ull) {                              \
            g_lv_wd[1] = (unsigned long long)(a); g_lv_wd[2] = (unsigned long long)(b);                    \
            g_lv_wd[3] = blockIdx.x; g_lv_wd[4] = threadIdx.x;                                             \
        }                                                                                                 \
        break;                                                                                            \
    }
#else
#define LV_WD_INIT()
#define LV_WD(id, a, b)
#endif

namespace lv {

enum { kMaxLevels = 4 };

struct VoxelLevel {
    const uint4* table;     /* hash slots, 2 x uint4 each          */
    uint32_t mask;          /* slots - 1                           */
    float cell;             /* voxel edge of this level            */
};

struct VoxelMapView {
    const float4* pts;      /* Morton-sorted points                */
    const float4* halo;     /* halo buckets of level 0             */
    VoxelLevel lv[kMaxLevels];
    int32_t n_levels;
    uint32_t n_points;
    float cell0;            /* finest voxel edge                   */
    float inv_cell0;        /* 1 / cell0 (fp32, used identically for build and query) */
};

#define LV_KEY_BIAS (1 << 20)
#define LV_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

/* biased (non-negative, 21-bit) finest-level voxel coordinate */
LV_HD uint32_t voxel_coord(float v, float inv_cell0) {
    int c = (int)floorf(fmul(v, inv_cell0)) + LV_KEY_BIAS;
    c = c < 0 ? 0 : (c > 0x1FFFFF ? 0x1FFFFF : c);
    return (uint32_t)c;
}
LV_HD uint64_t spread21(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
LV_HD uint32_t compact21(uint64_t x) {
    x &= 0x1249249249249249ull;
    x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
    x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
    x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
    x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
    x = (x ^ (x >> 32)) & 0x1fffffull;
    return (uint32_t)x;
}
/* Morton code of biased voxel coordinates (any level) */
LV_HD uint64_t morton3(uint32_t bx, uint32_t by, uint32_t bz) { return spread21(bx) | (spread21(by) << 1) | (spread21(bz) << 2); }

/* table key of a voxel given its biased coordinates at that level: any injective packing will do
 * (the Morton order only matters for the sort); this one costs a handful of instructions */
LV_HD uint64_t voxel_key(uint32_t bx, uint32_t by, uint32_t bz) { return (uint64_t)bx | ((uint64_t)by << 21) | ((uint64_t)bz << 42); }
/* the key of the level-l voxel containing the point with Morton code m at the finest level */
LV_HD uint64_t voxel_key_from_morton(uint64_t m, int l) {
    return voxel_key(compact21(m) >> l, compact21(m >> 1) >> l, compact21(m >> 2) >> l);
}
LV_HD uint32_t voxel_hash(uint64_t k) {   /* classic spatial hash of the three coordinates + a finaliser */
    uint32_t h = ((uint32_t)k & 0x1FFFFFu) * 73856093u ^ ((uint32_t)(k >> 21) & 0x1FFFFFu) * 19349663u ^
                 ((uint32_t)(k >> 42) & 0x1FFFFFu) * 83492791u;
    h ^= h >> 15;
    h *= 0x2c1b3c6du;
    h ^= h >> 12;
    return h;
}

LV_HD uint4 load_slot(const uint4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
LV_HD float4 load_point(const float4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

/* slot index of the voxel with table key `key` (voxel_key) in level L, or -1 */
LV_HD int voxel_find(const VoxelLevel& L, uint64_t key, uint32_t* start, uint32_t* count) {
    const uint32_t klo = (uint32_t)key, khi = (uint32_t)(key >> 32);
    uint32_t slot = voxel_hash(key) & L.mask;
    /* tables are built with load <= 0.6, so a probe sequence meets an empty slot long before it wraps; the bound
     * only guarantees termination whatever the table holds */
    LV_WD_INIT();
    for (uint32_t probes = 0; probes <= L.mask; ++probes) {
        LV_WD(1, key, L.mask)
        const uint4 e = load_slot(L.table + 2 * (size_t)slot);
        LV_PROBE_COUNT();
        if (e.x == klo && e.y == khi) { *start = e.z; *count = e.w; return (int)slot; }
        if ((e.x & e.y) == 0xFFFFFFFFu) break;
        slot = (slot + 1) & L.mask;
    }
    *count = 0;
    return -1;
}

/* ascending top-5 kept in registers.  id = position in the scanned array (-1 = none). */
struct Top5 {
    float d0, d1, d2, d3, d4;
    int i0, i1, i2, i3, i4;
    float d5;   /* smallest squared distance seen that is NOT in the list (the 6th neighbour so far): feeds the
                 * reuse test of later evaluations (query_reusable) */
};
LV_HD void top5_init(Top5& t, float bound) {
    t.d0 = t.d1 = t.d2 = t.d3 = t.d4 = bound;
    t.i0 = t.i1 = t.i2 = t.i3 = t.i4 = -1;
    t.d5 = INFINITY;
}
/* insert if strictly better than the current 5th (ikd_Tree.cpp:1087: dist < q.top().dist);
 * equal distances keep the earlier candidate in front.                                      */
LV_HD void top5_insert(Top5& t, float d, int id) {
    if (!(d < t.d4)) { t.d5 = d < t.d5 ? d : t.d5; return; }
    t.d5 = t.d4 < t.d5 ? t.d4 : t.d5;   /* the old 5th drops out (a `bound` placeholder is >= every real candidate kept) */
    /* branch-free bubble through the sorted list: every lane of a warp executes the same selects */
    bool s;
    float td; int ti;
    s = d < t.d0; td = s ? t.d0 : d; ti = s ? t.i0 : id; t.d0 = s ? d : t.d0; t.i0 = s ? id : t.i0; d = td; id = ti;
    s = d < t.d1; td = s ? t.d1 : d; ti = s ? t.i1 : id; t.d1 = s ? d : t.d1; t.i1 = s ? id : t.i1; d = td; id = ti;
    s = d < t.d2; td = s ? t.d2 : d; ti = s ? t.i2 : id; t.d2 = s ? d : t.d2; t.i2 = s ? id : t.i2; d = td; id = ti;
    s = d < t.d3; td = s ? t.d3 : d; ti = s ? t.i3 : id; t.d3 = s ? d : t.d3; t.i3 = s ? id : t.i3; d = td; id = ti;
    t.d4 = d; t.i4 = id;   /* d < old d4 is known: whatever falls out of slot 3 (or the new value) is the new 5th */
}

/* scan `n` consecutive points starting at p; ids are base + j */
LV_HD void scan_run(const float4* p, uint32_t n, int base, float gx, float gy, float gz, Top5& t) {
    uint32_t j = 0;
    for (; j + 4 <= n; j += 4) {   /* four independent 16-byte loads in flight, then the insertions */
        const float4 q0 = load_point(p + j), q1 = load_point(p + j + 1), q2 = load_point(p + j + 2), q3 = load_point(p + j + 3);
        top5_insert(t, sq_dist(gx, gy, gz, q0.x, q0.y, q0.z), base + (int)j);
        top5_insert(t, sq_dist(gx, gy, gz, q1.x, q1.y, q1.z), base + (int)j + 1);
        top5_insert(t, sq_dist(gx, gy, gz, q2.x, q2.y, q2.z), base + (int)j + 2);
        top5_insert(t, sq_dist(gx, gy, gz, q3.x, q3.y, q3.z), base + (int)j + 3);
    }
    for (; j < n; ++j) {
        const float4 q = load_point(p + j);
        top5_insert(t, sq_dist(gx, gy, gz, q.x, q.y, q.z), base + (int)j);
    }
}

LV_HD void top5_pop(Top5& t, float bound) {
    t.d0 = t.d1; t.i0 = t.i1;
    t.d1 = t.d2; t.i1 = t.i2;
    t.d2 = t.d3; t.i2 = t.i3;
    t.d3 = t.d4; t.i3 = t.i4;
    t.d4 = bound; t.i4 = -1;
}
/* smallest distance this list still holds or has seen dropped: what a merge leaves behind */
LV_HD float top5_rest(const Top5& t) { return t.d0 < t.d5 ? t.d0 : t.d5; }

/* ---- lane groups: the rare queries level 0 cannot settle are finished by a whole warp ---------- */
struct GroupSerial {   /* host / single lane */
    enum { size = 1 };
    LV_HD static int lane() { return 0; }
    LV_HD static uint32_t bcast(uint32_t v, int src) { (void)src; return v; }
    LV_HD static void merge(Top5& loc, float bound, Top5& out) { out = loc; (void)bound; }
    LV_HD static float min_all(float v) { return v; }
};
#if defined(__CUDACC__)
template <int G>
struct GroupLanes {    /* G consecutive lanes of a warp (G = 8 or 32); every lane of the WARP must call merge() */
    enum { size = G };
    __device__ __forceinline__ static int lane() { return (int)(threadIdx.x & (G - 1)); }
    __device__ __forceinline__ static uint32_t bcast(uint32_t v, int src) {
        return __shfl_sync(0xffffffffu, v, (int)((threadIdx.x & 31u) & ~(unsigned)(G - 1)) + src);
    }
    /* group-wide ascending top-5 of the union of the lanes' lists -> out (identical in all lanes) */
    __device__ __forceinline__ static void merge(Top5& loc, float bound, Top5& out) {
        const int l = lane();
        const unsigned gbase = (threadIdx.x & 31u) & ~(unsigned)(G - 1);
        float od[5];
        int oi[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            float m = loc.d0;
#pragma unroll
            for (int s = 1; s < G; s <<= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, s));
            const unsigned b = (__ballot_sync(0xffffffffu, loc.d0 == m) >> gbase) & (0xffffffffu >> (32 - G));
            const int winner = __ffs(b) - 1;                       /* lowest lane on ties */
            od[k] = m;
            oi[k] = __shfl_sync(0xffffffffu, loc.i0, (int)gbase + winner);
            if (l == winner) top5_pop(loc, bound);
        }
        out.d0 = od[0]; out.d1 = od[1]; out.d2 = od[2]; out.d3 = od[3]; out.d4 = od[4];
        out.i0 = oi[0]; out.i1 = oi[1]; out.i2 = oi[2]; out.i3 = oi[3]; out.i4 = oi[4];
        out.d5 = min_all(top5_rest(loc));   /* heads left after five extractions, and what the lanes dropped earlier */
    }
    __device__ __forceinline__ static float min_all(float v) {
#pragma unroll
        for (int s = 1; s < G; s <<= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, s));
        return v;
    }
};
typedef GroupLanes<32> GroupWarp;
#endif

/* the query inside its home voxel of level l: distances to the six faces, minus `slack` (a few ulp
 * of the largest coordinate: fp32 rounding of the voxel assignment floor(v * inv_cell) can never
 * make a bound optimistic) */
struct HomeGeom {
    float lo[3], hi[3];     /* distance to the low / high face per axis, >= 0 */
    float edge;
    float slack;
};
LV_HD HomeGeom home_geom(const VoxelMapView& m, int l, uint32_t bx0, uint32_t by0, uint32_t bz0, float gx, float gy,
                         float gz) {
    HomeGeom h;
    const float c = m.lv[l].cell;
    const uint32_t b[3] = {bx0, by0, bz0};
    const float g[3] = {gx, gy, gz};
    for (int a = 0; a < 3; ++a) {
        const float o = ((float)(int)((b[a] >> l) << l) - (float)LV_KEY_BIAS) * m.cell0;   /* low corner */
        float v = g[a] - o;
        v = v < 0.f ? 0.f : (v > c ? c : v);
        h.lo[a] = v;
        h.hi[a] = c - v;
    }
    float amax = fabsf(gx) > fabsf(gy) ? fabsf(gx) : fabsf(gy);
    amax = amax > fabsf(gz) ? amax : fabsf(gz);
    h.edge = c;
    h.slack = 2e-6f * (amax + 4.0f);
    return h;
}
/* squared radius certified once ring 1 of the home voxel has been searched: every other map point
 * is at least edge + (distance to the nearest face) away */
LV_HD float certified_d2(const HomeGeom& h) {
    float gap = h.lo[0] < h.hi[0] ? h.lo[0] : h.hi[0];
    gap = gap < h.lo[1] ? gap : h.lo[1]; gap = gap < h.hi[1] ? gap : h.hi[1];
    gap = gap < h.lo[2] ? gap : h.lo[2]; gap = gap < h.hi[2] ? gap : h.hi[2];
    float cert = h.edge + gap - h.slack;
    cert = cert > 0.f ? cert : 0.f;
    return cert * cert;
}
/* squared distance from the query to the neighbour voxel (dx, dy, dz) in {-1,0,1}^3, rounded down */
LV_HD float neighbour_box_d2(const HomeGeom& h, int dx, int dy, int dz) {
    const int d[3] = {dx, dy, dz};
    float s = 0.f;
    for (int a = 0; a < 3; ++a) {
        float v = d[a] < 0 ? h.lo[a] : (d[a] > 0 ? h.hi[a] : 0.f);
        v = d[a] == 0 ? 0.f : v - h.slack;
        v = v > 0.f ? v : 0.f;
        s += v * v;
    }
    return s;
}

/* where the 5 neighbours of a query were found */
enum { kSrcHalo = 0, kSrcPts = 1 };   /* ids index halo[] (level 0) or pts[] (upper levels) */

/*
 * Level 0 in two steps so that a thread block can put ALL its probes and bucket fetches in flight
 * before anything waits on them.
 *   level0_probe   ONE hash probe -> the halo bucket (start, count) of the query's home voxel.
 *                  Level-0 slots exist for every occupied voxel AND for every empty voxel adjacent
 *                  to one, so a query only misses here when it is farther than one voxel from all
 *                  map points (returns false).
 *   level0_scan    one GROUP of lanes per query (8 on the device: 8 x 16 B = one 128-byte line per
 *                  request, so the bucket streams through L1 in full lines; 1 on the host): ONE
 *                  contiguous scan + one merge.  Returns true when `out` (identical in all lanes of
 *                  the group) is final, i.e. certified exact.  `active` = the group has a query with a
 *                  bucket; inactive groups only take part in the collectives.
 */
LV_HD bool level0_probe(const VoxelMapView& m, float gx, float gy, float gz, uint32_t* bstart, uint32_t* bcount) {
    const uint32_t bx0 = voxel_coord(gx, m.inv_cell0), by0 = voxel_coord(gy, m.inv_cell0), bz0 = voxel_coord(gz, m.inv_cell0);
    uint32_t start, count;
    const int slot = voxel_find(m.lv[0], voxel_key(bx0, by0, bz0), &start, &count);
    *bstart = 0;
    *bcount = 0;
    if (slot < 0) return false;
    const uint4 b = load_slot(m.lv[0].table + 2 * (size_t)slot + 1);
    *bstart = b.x;
    *bcount = b.y;
    return true;
}

template <class Grp>
LV_HD bool level0_scan(const VoxelMapView& m, float gx, float gy, float gz, float max_d2, uint32_t bstart, uint32_t bcount,
                       bool active, Top5& out, float* region_d2 = nullptr) {
    Top5 loc;
    top5_init(loc, max_d2);
    if (active) {
        const float4* p = m.halo + bstart;
        const uint32_t n = bcount, step = (uint32_t)Grp::size;
        uint32_t j = (uint32_t)Grp::lane();
        /* eight, then four independent 16-byte loads in flight per lane: a typical bucket (~56 points over 4 lanes)
         * takes two round trips instead of four (K1 is latency-bound: -11 % with the 8-wide step) */
        LV_WD_INIT();
        for (; j + 7 * step < n; j += 8 * step) {
            LV_WD(2, bstart, bcount)
            float4 q[8];
            LV_UNROLL_N(8) for (int u = 0; u < 8; ++u) q[u] = load_point(p + j + (uint32_t)u * step);
            LV_UNROLL_N(8) for (int u = 0; u < 8; ++u)
                top5_insert(loc, sq_dist(gx, gy, gz, q[u].x, q[u].y, q[u].z), (int)(bstart + j + (uint32_t)u * step));
        }
        for (; j + 3 * step < n; j += 4 * step) {   /* four independent 16-byte loads in flight per lane */
            LV_WD(3, bstart, bcount)
            const float4 q0 = load_point(p + j), q1 = load_point(p + j + step), q2 = load_point(p + j + 2 * step),
                         q3 = load_point(p + j + 3 * step);
            top5_insert(loc, sq_dist(gx, gy, gz, q0.x, q0.y, q0.z), (int)(bstart + j));
            top5_insert(loc, sq_dist(gx, gy, gz, q1.x, q1.y, q1.z), (int)(bstart + j + step));
            top5_insert(loc, sq_dist(gx, gy, gz, q2.x, q2.y, q2.z), (int)(bstart + j + 2 * step));
            top5_insert(loc, sq_dist(gx, gy, gz, q3.x, q3.y, q3.z), (int)(bstart + j + 3 * step));
        }
        for (; j < n; j += step) {
            LV_WD(4, bstart, bcount)
            const float4 q = load_point(p + j);
            top5_insert(loc, sq_dist(gx, gy, gz, q.x, q.y, q.z), (int)(bstart + j));
        }
    }
    Grp::merge(loc, max_d2, out);
    if (!active) return false;
    const uint32_t bx0 = voxel_coord(gx, m.inv_cell0), by0 = voxel_coord(gy, m.inv_cell0), bz0 = voxel_coord(gz, m.inv_cell0);
    const float cert = certified_d2(home_geom(m, 0, bx0, by0, bz0, gx, gy, gz));
    if (region_d2) *region_d2 = cert;     /* every map point outside the bucket is at least this far (squared) */
    return out.d4 <= cert;                /* out.d4 <= max_d2 always */
}

/* strided scan: lane `first` of `step` lanes takes points first, first+step, ... */
LV_HD void scan_run_strided(const float4* p, uint32_t n, int base, uint32_t first, uint32_t step, float gx, float gy,
                            float gz, Top5& t) {
    for (uint32_t j = first; j < n; j += step) {
        const float4 q = load_point(p + j);
        top5_insert(t, sq_dist(gx, gy, gz, q.x, q.y, q.z), base + (int)j);
    }
}

/*
 * The queries level 0 could not settle, executed by a whole warp on the device (one lane on the
 * host): ring-1 search on ONE coarser level, chosen so that it is conclusive.
 *   bound0  squared 5th distance level 0 found among real points (an upper bound of the answer), or
 *           max_d2 if it found fewer than 5.  The level is the first whose edge >= sqrt(bound0): every
 *           point that can still matter lies within ring 1 of the query's home voxel there.  With
 *           bound0 = max_d2 that is the last level (edge >= search radius).
 *   the lanes probe the 27 voxels (skipping those farther than the bound), a warp scan of the counts
 *   flattens all their points into one index range that the lanes stride over with independent
 *   loads (binary search of the prefix sums by shuffles), then one merge.
 * Ids in `out` index pts[].
 */
template <class Grp>
LV_HD void knn5_upper(const VoxelMapView& m, float gx, float gy, float gz, float max_d2, float bound0, Top5& out,
                      float* region_d2 = nullptr) {
    const uint32_t bx0 = voxel_coord(gx, m.inv_cell0), by0 = voxel_coord(gy, m.inv_cell0), bz0 = voxel_coord(gz, m.inv_cell0);
    const int last = m.n_levels - 1;
    int l = last > 0 ? 1 : 0;
    float covered;     /* squared radius ring 1 of level l is guaranteed to contain */
    {
        float amax = fabsf(gx) > fabsf(gy) ? fabsf(gx) : fabsf(gy);
        amax = amax > fabsf(gz) ? amax : fabsf(gz);
        const float slack = 2e-6f * (amax + 4.0f);
        while (l < last && !(bound0 <= (m.lv[l].cell - slack) * (m.lv[l].cell - slack))) ++l;
        covered = (m.lv[l].cell - slack) * (m.lv[l].cell - slack);
    }
    /* strictly above bound0 so that the points level 0 saw are found again; 10 % farther than needed (as far as
     * the level covers) so that the answer comes with a margin to the nearest point NOT in it: later evaluations
     * of the same sweep reuse it while the iterate moves less than that margin (query_reusable) */
    float bound = nextafterf(bound0, INFINITY);
    {
        float wide = bound0 * 1.21f;
        wide = wide < covered ? wide : covered;
        bound = bound > wide ? bound : wide;
    }
    bound = bound < max_d2 ? bound : max_d2;
    if (region_d2) *region_d2 = bound < covered ? bound : covered;
    const VoxelLevel& L = m.lv[l];
    const HomeGeom h = home_geom(m, l, bx0, by0, bz0, gx, gy, gz);
    const int hx = (int)(bx0 >> l), hy = (int)(by0 >> l), hz = (int)(bz0 >> l);
    Top5 loc;
    top5_init(loc, bound);
#if defined(__CUDA_ARCH__)
    if (Grp::size == 32) {
        const int lane = Grp::lane();
        uint32_t s = 0, cnt = 0;
        if (lane < 27) {
            const int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
            const int cx = hx + dx, cy = hy + dy, cz = hz + dz;
            if (cx >= 0 && cy >= 0 && cz >= 0 && neighbour_box_d2(h, dx, dy, dz) < bound)
                if (voxel_find(L, voxel_key((uint32_t)cx, (uint32_t)cy, (uint32_t)cz), &s, &cnt) < 0) cnt = 0;
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t excl = incl - cnt;
        LV_WD_INIT();
        for (uint32_t j0 = 0; j0 < total; j0 += 128) {           /* four independent loads in flight per lane */
            LV_WD(5, total, l)
            uint32_t idx[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t j = j0 + 32u * (uint32_t)u + (uint32_t)lane;
                ok[u] = j < total;
                /* first voxel whose inclusive prefix exceeds j: binary search over the 32 lanes */
                int lo = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const uint32_t v = __shfl_sync(0xffffffffu, incl, lo + step - 1);
                    if (v <= j) lo += step;
                }
                lo = lo > 31 ? 31 : lo;
                const uint32_t sv = __shfl_sync(0xffffffffu, s, lo), ev = __shfl_sync(0xffffffffu, excl, lo);
                idx[u] = sv + (j - ev);
            }
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = ok[u] ? load_point(m.pts + idx[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ok[u]) top5_insert(loc, sq_dist(gx, gy, gz, q[u].x, q[u].y, q[u].z), (int)idx[u]);
        }
        Grp::merge(loc, bound, out);
        return;
    }
#endif
    for (int n = 0; n < 27; ++n) {   /* single-lane form of the same search */
        const int dx = n % 3 - 1, dy = (n / 3) % 3 - 1, dz = n / 9 - 1;
        const int cx = hx + dx, cy = hy + dy, cz = hz + dz;
        uint32_t s, cnt;
        if (cx < 0 || cy < 0 || cz < 0 || !(neighbour_box_d2(h, dx, dy, dz) < bound)) continue;
        if (voxel_find(L, voxel_key((uint32_t)cx, (uint32_t)cy, (uint32_t)cz), &s, &cnt) >= 0) scan_run(m.pts + s, cnt, (int)s, gx, gy, gz, loc);
    }
    Grp::merge(loc, bound, out);
}

/*
 * Reuse of a query's neighbours by a later evaluation of the SAME sweep (the iterate moved, the map did not).
 *   ref    (gx, gy, gz, lb): the world position the neighbours were searched from and lb, a lower bound (metres,
 *          rounded down) of the distance from there to every map point that is not one of the five
 *          (outsider_bound()).
 *   g      the query's world position now; q[k], id[k] the five stored neighbours.
 * A map point outside the five is now at least lb - |g - ref| away, so if the farthest of the five is closer
 * than that (with room for the fp32 rounding of every computed distance) the exact search would return the
 * same five: they only need their distances recomputed and their order restored.  Returns false (search
 * again) on any doubt: fewer than five, a tie among the new distances, the 5th beyond the search radius.
 */
LV_HD float outsider_bound(float d5_sq, float region_sq) {
    const float m = d5_sq < region_sq ? d5_sq : region_sq;
    const float r = fsqrt(m > 0.f ? m : 0.f);
    return r * (1.0f - 1e-6f);
}
LV_HD bool query_reusable(const float ref[4], const float g[3], const float (*q)[3], const int* id, float max_d2,
                          Top5& out) {
    if (!(ref[3] > 0.f)) return false;
    float d[5];
    int o[5];
    for (int k = 0; k < 5; ++k) {
        if (id[k] == -1) return false;
        d[k] = sq_dist(g[0], g[1], g[2], q[k][0], q[k][1], q[k][2]);
        o[k] = id[k];
    }
    /* 9-comparator sorting network for 5 keys */
#define LV_CE(a, b) { if (d[b] < d[a]) { const float td = d[a]; d[a] = d[b]; d[b] = td; const int ti = o[a]; o[a] = o[b]; o[b] = ti; } }
    LV_CE(0, 1) LV_CE(3, 4) LV_CE(2, 4) LV_CE(2, 3) LV_CE(0, 3) LV_CE(0, 2) LV_CE(1, 4) LV_CE(1, 3) LV_CE(1, 2)
#undef LV_CE
    if (d[0] == d[1] || d[1] == d[2] || d[2] == d[3] || d[3] == d[4]) return false;
    if (!(d[4] < max_d2)) return false;
    const float dx = g[0] - ref[0], dy = g[1] - ref[1], dz = g[2] - ref[2];
    const float moved = fsqrt(dx * dx + dy * dy + dz * dz);
    const float far5 = fsqrt(d[4]);
    const float need = far5 + moved + 1e-5f * (far5 + moved + ref[3]) + 1e-5f;
    if (!(need < ref[3])) return false;
    out.d0 = d[0]; out.d1 = d[1]; out.d2 = d[2]; out.d3 = d[3]; out.d4 = d[4];
    out.i0 = o[0]; out.i1 = o[1]; out.i2 = o[2]; out.i3 = o[3]; out.i4 = o[4];
    out.d5 = INFINITY;
    return true;
}

}  // namespace lv
#endif
