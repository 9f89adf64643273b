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
 * kd-tree of ikd_Tree.h:66-89):
 *   pts[]    float4 (x, y, z, bits(map index)), sorted by voxel key -> each voxel's points are
 *            one contiguous, 16-byte aligned run (one or two 128-byte lines for a typical voxel)
 *   table[]  open-addressing hash table, one 16-byte slot per occupied voxel:
 *            {key_lo, key_hi, start, count}; empty slot = key 0xFFFFFFFFFFFFFFFF
 * Voxel edge c (lv_params.voxel_size, default 0.5 m).  Search: visit the query's own voxel, then
 * shells of Chebyshev radius r = 1, 2, ... ; a voxel is probed only if its box distance to the
 * query is smaller than the current 5th best (the kd-tree's pruning rule, ikd_Tree.cpp:1098,
 * applied to voxels); stop as soon as the 5th best is within the radius certified by the
 * completed shells, or that radius reaches MAX_DIST_PLANE.
 */
#ifndef LV_VOXEL_SEARCH_H_
#define LV_VOXEL_SEARCH_H_

#include "lv_point_math.h"

#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };
struct uint4 { unsigned int x, y, z, w; };
#endif

namespace lv {

struct VoxelMapView {
    const float4* pts;      /* sorted points */
    const uint4* table;     /* hash slots    */
    uint32_t mask;          /* capacity - 1  */
    uint32_t n_points;
    float cell;             /* voxel edge                        */
    float inv_cell;         /* 1 / cell (fp32, used identically for build and query) */
};

#define LV_KEY_BIAS (1 << 20)
#define LV_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

LV_HD int voxel_coord(float v, float inv_cell) { return (int)floorf(fmul(v, inv_cell)); }

LV_HD uint64_t voxel_key(int ix, int iy, int iz) {
    return ((uint64_t)(uint32_t)(iz + LV_KEY_BIAS) << 42) | ((uint64_t)(uint32_t)(iy + LV_KEY_BIAS) << 21) |
           (uint64_t)(uint32_t)(ix + LV_KEY_BIAS);
}
LV_HD uint32_t voxel_hash(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k;
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

/* returns count (0 if the voxel is empty) and its first point index in *start */
LV_HD uint32_t voxel_lookup(const VoxelMapView& m, int ix, int iy, int iz, uint32_t* start) {
    const uint64_t key = voxel_key(ix, iy, iz);
    const uint32_t klo = (uint32_t)key, khi = (uint32_t)(key >> 32);
    uint32_t slot = voxel_hash(key) & m.mask;
    for (;;) {
        const uint4 e = load_slot(m.table + slot);
        if (e.x == klo && e.y == khi) { *start = e.z; return e.w; }
        if ((e.x & e.y) == 0xFFFFFFFFu) return 0;
        slot = (slot + 1) & m.mask;
    }
}

/* ascending top-5 kept in registers.  id = position in the sorted point array (-1 = none). */
struct Top5 {
    float d0, d1, d2, d3, d4;
    int i0, i1, i2, i3, i4;
};
LV_HD void top5_init(Top5& t, float bound) {
    t.d0 = t.d1 = t.d2 = t.d3 = t.d4 = bound;
    t.i0 = t.i1 = t.i2 = t.i3 = t.i4 = -1;
}
/* insert if strictly better than the current 5th (ikd_Tree.cpp:1087: dist < q.top().dist);
 * equal distances keep the earlier candidate in front.                                      */
LV_HD void top5_insert(Top5& t, float d, int id) {
    if (!(d < t.d4)) return;
    if (d < t.d3) {
        t.d4 = t.d3; t.i4 = t.i3;
        if (d < t.d2) {
            t.d3 = t.d2; t.i3 = t.i2;
            if (d < t.d1) {
                t.d2 = t.d1; t.i2 = t.i1;
                if (d < t.d0) { t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = id; }
                else { t.d1 = d; t.i1 = id; }
            } else { t.d2 = d; t.i2 = id; }
        } else { t.d3 = d; t.i3 = id; }
    } else { t.d4 = d; t.i4 = id; }
}

LV_HD void scan_voxel(const VoxelMapView& m, int ix, int iy, int iz, float gx, float gy, float gz, Top5& t) {
    uint32_t start;
    const uint32_t cnt = voxel_lookup(m, ix, iy, iz, &start);
    for (uint32_t j = 0; j < cnt; ++j) {
        const float4 q = load_point(m.pts + start + j);
        top5_insert(t, sq_dist(gx, gy, gz, q.x, q.y, q.z), (int)(start + j));
    }
}

/* distance from the query (offsets lo/hi to the faces of its home voxel) to the near face of the
 * voxel d steps away along one axis, reduced by `slack` so that fp32 rounding of the voxel
 * assignment (floor(v * inv_cell)) can never make the bound optimistic.                      */
LV_HD float axis_gap(int d, float lo, float hi, float c, float slack) {
    if (d == 0) return 0.f;
    const float g = (d > 0 ? (float)(d - 1) * c + hi : (float)(-d - 1) * c + lo) - slack;
    return g > 0.f ? g : 0.f;
}

/*
 * Exact 5 nearest map points of g within squared radius max_d2 (exclusive).
 * max_ring = ceil(max_dist / c).  Result ascending in t; t.i4 < 0 means "fewer than 5".
 */
LV_HD void knn5(const VoxelMapView& m, float gx, float gy, float gz, float max_d2, int max_ring, Top5& t) {
    top5_init(t, max_d2);
    const float c = m.cell;
    const int cx = voxel_coord(gx, m.inv_cell), cy = voxel_coord(gy, m.inv_cell), cz = voxel_coord(gz, m.inv_cell);
    /* offsets of g inside its voxel, clamped to [0, c] against fp rounding of the floor */
    float lx = gx - (float)cx * c, ly = gy - (float)cy * c, lz = gz - (float)cz * c;
    lx = lx < 0.f ? 0.f : (lx > c ? c : lx);
    ly = ly < 0.f ? 0.f : (ly > c ? c : ly);
    lz = lz < 0.f ? 0.f : (lz > c ? c : lz);
    const float hx = c - lx, hy = c - ly, hz = c - lz;
    float gap = lx < hx ? lx : hx;
    gap = gap < ly ? gap : ly; gap = gap < hy ? gap : hy;
    gap = gap < lz ? gap : lz; gap = gap < hz ? gap : hz;

    /* a few ulp of the largest coordinate involved (2^-23 ~ 1.2e-7 relative) */
    float amax = fabsf(gx) > fabsf(gy) ? fabsf(gx) : fabsf(gy);
    amax = amax > fabsf(gz) ? amax : fabsf(gz);
    const float slack = 2e-6f * (amax + 4.0f);

    scan_voxel(m, cx, cy, cz, gx, gy, gz, t);
    for (int r = 1; r <= max_ring; ++r) {
        /* every unvisited point is at least cert away */
        float cert = (float)(r - 1) * c + gap - slack;
        cert = cert > 0.f ? cert : 0.f;
        const float cert2 = cert * cert;
        if (t.d4 <= cert2) break;          /* also covers cert2 >= max_d2 (d4 <= max_d2 always) */
        for (int dz = -r; dz <= r; ++dz) {
            const float az = axis_gap(dz, lz, hz, c, slack);
            const float az2 = az * az;
            if (!(az2 < t.d4)) continue;
            const bool zface = (dz == -r || dz == r);
            for (int dy = -r; dy <= r; ++dy) {
                const float ay = axis_gap(dy, ly, hy, c, slack);
                const float ayz2 = az2 + ay * ay;
                if (!(ayz2 < t.d4)) continue;
                const bool face = zface || dy == -r || dy == r;
                const int step = face ? 1 : 2 * r;
                for (int dx = -r; dx <= r; dx += step) {
                    const float ax = axis_gap(dx, lx, hx, c, slack);
                    if (!(ayz2 + ax * ax < t.d4)) continue;
                    scan_voxel(m, cx + dx, cy + dy, cz + dz, gx, gy, gz, t);
                }
            }
        }
    }
}

}  // namespace lv
#endif
