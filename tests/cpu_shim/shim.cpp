/*
 * tests/cpu_shim/shim.cpp — TEST-ONLY host build of the product's host+device headers.
 *
 * The device math of liblimovelo_b200.so lives in limo-velo_b200/csrc/ headers that compile for
 * both nvcc and g++ (lv_voxel_search.h, lv_point_math.h, lv_ieskf.h, ...).  This file instantiates
 * them on the CPU (ExecSerial = one "thread", no barriers) so that `pytest -m "not gpu"` can check
 * the very code the kernels run against the oracle, in a container without a GPU.
 *
 * It is NOT part of the product, is not linked into liblimovelo_b200.so and is not reachable from
 * the C ABI; the product has no CPU path.  The hash-table builder below is a plain host loop that
 * produces the same layout lv_map_build.cu produces on the device.
 */
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../limo-velo_b200/csrc/lv_ieskf.h"
#include "../../limo-velo_b200/csrc/lv_voxel_search.h"

using namespace lv;

struct ShimMap {
    std::vector<float4> pts;
    std::vector<uint4> table;
    VoxelMapView view;
};

struct ShimParams {
    int32_t max_iter;
    int32_t estimate_extrinsics;
    double max_dist_plane;
    float planes_threshold;
    float voxel_size;
    double R, D;
    double limits[23];
};

extern "C" {

ShimMap* shim_map_create(const float* xyz, int64_t m, float cell) {
    ShimMap* sm = new ShimMap();
    const float inv = 1.0f / cell;
    std::vector<uint64_t> keys(m);
    for (int64_t i = 0; i < m; ++i)
        keys[i] = voxel_key(voxel_coord(xyz[3 * i], inv), voxel_coord(xyz[3 * i + 1], inv), voxel_coord(xyz[3 * i + 2], inv));
    std::vector<uint32_t> order(m);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    sm->pts.resize(m);
    int64_t heads = 0;
    for (int64_t j = 0; j < m; ++j) {
        const uint32_t s = order[j];
        float4 p;
        p.x = xyz[3 * s]; p.y = xyz[3 * s + 1]; p.z = xyz[3 * s + 2];
        int32_t si = (int32_t)s;
        memcpy(&p.w, &si, 4);
        sm->pts[j] = p;
        if (j == 0 || keys[order[j - 1]] != keys[s]) ++heads;
    }
    uint32_t slots = 1024;
    while (slots < 2 * heads) slots <<= 1;
    uint4 empty; empty.x = empty.y = 0xFFFFFFFFu; empty.z = empty.w = 0;
    sm->table.assign(slots, empty);
    const uint32_t mask = slots - 1;
    for (int64_t j = 0; j < m;) {
        const uint64_t key = keys[order[j]];
        int64_t e = j + 1;
        while (e < m && keys[order[e]] == key) ++e;
        uint32_t slot = voxel_hash(key) & mask;
        while (!((sm->table[slot].x & sm->table[slot].y) == 0xFFFFFFFFu)) slot = (slot + 1) & mask;
        sm->table[slot].x = (uint32_t)key; sm->table[slot].y = (uint32_t)(key >> 32);
        sm->table[slot].z = (uint32_t)j; sm->table[slot].w = (uint32_t)(e - j);
        j = e;
    }
    sm->view.pts = sm->pts.data();
    sm->view.table = sm->table.data();
    sm->view.mask = mask;
    sm->view.n_points = (uint32_t)m;
    sm->view.cell = cell;
    sm->view.inv_cell = inv;
    return sm;
}
void shim_map_destroy(ShimMap* m) { delete m; }

static void search_setup(const ShimParams* p, float cell, float* max_d2, double* gate, int* max_ring) {
    *gate = p->max_dist_plane * p->max_dist_plane;
    float f = (float)*gate;
    if ((double)f < *gate) f = nextafterf(f, INFINITY);
    *max_d2 = f;
    *max_ring = (int)ceil(p->max_dist_plane / (double)cell);
    if (*max_ring < 1) *max_ring = 1;
}

/* per-point outputs like lv_match_all; rows (n x 13) optional */
void shim_match_all(const ShimMap* sm, const double* x, const ShimParams* p, const float* xyz, int64_t n,
                    uint8_t* valid, int32_t* nn_idx, float* nn_sqd, float* plane, float* dist, float* g_world,
                    double* rows) {
    Frame fr;
    make_frame(x, &fr);
    float max_d2; double gate; int max_ring;
    search_setup(p, sm->view.cell, &max_d2, &gate, &max_ring);
    for (int64_t i = 0; i < n; ++i) {
        float g[3];
        rt_apply(fr.lidar_to_world, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], g);
        Top5 t;
        knn5(sm->view, g[0], g[1], g[2], max_d2, max_ring, t);
        float abcd[4] = {0, 0, 0, 0}, d = 0;
        double row[12] = {0}, h = 0;
        bool chosen = false;
        const int ids[5] = {t.i0, t.i1, t.i2, t.i3, t.i4};
        const float ds[5] = {t.d0, t.d1, t.d2, t.d3, t.d4};
        if (t.i4 >= 0 && (double)t.d4 < gate) {
            float q[5][3];
            for (int k = 0; k < 5; ++k) { q[k][0] = sm->pts[ids[k]].x; q[k][1] = sm->pts[ids[k]].y; q[k][2] = sm->pts[ids[k]].z; }
            chosen = plane_fit(q, p->planes_threshold, abcd);
            if (chosen) {
                d = plane_dist(abcd, g);
                jacobian_row(fr, g, abcd, d, p->estimate_extrinsics != 0, row, &h);
            } else {
                abcd[0] = abcd[1] = abcd[2] = abcd[3] = 0;
            }
        }
        if (valid) valid[i] = chosen;
        if (g_world) memcpy(g_world + 3 * i, g, sizeof(g));
        for (int k = 0; k < 5; ++k) {
            int32_t orig = -1;
            if (t.i4 >= 0) memcpy(&orig, &sm->pts[ids[k]].w, 4);
            if (nn_idx) nn_idx[5 * i + k] = orig;
            if (nn_sqd) nn_sqd[5 * i + k] = t.i4 >= 0 ? ds[k] : INFINITY;
        }
        if (plane) memcpy(plane + 4 * i, abcd, sizeof(abcd));
        if (dist) dist[i] = d;
        if (rows) {
            for (int k = 0; k < 12; ++k) rows[13 * i + k] = chosen ? row[k] : 0.0;
            rows[13 * i + 12] = chosen ? h : 0.0;
        }
    }
}

static void reduce_rows(const double* rows, const uint8_t* valid, int64_t n, double* HTH, double* HTh, int64_t* nm) {
    for (int i = 0; i < 144; ++i) HTH[i] = 0;
    for (int i = 0; i < 12; ++i) HTh[i] = 0;
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        const double* r = rows + 13 * i;
        for (int a = 0; a < 12; ++a) {
            for (int b = 0; b < 12; ++b) HTH[a * 12 + b] += r[a] * r[b];
            HTh[a] += r[a] * r[12];
        }
        ++cnt;
    }
    *nm = cnt;
}

/* whole update with the product's ieskf_begin / ieskf_step run serially; returns status */
int shim_update(const ShimMap* sm, double* x, double* P, const ShimParams* p, const float* xyz, int64_t n,
                IterLog* logs, int32_t* n_evals) {
    UpdateCtrl* c = new UpdateCtrl();
    IeskfWork* w = new IeskfWork();
    memset(c, 0, sizeof(*c));
    memcpy(c->x, x, sizeof(double) * kStateLen);
    memcpy(c->P, P, sizeof(double) * kN * kN);
    IeskfParams prm;
    prm.R = p->R; prm.D = p->D; prm.max_iter = p->max_iter; prm.estimate_extrinsics = p->estimate_extrinsics;
    for (int i = 0; i < kN; ++i) prm.limits[i] = p->limits[i];
    ExecSerial ex;
    ieskf_begin(ex, c);
    std::vector<double> rows(13 * n);
    std::vector<uint8_t> valid(n);
    for (int e = 0; e <= p->max_iter && !c->done; ++e) {
        shim_match_all(sm, c->x, p, xyz, n, valid.data(), nullptr, nullptr, nullptr, nullptr, nullptr, rows.data());
        reduce_rows(rows.data(), valid.data(), n, w->HTH, w->HTh, &w->n_matches);
        ieskf_step(ex, prm, c, w);
    }
    memcpy(x, c->x, sizeof(double) * kStateLen);
    if (c->status == 0) memcpy(P, c->P, sizeof(double) * kN * kN);
    *n_evals = c->n_evals;
    memcpy(logs, c->logs, sizeof(IterLog) * kMaxEvals);
    const int st = c->status;
    delete c;
    delete w;
    return st;
}

/* the 23x23 step alone, measurement supplied */
int shim_step(const double* x_prop, const double* P_prop, const double* x_cur, const ShimParams* p, const double* HTH,
              const double* HTh, int64_t nm, int iter, int t_in, double* dx_out, double* x_new, double* P_out,
              int32_t* done) {
    UpdateCtrl* c = new UpdateCtrl();
    IeskfWork* w = new IeskfWork();
    memset(c, 0, sizeof(*c));
    memcpy(c->x_prop, x_prop, sizeof(double) * kStateLen);
    memcpy(c->P_prop, P_prop, sizeof(double) * kN * kN);
    memcpy(c->x, x_cur, sizeof(double) * kStateLen);
    c->iter = iter; c->t = t_in;
    IeskfParams prm;
    prm.R = p->R; prm.D = p->D; prm.max_iter = p->max_iter; prm.estimate_extrinsics = p->estimate_extrinsics;
    for (int i = 0; i < kN; ++i) prm.limits[i] = p->limits[i];
    memcpy(w->HTH, HTH, sizeof(double) * 144);
    memcpy(w->HTh, HTh, sizeof(double) * 12);
    w->n_matches = nm;
    ExecSerial ex;
    ieskf_step(ex, prm, c, w);
    memcpy(dx_out, c->logs[0].dx, sizeof(double) * kN);
    memcpy(x_new, c->x, sizeof(double) * kStateLen);
    if (c->done && c->status == 0) memcpy(P_out, c->P, sizeof(double) * kN * kN);
    *done = c->done;
    const int st = c->status;
    delete c;
    delete w;
    return st;
}

void shim_boxplus(double* x, const double* d) { state_boxplus(x, d); }
void shim_boxminus(const double* x, const double* y, double* d) { state_boxminus(x, y, d); }
void shim_plane_fit(const float* pts5, float thr, float* abcd, int* ok) {
    float q[5][3];
    memcpy(q, pts5, sizeof(q));
    *ok = plane_fit(q, thr, abcd) ? 1 : 0;
}
int shim_sizeof_iterlog() { return (int)sizeof(IterLog); }

}  // extern "C"
