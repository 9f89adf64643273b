/*
 * lv_synth.cpp — synthetic reader: seeded "city block" world + LiDAR ray caster.
 *
 * Replaces the ROS subscribers / rosbag input of the reference driver (src/main.cpp:27-39,
 * Accumulator::receive_lidar, src/Modules/Accumulator.cpp:39-48) for tests and benchmarks: the hot
 * path's inputs are (a) a map cloud and (b) deskewed sweeps in the LiDAR frame, which is exactly
 * what this produces (SURVEY.md 8d "concrete synthetic inputs").
 *
 *   world  ground plane z = -1.8 m over an L x L square, axis-aligned buildings on a jittered
 *          40 m grid (8-30 m footprint, 5-20 m tall) that leave a road corridor |y| < 10 m free;
 *          map = one surface sample per 0.2 m lattice cell + N(0, 1 cm) noise along the normal,
 *          L chosen so that the sample count reaches the requested map size exactly
 *   sweep  rings x azimuths beams in firing order (azimuth-major), elevation measured in the
 *          (levelled) body frame, nearest hit of ground / walls, range noise, returns closer
 *          than min_dist or missing are re-drawn so that the sweep size is constant
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/lv_synth.h"
#include "../csrc/lv_host.h"
#include "../csrc/lv_manifold.h"

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double uni(double a, double b) { return a + (b - a) * uni(); }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
};

struct Box { double x0, x1, y0, y1, z1; };   /* footprint + top; bottom = ground */

const double kGroundZ = -1.8;
const double kLattice = 0.2;
const double kRoadHalf = 10.0;

}  // namespace

struct lv_synth_world {
    uint64_t seed;
    double L;
    std::vector<Box> boxes;
    std::vector<float> map;   /* xyz */
};

namespace {

void layout_boxes(lv_synth_world* w) {
    w->boxes.clear();
    Rng rng(w->seed ^ 0xB0B0ull);
    const double half = w->L / 2;
    int n = (int)floor(w->L / 40.0);
    if (n < 3) n = 3;                              /* small test worlds still get walls */
    const double pitch = w->L / n;
    const double scale = pitch >= 40.0 ? 1.0 : pitch / 40.0;
    const double road = kRoadHalf * scale;
    for (int iy = 0; iy < n; ++iy)
        for (int ix = 0; ix < n; ++ix) {
            const double cx = -half + (ix + 0.5) * pitch + rng.uni(-5, 5) * scale;
            const double cy = -half + (iy + 0.5) * pitch + rng.uni(-5, 5) * scale;
            const double sx = rng.uni(8, 30) * scale, sy = rng.uni(8, 30) * scale, h = rng.uni(5, 20);
            Box b;
            b.x0 = cx - sx / 2; b.x1 = cx + sx / 2; b.y0 = cy - sy / 2; b.y1 = cy + sy / 2; b.z1 = kGroundZ + h;
            if (b.y0 < road && b.y1 > -road) continue;   /* keep the road free */
            if (b.x0 < -half + 1 || b.x1 > half - 1 || b.y0 < -half + 1 || b.y1 > half - 1) continue;
            w->boxes.push_back(b);
        }
}

bool inside_footprint(const lv_synth_world* w, double x, double y) {
    for (const Box& b : w->boxes)
        if (x > b.x0 && x < b.x1 && y > b.y0 && y < b.y1) return true;
    return false;
}

void sample_surfaces(lv_synth_world* w) {
    w->map.clear();
    Rng rng(w->seed ^ 0x5A5Aull);
    const double half = w->L / 2;
    const int n = (int)floor(w->L / kLattice);
    auto push = [&](double x, double y, double z) {
        w->map.push_back((float)x); w->map.push_back((float)y); w->map.push_back((float)z);
    };
    for (int iy = 0; iy < n; ++iy)
        for (int ix = 0; ix < n; ++ix) {
            const double x = -half + (ix + 0.5 + rng.uni(-0.4, 0.4)) * kLattice;
            const double y = -half + (iy + 0.5 + rng.uni(-0.4, 0.4)) * kLattice;
            const double z = kGroundZ + 0.01 * rng.normal();
            if (inside_footprint(w, x, y)) continue;
            push(x, y, z);
        }
    for (const Box& b : w->boxes) {
        const int nz = (int)floor((b.z1 - kGroundZ) / kLattice);
        const int nx = (int)floor((b.x1 - b.x0) / kLattice), ny = (int)floor((b.y1 - b.y0) / kLattice);
        for (int iz = 0; iz < nz; ++iz) {
            for (int i = 0; i < nx; ++i)
                for (int side = 0; side < 2; ++side) {
                    const double x = b.x0 + (i + 0.5 + rng.uni(-0.4, 0.4)) * kLattice;
                    const double z = kGroundZ + (iz + 0.5 + rng.uni(-0.4, 0.4)) * kLattice;
                    push(x, (side ? b.y1 : b.y0) + 0.01 * rng.normal(), z);
                }
            for (int i = 0; i < ny; ++i)
                for (int side = 0; side < 2; ++side) {
                    const double y = b.y0 + (i + 0.5 + rng.uni(-0.4, 0.4)) * kLattice;
                    const double z = kGroundZ + (iz + 0.5 + rng.uni(-0.4, 0.4)) * kLattice;
                    push((side ? b.x1 : b.x0) + 0.01 * rng.normal(), y, z);
                }
        }
    }
}

}  // namespace

extern "C" {

lv_synth_world* lv_synth_world_create(uint64_t seed, int64_t m) {
    if (m <= 0) return NULL;
    lv_synth_world* w = new lv_synth_world();
    w->seed = seed;
    w->L = sqrt((double)m * kLattice * kLattice);   /* ground alone */
    if (w->L < 30) w->L = 30;
    for (int it = 0; it < 12; ++it) {
        layout_boxes(w);
        sample_surfaces(w);
        const double have = (double)(w->map.size() / 3);
        if (have >= (double)m && have <= 1.08 * (double)m) break;
        w->L *= sqrt(1.03 * (double)m / have);
    }
    int64_t have = (int64_t)(w->map.size() / 3);
    if (have < m) {   /* tiny maps: pad by re-jittering ground samples */
        Rng rng(seed ^ 0x77ull);
        while (have < m) {
            const double x = rng.uni(-w->L / 2, w->L / 2), y = rng.uni(-w->L / 2, w->L / 2);
            if (inside_footprint(w, x, y)) continue;
            w->map.push_back((float)x); w->map.push_back((float)y); w->map.push_back((float)(kGroundZ + 0.01 * rng.normal()));
            ++have;
        }
    }
    if (have > m) {   /* drop a seeded random subset, keep the order otherwise */
        Rng rng(seed ^ 0x99ull);
        std::vector<uint8_t> drop((size_t)have, 0);
        int64_t todo = have - m;
        while (todo > 0) {
            const int64_t i = (int64_t)(rng.uni() * (double)have);
            if (i < have && !drop[(size_t)i]) { drop[(size_t)i] = 1; --todo; }
        }
        size_t k = 0;
        for (int64_t i = 0; i < have; ++i)
            if (!drop[(size_t)i]) {
                w->map[3 * k] = w->map[3 * i]; w->map[3 * k + 1] = w->map[3 * i + 1]; w->map[3 * k + 2] = w->map[3 * i + 2];
                ++k;
            }
        w->map.resize(3 * (size_t)m);
    }
    return w;
}

void lv_synth_world_destroy(lv_synth_world* w) { delete w; }

int64_t lv_synth_world_map(const lv_synth_world* w, float* out, int64_t cap) {
    if (!w) return 0;
    const int64_t n = (int64_t)(w->map.size() / 3);
    if (out) memcpy(out, w->map.data(), sizeof(float) * 3 * (size_t)(n < cap ? n : cap));
    return n;
}

double lv_synth_world_extent(const lv_synth_world* w) { return w ? w->L : 0; }

void lv_synth_pose(const lv_synth_world* w, double s, const lv_params* p, double* x) {
    const float q0[4] = {0.f, 0.f, 0.f, 1.f};
    double P[LV_DOF * LV_DOF];
    lvh_init_state(*p, q0, x, P);
    const double half = w->L / 2;
    const double px = -half + 20.0 + s;
    const double py = 3.0 * sin(6.283185307179586 * s / 120.0);
    const double dy = 3.0 * 6.283185307179586 / 120.0 * cos(6.283185307179586 * s / 120.0);
    const double yaw = atan2(dy, 1.0);
    x[lv::kPos] = px; x[lv::kPos + 1] = py; x[lv::kPos + 2] = 0.0;
    x[lv::kRot] = 0; x[lv::kRot + 1] = 0; x[lv::kRot + 2] = sin(yaw / 2); x[lv::kRot + 3] = cos(yaw / 2);
}

int64_t lv_synth_sweep(const lv_synth_world* w, const double* x, int rings, int azimuths, double elev_lo_deg,
                       double elev_hi_deg, double min_dist, double range_sigma, uint64_t seed, float* out) {
    using namespace lv;
    if (!w || !x || !out || rings <= 0 || azimuths <= 0) return 0;
    Rng rng(seed ^ 0xC0FFEEull);
    const Mat3d R = quat_to_rot(load_quat(x + kRot));
    const Mat3d RLI = quat_to_rot(load_quat(x + kOffR));
    const Mat3d RLIt = mat3_transpose(RLI);
    const Vec3d tLI = load_vec3(x + kOffT);
    const Vec3d Rt = mat3_apply(R, tLI);
    const double ox = Rt.x + x[kPos], oy = Rt.y + x[kPos + 1], oz = Rt.z + x[kPos + 2];
    const double half = w->L / 2, max_range = 120.0, deg = 0.017453292519943295;
    int64_t k = 0;
    for (int a = 0; a < azimuths; ++a)
        for (int r = 0; r < rings; ++r) {
            double range = -1;
            Vec3d dir_body;
            dir_body.x = 1; dir_body.y = 0; dir_body.z = 0;
            for (int attempt = 0; attempt < 64 && range < 0; ++attempt) {
                int ring = r;
                double az = 6.283185307179586 * ((double)a + (attempt ? rng.uni() * azimuths : 0.0)) / azimuths;
                if (attempt >= 6) ring = (int)(rng.uni() * rings);
                if (attempt >= 16) ring = (int)(rng.uni() * rings * 0.5);   /* the lower half looks at the ground */
                const double el = deg * (rings > 1 ? elev_lo_deg + (elev_hi_deg - elev_lo_deg) * ring / (rings - 1) : elev_lo_deg);
                dir_body.x = cos(el) * cos(az); dir_body.y = cos(el) * sin(az); dir_body.z = sin(el);
                const Vec3d d = mat3_apply(R, dir_body);
                double best = 1e30;
                if (d.z < -1e-9) {
                    const double t = (kGroundZ - oz) / d.z;
                    const double hx = ox + t * d.x, hy = oy + t * d.y;
                    if (t > 0 && fabs(hx) < half && fabs(hy) < half && !inside_footprint(w, hx, hy)) best = t;
                }
                for (const Box& b : w->boxes) {
                    double t0 = 0, t1 = best;
                    const double lo[3] = {b.x0, b.y0, kGroundZ}, hi[3] = {b.x1, b.y1, b.z1};
                    const double o[3] = {ox, oy, oz}, dd[3] = {d.x, d.y, d.z};
                    bool hit = true;
                    for (int c = 0; c < 3 && hit; ++c) {
                        if (fabs(dd[c]) < 1e-12) { if (o[c] < lo[c] || o[c] > hi[c]) hit = false; continue; }
                        double ta = (lo[c] - o[c]) / dd[c], tb = (hi[c] - o[c]) / dd[c];
                        if (ta > tb) std::swap(ta, tb);
                        if (ta > t0) t0 = ta;
                        if (tb < t1) t1 = tb;
                        if (t0 > t1) hit = false;
                    }
                    if (hit && t0 > 1e-6 && t0 < best) best = t0;
                }
                if (best < max_range && best >= min_dist) range = best;
            }
            if (range < 0) range = min_dist + 1.0;   /* pathological pose: keep the count, point is an outlier */
            const double rr = range + range_sigma * rng.normal();
            const Vec3d dl = mat3_apply(RLIt, dir_body);
            out[3 * k] = (float)(dl.x * rr); out[3 * k + 1] = (float)(dl.y * rr); out[3 * k + 2] = (float)(dl.z * rr);
            ++k;
        }
    return k;
}

}  // extern "C"
