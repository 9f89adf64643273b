/*
 * Modules.hpp — LIMO-Velo's module boundaries on top of the C ABI (header-only C++).
 *
 * The reference driver (src/main.cpp:52-130) talks to four singletons: Accumulator, Compensator,
 * Localizator, Mapper.  These classes keep their names and the method names / argument meaning of the
 * calls that sit on the per-sweep path, so the tick of main.cpp reads the same; the bodies forward to
 * liblimovelo_b200.so.  One `Context` (= one lv_handle) replaces the singletons' hidden coupling
 * (use-ikfom.cpp:18-19 reaches Localizator::getInstance() / Mapper::getInstance()).
 *
 *   reference                                      here
 *   Points = std::deque<Point> (Common.hpp:231)    lv::Points = std::vector<lv::Point> (xyz + time)
 *   State (Objects.hpp:97-137, fp32)               lv::State  = flat state_ikfom layout (26 doubles) + time
 *   Mapper::add(Points&, t, downsample)            Mapper::add(...)          -> lv_map_build / lv_map_add
 *   Mapper::exists() / size()                      same                      -> lv_map_exists / lv_map_size
 *   Localizator::correct(const Points&, t)         same                      -> lv_correct
 *   Localizator::propagate_to(t)                   propagate(IMU) per sample -> lv_predict
 *   Localizator::latest_state()                    same                      -> lv_get_state
 *   Compensator::compensate / downsample           identity hooks (input is already deskewed)
 *   Accumulator::receive_* / get_*                 plain buffers fed by the synthetic / rosbag reader
 */
#ifndef LIMOVELO_B200_MODULES_HPP_
#define LIMOVELO_B200_MODULES_HPP_

#include <deque>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/limovelo_b200.h"

namespace lv {

struct Point { float x, y, z; double time; };
typedef std::vector<Point> Points;
struct IMU { double a[3]; double w[3]; double time; };
typedef std::deque<IMU> IMUs;

struct State {
    double x[LV_STATE_LEN];
    double time;
    const double* pos() const { return x; }
    const double* rot() const { return x + 3; }       /* quaternion x, y, z, w */
};

inline void check(lv_status s, const char* what) {
    if (s != LV_OK) throw std::runtime_error(std::string(what) + ": " + lv_last_error());
}

/* one sequence on one GPU: owns the handle the modules share */
class Context {
   public:
    explicit Context(const lv_params& p) : params(p) { check(lv_create(&params, &h), "lv_create"); }
    explicit Context(const std::string& yaml) {
        lv_default_params(&params);
        check(lv_params_from_yaml(yaml.c_str(), &params), "lv_params_from_yaml");
        check(lv_create(&params, &h), "lv_create");
    }
    ~Context() { lv_destroy(h); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    lv_handle h = nullptr;
    lv_params params;
};

class Mapper {
   public:
    explicit Mapper(Context& c) : ctx(c) {}
    double last_map_time = -1;
    bool exists() { return lv_map_exists(ctx.h) != 0; }                       /* Mapper.cpp:36-38 */
    int size() { return (int)lv_map_size(ctx.h); }                            /* Mapper.cpp:32-34 */
    /* Mapper.cpp:22-30: build on the first call, Add_Points (0.2 m voxel rule if downsample) afterwards */
    void add(const Points& points, double time, bool downsample = false) {
        if (points.empty()) return;
        std::vector<float> xyz(3 * points.size());
        for (size_t i = 0; i < points.size(); ++i) { xyz[3 * i] = points[i].x; xyz[3 * i + 1] = points[i].y; xyz[3 * i + 2] = points[i].z; }
        lv_status s = exists() ? lv_map_add(ctx.h, xyz.data(), (int64_t)points.size(), downsample ? 1 : 0)
                               : lv_map_build(ctx.h, xyz.data(), (int64_t)points.size());
        check(s, "Mapper::add");
        last_map_time = time;
    }
    /* main.cpp:99-105 without leaving the GPU: the sweep Localizator::correct has just used (still resident on the device)
     * is transformed by the corrected state and merged (lv_map_add_last_sweep); nothing is copied to or from the host */
    void add_corrected_sweep(double time, bool downsample = true) {
        check(lv_map_add_last_sweep(ctx.h, downsample ? 1 : 0), "Mapper::add_corrected_sweep");
        last_map_time = time;
    }

   private:
    Context& ctx;
};

class Localizator {
   public:
    explicit Localizator(Context& c) : ctx(c) {}
    double last_time_integrated = -1;
    double last_time_updated = -1;
    bool initialized = false;
    /* Localizator.cpp:119-127,135-153 */
    void initialize(const float q_imu[4], double t) {
        check(lv_init_state(ctx.h, q_imu), "Localizator::initialize");
        last_time_integrated = t;
        initialized = true;
    }
    /* Localizator.cpp:159-173: one IMU sample */
    void propagate(const IMU& imu) {
        if (last_time_integrated < 0) last_time_integrated = imu.time;
        check(lv_predict(ctx.h, imu.a, imu.w, imu.time - last_time_integrated), "Localizator::propagate");
        last_time_integrated = imu.time;
    }
    /* Localizator.cpp:59-75 over an explicit IMU list (the Accumulator lookup is the caller's) */
    void propagate_to(const IMUs& imus, double t) {
        for (const IMU& imu : imus) propagate(imu);
        if (!imus.empty()) { IMU last = imus.back(); last.time = t; propagate(last); }
    }
    /* the same with every IMU sample in ONE device launch (lv_propagate_device): the filter state stays in HBM between the
     * previous update and the next one */
    void propagate_to_device(const IMUs& imus, double t) {
        if (imus.empty()) return;
        std::vector<double> a, w, dt;
        double last = last_time_integrated < 0 ? imus.front().time : last_time_integrated;
        for (const IMU& imu : imus) {
            for (int k = 0; k < 3; ++k) { a.push_back(imu.a[k]); w.push_back(imu.w[k]); }
            dt.push_back(imu.time - last);
            last = imu.time;
        }
        for (int k = 0; k < 3; ++k) { a.push_back(imus.back().a[k]); w.push_back(imus.back().w[k]); }   /* Localizator.cpp:70-74: the last sample up to t */
        dt.push_back(t - last);
        check(lv_propagate_device(ctx.h, a.data(), w.data(), dt.data(), (int32_t)dt.size()), "Localizator::propagate_to_device");
        last_time_integrated = t;
    }
    /* Localizator.cpp:23-27.  Returns the number of h-evaluations (0 when the map is empty). */
    int correct(const Points& points, double time) {
        if (points.empty()) return 0;
        std::vector<float> xyz(3 * points.size());
        for (size_t i = 0; i < points.size(); ++i) { xyz[3 * i] = points[i].x; xyz[3 * i + 1] = points[i].y; xyz[3 * i + 2] = points[i].z; }
        int32_t n_evals = 0;
        const lv_status s = lv_correct(ctx.h, xyz.data(), (int64_t)points.size(), time, logs, &n_evals, nullptr, nullptr);
        if (s == LV_EMPTY_MAP) return 0;                                       /* Localizator.cpp:24 */
        if (s != LV_TOO_FEW_MATCHES) check(s, "Localizator::correct");
        last_time_updated = time;
        return n_evals;
    }
    State latest_state() {                                                     /* Localizator.cpp:77-98 */
        State s;
        check(lv_get_state(ctx.h, s.x, nullptr), "Localizator::latest_state");
        s.time = last_time_updated >= 0 ? last_time_updated : last_time_integrated;
        return s;
    }
    lv_iter_log logs[LV_MAX_EVALS];

   private:
    Context& ctx;
};

/* Deskew (lv_compensate) and voxel-grid downsample (lv_voxelgrid_downsample) on the GPU. */
class Compensator {
   public:
    explicit Compensator(Context& c) : ctx(c) {}
    /* Compensator::compensate(states, Xt2, points) (Compensator.cpp:123-146): `path` as Compensator::upsample
     * leaves it (lv_compensator_upsample), `Xt2` from lv_compensator_get_t2; points sorted by time */
    Points compensate(const std::vector<lv_state32>& path, const lv_state32& Xt2, const Points& sweep) {
        if (sweep.empty()) return Points();                                    /* Compensator.cpp:24 */
        std::vector<float> xyz(3 * sweep.size()), out(3 * sweep.size());
        std::vector<double> t(sweep.size());
        for (size_t i = 0; i < sweep.size(); ++i) {
            xyz[3 * i] = sweep[i].x; xyz[3 * i + 1] = sweep[i].y; xyz[3 * i + 2] = sweep[i].z;
            t[i] = sweep[i].time;
        }
        if (lv_compensate(ctx.h, path.data(), (int32_t)path.size(), &Xt2, xyz.data(), t.data(), (int64_t)sweep.size(),
                          out.data()) != LV_OK)
            throw std::runtime_error(lv_last_error());
        Points res = sweep;                                                    /* attributes ride along (RotTransl.cpp:43-48) */
        for (size_t i = 0; i < res.size(); ++i) { res[i].x = out[3 * i]; res[i].y = out[3 * i + 1]; res[i].z = out[3 * i + 2]; }
        return res;
    }
    Points compensate(const Points& sweep) { return sweep; }                   /* a sweep that arrives deskewed (synthetic reader) */
    /* Compensator::downsample -> voxelgrid_downsample (Compensator.cpp:115-118,148-163): centroids carry xyz only */
    Points downsample(const Points& sweep, float downsample_prec) {
        if (sweep.empty()) return Points();
        std::vector<float> xyz(3 * sweep.size()), out(3 * sweep.size());
        for (size_t i = 0; i < sweep.size(); ++i) { xyz[3 * i] = sweep[i].x; xyz[3 * i + 1] = sweep[i].y; xyz[3 * i + 2] = sweep[i].z; }
        int64_t m = 0;
        if (lv_voxelgrid_downsample(ctx.h, xyz.data(), (int64_t)sweep.size(), downsample_prec, out.data(), &m) != LV_OK)
            return sweep;                                                      /* PCL returns the input when it refuses */
        Points res((size_t)m);
        for (int64_t i = 0; i < m; ++i) res[(size_t)i] = Point{out[3 * i], out[3 * i + 1], out[3 * i + 2], 0.0};
        return res;
    }
    Points downsample(const Points& sweep) { return sweep; }                   /* a sweep that arrives downsampled */

   private:
    Context& ctx;
};

/* Plain time-ordered buffers (Accumulator.hpp:61-74), fed by a reader instead of ROS callbacks. */
class Accumulator {
   public:
    void receive_lidar(const Points& sweep) { points = sweep; }                /* Accumulator.cpp:39-48 */
    void receive_imu(const IMU& imu) { imus.push_back(imu); }
    IMUs get_imus(double t1, double t2) {
        IMUs out;
        for (const IMU& i : imus) if (i.time > t1 && i.time <= t2) out.push_back(i);
        return out;
    }
    Points points;
    IMUs imus;
};

}  // namespace lv
#endif
