/*
 * lv_synth.h — C ABI of liblv_synth.so: the synthetic reader of tests, benchmarks and the limovelo_synth driver.
 *
 * Replaces the ROS subscribers / rosbag input of the reference driver (src/main.cpp:27-39,
 * Accumulator::receive_lidar, src/Modules/Accumulator.cpp:39-48).  It is a library of its own (plain C++, no CUDA)
 * so that a process which only needs inputs — the CPU reference arm of bench.py — loads no product code.  The
 * library also carries the YAML reader (lv_default_params, lv_params_from_yaml; same source as the product's).
 */
#ifndef LV_SYNTH_H_
#define LV_SYNTH_H_

#include "limovelo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic reader (replaces the ROS subscribers of src/main.cpp:27-39) ---- */
typedef struct lv_synth_world lv_synth_world;
/* seeded "city-block" world whose surface sampling holds exactly m map points              */
lv_synth_world* lv_synth_world_create(uint64_t seed, int64_t m);
void lv_synth_world_destroy(lv_synth_world* w);
int64_t lv_synth_world_map(const lv_synth_world* w, float* xyz_out, int64_t cap);
double lv_synth_world_extent(const lv_synth_world* w);
/* pose of the sensor platform at arc-length s along the world's road (state layout above)   */
void lv_synth_pose(const lv_synth_world* w, double s, const lv_params* p, double* x_out);
/* ray-cast one sweep of `rings` x `azimuths` beams (elevations elev_lo..elev_hi degrees)
 * from state x; writes exactly rings*azimuths points in the LiDAR frame, firing order.       */
int64_t lv_synth_sweep(const lv_synth_world* w, const double* x, int rings, int azimuths,
                       double elev_lo_deg, double elev_hi_deg, double min_dist, double range_sigma,
                       uint64_t seed, float* xyz_out);

#ifdef __cplusplus
}
#endif
#endif /* LV_SYNTH_H_ */
