/*
 * lv_compensator_host.cpp — the serial parts of the Compensator, on the host as in the reference:
 *   State(const state_ikfom&, double)   src/Objects/State.cpp:40-62
 *   State::operator+=(IMU)              src/Objects/State.cpp:73-75,103-132
 *   Compensator::upsample               src/Modules/Compensator.cpp:73-113
 *   Compensator::get_t2                 src/Modules/Compensator.cpp:55-63
 * A path has a few dozen states per sweep; the per-point work (Compensator::compensate) is the kernel in
 * csrc/lv_deskew.cu.  The arithmetic lives in csrc/lv_deskew.h and is shared with the device.
 */
#include <string.h>

#include "../csrc/lv_deskew.h"
#include "../csrc/lv_host.h"

using namespace lv;

extern "C" {

void lv_state_from_ikfom(const lv_params* p, const double* x, double time, const float a[3], const float w[3],
                         lv_state32* out) {
    memset(out, 0, sizeof(*out));
    const Mat3d R = quat_to_rot(load_quat(x + kRot)), RL = quat_to_rot(load_quat(x + kOffR));
    for (int i = 0; i < 9; ++i) { out->R[i] = (float)R.m[i]; out->RLI[i] = (float)RL.m[i]; }   /* .cast<float>() */
    for (int i = 0; i < 3; ++i) {
        out->pos[i] = (float)x[kPos + i];
        out->vel[i] = (float)x[kVel + i];
        out->bw[i] = (float)x[kBg + i];
        out->ba[i] = (float)x[kBa + i];
        out->tLI[i] = (float)x[kOffT + i];
        out->g[i] = p->initial_gravity[i];          /* State.cpp:21; never overwritten */
        out->a[i] = a[i];
        out->w[i] = w[i];
    }
    out->time = time;
}

void lv_state_add_imu(lv_state32* s, const float a[3], const float w[3], double time) { state_add_imu(s, a, w, time); }

int32_t lv_compensator_upsample(const lv_state32* states, int32_t ns, const float* imu_a, const float* imu_w,
                                const double* imu_t, int32_t ni, lv_state32* out, int32_t cap) {
    if (!states || ns < 1 || !imu_t || ni < 1) return 0;
    int32_t s = 0, u = 0, no = 0;
    lv_state32 cur = states[0];
    while (s < ns - 1) {
        if (no < cap) out[no] = states[s];
        ++no;
        while (u < ni && imu_t[u] < states[s + 1].time) {
            state_add_imu(&cur, imu_a + 3 * u, imu_w + 3 * u, imu_t[u]);
            ++u;
            if (no < cap) out[no] = cur;
            ++no;
        }
        cur = states[s++];      /* Compensator.cpp:98 restarts from the state BEFORE the increment; kept as it is */
    }
    if (u >= ni) u = ni - 1;
    if (no < cap) out[no] = states[ns - 1];
    ++no;
    cur = states[ns - 1];
    while (cur.time < imu_t[ni - 1] && u < ni) {
        state_add_imu(&cur, imu_a + 3 * u, imu_w + 3 * u, imu_t[u]);
        ++u;
        if (no < cap) out[no] = cur;
        ++no;
    }
    return no;
}

void lv_compensator_get_t2(const lv_state32* path, int32_t ns, double t2, lv_state32* out) {
    int32_t s = ns - 1;
    while (s > 0 && t2 < path[s].time) --s;
    *out = path[s];
    const float a[3] = {out->a[0], out->a[1], out->a[2]}, w[3] = {out->w[0], out->w[1], out->w[2]};
    state_add_imu(out, a, w, t2);
}

}  // extern "C"
