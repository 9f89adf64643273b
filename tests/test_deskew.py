"""Compensator (deskew) row of SURVEY 8f: State arithmetic, upsample / get_t2 on the host, compensate on the GPU.

The oracle side (oracle/lv_oracle.cpp, "Deskew") restates State.cpp / Compensator.cpp line by line; the reference's own
Compensator needs ROS + PCL + Eigen and holds no test vectors, so this row is "parity unpinned" (DESIGN.md).
Host functions share libm with the oracle and must agree bit for bit; the kernel uses CUDA's sinf / cosf and is
held to 2e-5 m (fp32 ulp of a 100 m coordinate is 8e-6 m)."""
import numpy as np
import pytest


def make_case(lv, O, sc, n_states=4, imu_hz=400.0, seed=3, n_pts=None, gyro=0.3):
    """a plausible 0.1 s sweep: KF states every ~33 ms, IMU at 400 Hz, points time-stamped along the sweep"""
    rng = np.random.default_rng(seed)
    t1, t2 = 10.0, 10.1
    st_times = np.linspace(t1 - 0.012, t2 - 0.004, n_states)
    # Compensator::path: IMU samples from the first state on (get_imus(states.front().time, t2)) plus the first one after t2
    imu_t = np.arange(st_times[0] + 0.2 / imu_hz, t2 + 1.5 / imu_hz, 1.0 / imu_hz)
    imu_a = (np.array([0.3, -0.2, 9.8]) + rng.normal(0, 0.05, (len(imu_t), 3))).astype(np.float32)
    imu_w = (np.array([0.02, -0.01, gyro]) + rng.normal(0, 0.01, (len(imu_t), 3))).astype(np.float32)
    states_p, states_o = [], []
    for k, ts in enumerate(st_times):
        x = sc.world.pose(8.0 + 15.0 * (ts - t1), sc.prm).copy()
        x[14:17] = [15.0, 0.3, -0.1]                                    # vel
        x[17:20] = [1e-3, -2e-3, 5e-4]                                  # bg
        x[20:23] = [0.02, 0.01, -0.03]                                  # ba
        j = int(np.searchsorted(imu_t, ts))                             # Accumulator::get_next_imu
        states_p.append(lv.state_from_ikfom(sc.prm, x, ts, imu_a[j], imu_w[j]))
        states_o.append(O.state_from_ikfom(x, ts, imu_a[j], imu_w[j], sc.prm.initial_gravity[:]))
    xyz = sc.sweep if n_pts is None else sc.sweep[:n_pts]
    t = np.linspace(t1, t2, len(xyz))
    return dict(t1=t1, t2=t2, states_p=states_p, states_o=states_o, imu_a=imu_a, imu_w=imu_w, imu_t=imu_t, xyz=xyz, t=t)


def same(a, b):
    return a.as_tuple() == b.as_tuple()


def test_state_arithmetic_matches_oracle_bitwise(lv, O, scene_xaloc):
    c = make_case(lv, O, scene_xaloc)
    for sp, so in zip(c["states_p"], c["states_o"]):
        assert same(sp, so)
    rng = np.random.default_rng(0)
    sp, so = c["states_p"][0], c["states_o"][0]
    for k in range(200):
        a = rng.normal(0, 3, 3).astype(np.float32) + np.float32([0, 0, 9.8])
        w = rng.normal(0, 0.5, 3).astype(np.float32) if k % 7 else np.zeros(3, np.float32)   # the |w| < 1e-7 branch too
        tn = sp.time + float(rng.uniform(1e-4, 2e-2))
        sp, so = lv.state_add_imu(sp, a, w, tn), O.state_add_imu(so, a, w, tn)
        assert same(sp, so), k
    R = np.array(sp.R[:], np.float64).reshape(3, 3)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-3                     # 200 fp32 Rodrigues steps stay near SO(3)


@pytest.mark.parametrize("n_states,imu_hz", [(4, 400.0), (2, 100.0), (7, 1000.0)])
def test_upsample_and_get_t2_match_oracle_bitwise(lv, O, scene_xaloc, n_states, imu_hz):
    c = make_case(lv, O, scene_xaloc, n_states=n_states, imu_hz=imu_hz)
    pp = lv.compensator_upsample(c["states_p"], c["imu_a"], c["imu_w"], c["imu_t"])
    po = O.upsample(c["states_o"], c["imu_a"], c["imu_w"], c["imu_t"])
    assert len(pp) == len(po) > n_states
    assert all(same(a, b) for a, b in zip(pp, po))
    times = [s.time for s in pp]
    assert times[0] <= c["t1"] and times[-1] >= c["t2"]                 # the path surrounds the sweep
    assert same(lv.compensator_get_t2(pp, c["t2"]), O.get_t2(po, c["t2"]))
    # oracle compensate == per-point restatement in numpy-free form: count and order are preserved
    out = O.compensate(po, O.get_t2(po, c["t2"]), c["xyz"][:500], c["t"][:500] * 0 + np.linspace(c["t1"], c["t2"], 500))
    assert out.shape == (500, 3) and np.isfinite(out).all()


@pytest.mark.gpu
def test_compensate_matches_oracle(lv, O, scene_xaloc):
    sc = scene_xaloc
    c = make_case(lv, O, sc)
    pp = lv.compensator_upsample(c["states_p"], c["imu_a"], c["imu_w"], c["imu_t"])
    po = O.upsample(c["states_o"], c["imu_a"], c["imu_w"], c["imu_t"])
    x2p, x2o = lv.compensator_get_t2(pp, c["t2"]), O.get_t2(po, c["t2"])
    loc = lv.Localizer(sc.prm)
    got = loc.compensate(pp, x2p, c["xyz"], c["t"])
    ref = O.compensate(po, x2o, c["xyz"], c["t"])
    assert got.shape == ref.shape == c["xyz"].shape
    err = np.abs(got - ref).max(axis=1)
    assert err.max() < 2e-5, err.max()
    assert (err == 0).mean() > 0.5                                       # most points are bit-identical
    # the sweep really was skewed: 15 m/s over 0.1 s
    moved = np.linalg.norm(got - c["xyz"], axis=1)
    assert 0.5 < moved.max() < 5.0 and moved[-1] < 0.2                   # points stamped at t2 barely move
    # device-resident variant, in place
    n = len(c["xyz"])
    d_xyz, d_t = loc.upload(c["xyz"]), loc.upload(c["t"])
    loc.compensate_device(pp, x2p, d_xyz, d_t, n, d_xyz)
    back = np.empty_like(c["xyz"])
    import ctypes as C
    import torch
    torch.cuda.synchronize()
    buf = torch.empty(n * 3, dtype=torch.float32, device="cuda")
    C.cdll.LoadLibrary("libcudart.so").cudaMemcpy(C.c_void_p(buf.data_ptr()), C.c_void_p(d_xyz), C.c_size_t(n * 12), 3)
    assert (buf.cpu().numpy().reshape(-1, 3) == got).all()
    loc.device_free(d_xyz); loc.device_free(d_t)
    # errors the reference asserts on
    with pytest.raises(RuntimeError):
        loc.compensate(pp, x2p, c["xyz"][:100], np.linspace(c["t2"], c["t2"] + 1.0, 100))     # beyond the path
    with pytest.raises(RuntimeError):
        loc.compensate(pp, x2p, c["xyz"][:100], c["t"][:100][::-1].copy())                      # not time-sorted
    loc.close()


@pytest.mark.gpu
def test_compensate_of_a_standing_sensor_is_the_identity(lv, O, scene_xaloc):
    sc = scene_xaloc
    c = make_case(lv, O, sc, gyro=0.0)
    still = []
    for s in c["states_p"]:
        s = lv.State32.from_buffer_copy(s)
        for k in range(3):
            s.vel[k] = 0.0; s.a[k] = 0.0; s.w[k] = 0.0; s.g[k] = 0.0; s.ba[k] = 0.0; s.bw[k] = 0.0
            s.pos[k] = c["states_p"][0].pos[k]
        for k in range(9):
            s.R[k] = c["states_p"][0].R[k]
        still.append(s)
    loc = lv.Localizer(sc.prm)
    got = loc.compensate(still, still[-1], c["xyz"], np.linspace(still[0].time, still[-1].time, len(c["xyz"])))
    assert np.abs(got - c["xyz"]).max() < 2e-4                           # X^-1 X in fp32 at 100 m
    loc.close()
