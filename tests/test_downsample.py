"""Downsamplers of SURVEY 8f row 3: temporal decimator (bit-exact: integer / compare work) and PCL voxel grid
(fp32 centroids, leaves in PCL cell order; bit-exact against the oracle's restatement, which sums in input order —
PCL's own unstable sort leaves that order unspecified)."""
import numpy as np
import pytest


def test_oracle_downsamplers_basic_properties(O, scene_xaloc):
    xyz = scene_xaloc.sweep
    idx = O.temporal_downsample(xyz, 4, 4.0)
    assert len(idx) > 0 and ((idx + 1) % 4 == 0).all()
    assert (np.linalg.norm(xyz[idx].astype(np.float64), axis=1) > 4.0 - 1e-5).all()
    assert (O.temporal_downsample(xyz, 1, 0.0) == np.arange(len(xyz))).all()
    c = O.voxelgrid_downsample(xyz, 0.5)
    assert 0 < len(c) < len(xyz)
    # every centroid lies inside the bounding box, and the grid is idempotent enough: a second pass keeps the count close
    assert (c.min(0) >= xyz.min(0) - 1e-4).all() and (c.max(0) <= xyz.max(0) + 1e-4).all()
    assert len(O.voxelgrid_downsample(c, 0.5)) <= len(c)
    one = O.voxelgrid_downsample(xyz[:1], 0.5)
    assert one.shape == (1, 3) and (one == xyz[:1]).all()
    with pytest.raises(ValueError):
        O.voxelgrid_downsample(xyz, 1e-5)                                # PCL: cell index would overflow


@pytest.mark.gpu
@pytest.mark.parametrize("rate,min_dist", [(4, 4.0), (1, 0.0), (32, 10.0), (3, 1e9)])
def test_temporal_downsample_bit_exact(lv, O, scene_xaloc, rate, min_dist):
    loc = lv.Localizer(scene_xaloc.prm)
    xyz = scene_xaloc.sweep
    got, gi = loc.temporal_downsample(xyz, rate, min_dist)
    ri = O.temporal_downsample(xyz, rate, min_dist)
    assert (gi == ri).all() and (got == xyz[ri]).all()
    loc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("leaf", [0.5, 0.2, 2.0])
def test_voxelgrid_downsample_matches_oracle(lv, O, scene_xaloc, scene_kitti, leaf):
    for sc in (scene_xaloc, scene_kitti):
        loc = lv.Localizer(sc.prm)
        for xyz in (sc.sweep, sc.sweep[:1], sc.sweep[:777][::-1].copy(), sc.map[:50000]):
            got = loc.voxelgrid_downsample(xyz, leaf)
            ref = O.voxelgrid_downsample(xyz, leaf)
            assert got.shape == ref.shape
            assert (got == ref).all()
        # PCL warns and hands the input on when the cell index would overflow (voxel_grid.hpp applyFilter); so does the library
        passed = loc.voxelgrid_downsample(sc.sweep, 1e-5)
        assert passed.shape == sc.sweep.shape and (passed == sc.sweep).all()
        loc.close()
