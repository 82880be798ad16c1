"""N4: the PointCloud2 hand-off between feature extraction and SLAM, byte layout and sign convention, without ROS
(mirror of utils/conversions.py:240-243,297-308, feature_extraction.py:181-190, slam_ros.py:169-170)."""
import struct

import numpy as np
import pytest

from sonar_slam_b200.bruce_slam import conversions as cv


def test_pointcloud_xyz32_layout_is_create_cloud_xyz32():
    pts = np.array([[1.5, 0.0, -2.25], [3.0, 0.0, 4.0]])
    m = cv.n2r(pts, "PointCloudXYZ")
    assert (m.height, m.width, m.point_step, m.row_step, m.is_bigendian, m.is_dense) == (1, 2, 12, 24, False, False)
    assert [(f.name, f.offset, f.datatype, f.count) for f in m.fields] == [("x", 0, 7, 1), ("y", 4, 7, 1), ("z", 8, 7, 1)]
    assert m.data == struct.pack("<6f", 1.5, 0.0, -2.25, 3.0, 0.0, 4.0)      # what pc2.create_cloud packs
    back = cv.r2n(m)
    assert back.dtype == np.float64 and np.array_equal(back, pts)
    mi = cv.n2r(np.c_[pts, [7.0, 9.0]], "PointCloudXYZI")
    assert mi.point_step == 16 and mi.fields[3].name == "i" and np.array_equal(cv.r2n(mi)[:, 3], [7.0, 9.0])
    with pytest.raises(NotImplementedError):
        cv.n2r(pts, "Image")


def test_feature_handoff_sign_convention_and_skipped_frame():
    class Ping:
        class header:
            stamp = 12.5
    pts = np.array([[10.0, 2.0], [11.0, -3.0]], np.float32)
    m = cv.feature_msg(Ping, pts)
    assert m.header.stamp == 12.5 and m.header.frame_id == "base_link"
    assert np.array_equal(cv.r2n(m), [[10, 0, 2], [11, 0, -3]])              # [p0, 0, p1]
    assert np.array_equal(cv.keyframe_points(m), [[10, -2], [11, 3]])        # (x, -z)
    # a skipped ping publishes one NaN point (feature_extraction.py:206-208); ros_numpy drops it
    nan = cv.feature_msg(Ping, np.array([[np.nan, np.nan]]))
    assert nan.width == 1 and len(cv.keyframe_points(nan)) == 0


def test_oculus_ping_gamma():
    class Fire:
        gamma = 127.5
    class Msg:
        _type = "sonar_oculus/OculusPing"
        ping = np.array([[0, 64], [128, 255]], np.uint8)
        fire_msg = Fire
    out = cv.r2n(Msg)
    assert out.dtype == np.float32 and np.allclose(out, 255.0 * (Msg.ping / 255.0) ** 2.0, atol=1e-4)
