"""`bruce_slam.feature_extraction.FeatureExtraction` -- ROS-free mirror of the reference's node class
(bruce_slam/src/bruce_slam/feature_extraction.py:26-252): same attributes, same method names, same
numeric results; the ROS plumbing (rospy params, subscribers, publishers, cv_bridge) is replaced by
plain arguments and return values.

    fe = FeatureExtraction(); fe.init_node(params_dict_or_feature_yaml); points = fe.callback(ping)

`ping` is any object with the OculusPing fields the reference reads: ping_id, range_resolution,
num_ranges, bearings (int16 centi-degrees) and the image as `image` (uint8 [num_ranges, num_beams]) --
see sonar_slam_b200.synth.Ping.  callback() returns (and keeps in `self.points`) the float32 [K,2]
cloud the reference would publish: column 0 forward range (m), column 1 lateral (m); a skipped ping
gives the same single NaN point the reference publishes (feature_extraction.py:201-207).

The polar->Cartesian maps stay host-side numpy/scipy exactly as in the reference
(generate_map_xy, :134-173, re-run only when the geometry changes); everything per ping runs on the
GPU: CFAR with the amplitude gate fused (:223-224), the remap/nonzero/metres step (:231-238),
pcl.downsample and pcl.remove_outlier (:241-249).
"""
import numpy as np
from scipy.interpolate import interp1d

from .. import _lib
from . import pcl
from .CFAR import CFAR

_ALG = {"CA": 0, "SOCA": 1, "GOCA": 2, "OS": 3}


class FeatureExtraction(object):
    """Extract an in-plane 2-D point cloud from polar sonar images with CFAR."""

    def __init__(self):
        # default parameters for CFAR (feature_extraction.py:38-46)
        self.Ntc = 40
        self.Ngc = 10
        self.Pfa = 1e-2
        self.rank = None
        self.alg = "SOCA"
        self.detector = None
        self.threshold = 0
        self.cimg = None

        # default parameters for the point cloud (:48-54)
        self.colormap = "RdBu_r"
        self.pub_rect = True
        self.resolution = 0.5
        self.outlier_filter_radius = 1.0
        self.outlier_filter_min_points = 5
        self.skip = 5

        self.feature_img = None

        # polar -> Cartesian remapping state (:59-70)
        self.res = None
        self.height = None
        self.rows = None
        self.width = None
        self.cols = None
        self.map_x = None
        self.map_y = None
        self.f_bearings = None
        self.to_rad = lambda bearing: bearing * np.pi / 18000
        self.REVERSE_Z = 1
        self.maxRange = None

        self.compressed_images = True
        self.rov_id = ""

        # device side
        self._maps = None
        self._maps_shape = None
        self.points = None
        self.locs = None

    def configure(self):
        """Build the CFAR detector object from the current parameters."""
        self.detector = CFAR(self.Ntc, self.Ngc, self.Pfa, self.rank)

    def init_node(self, params=None, ns="~"):
        """`params`: the dict rosparam would hold (the content of config/feature.yaml) or a path to that
        YAML.  Keys as in feature_extraction.py:86-114 (CFAR/{Ntc,Ngc,Pfa,rank,alg},
        filter/{threshold,resolution,radius,min_points,skip}, compressed_images)."""
        if isinstance(params, str):
            import yaml
            with open(params) as f:
                params = yaml.safe_load(f)
        params = params or {}
        cf, fl = params.get("CFAR", {}), params.get("filter", {})
        self.Ntc = cf.get("Ntc", self.Ntc)
        self.Ngc = cf.get("Ngc", self.Ngc)
        self.Pfa = cf.get("Pfa", self.Pfa)
        self.rank = cf.get("rank", self.rank)
        self.alg = cf.get("alg", "SOCA")
        self.threshold = fl.get("threshold", self.threshold)
        self.resolution = fl.get("resolution", self.resolution)
        self.outlier_filter_radius = fl.get("radius", self.outlier_filter_radius)
        self.outlier_filter_min_points = fl.get("min_points", self.outlier_filter_min_points)
        self.skip = fl.get("skip", self.skip)
        self.compressed_images = params.get("compressed_images", False)
        self.configure()

    def generate_map_xy(self, ping):
        """Sampling maps from the Cartesian image back into (range bin, beam) coordinates; cached per
        geometry.  Same expressions as the reference (:142-173)."""
        _res = ping.range_resolution
        _height = ping.num_ranges * _res
        _rows = ping.num_ranges
        _width = np.sin(self.to_rad(ping.bearings[-1] - ping.bearings[0]) / 2) * _height * 2
        _cols = int(np.ceil(_width / _res))
        if self.res == _res and self.height == _height and self.rows == _rows and self.width == _width \
                and self.cols == _cols:
            return
        self.res, self.height, self.rows, self.width, self.cols = _res, _height, _rows, _width, _cols

        bearings = self.to_rad(np.asarray(ping.bearings, dtype=np.float32))
        f_bearings = interp1d(bearings, range(len(bearings)), kind="linear", bounds_error=False, fill_value=-1,
                              assume_sorted=True)
        self.f_bearings = f_bearings
        XX, YY = np.meshgrid(range(self.cols), range(self.rows))
        x = self.res * (self.rows - YY)
        y = self.res * (-self.cols / 2.0 + XX + 0.5)
        b = np.arctan2(y, x) * self.REVERSE_Z
        r = np.sqrt(np.square(x) + np.square(y))
        self.map_y = np.asarray(r / self.res, dtype=np.float32)
        self.map_x = np.asarray(f_bearings(b), dtype=np.float32)
        self._maps = None  # device table is rebuilt on next use

    def device_maps(self, ctx, num_beams):
        """The geometry's sampling table on the GPU (built lazily, cached)."""
        key = (id(ctx), self.rows, self.cols, num_beams, self.width, self.height)
        if self._maps is None or self._maps_shape != key:
            self._maps = _lib.Maps(ctx, self.map_x, self.map_y, self.rows, num_beams, self.width, self.height)
            self._maps_shape = key
        return self._maps

    def publish_features(self, ping, points):
        """The reference publishes a PointCloud2 here (:175-193).  The cloud is kept and returned as before; the
        message itself (same bytes, no ROS: bruce_slam/conversions.py) is left in `self.feature_msg` for a caller
        that forwards it."""
        self.points = points
        from . import conversions
        self.feature_msg = conversions.feature_msg(ping, points)
        return points

    def ping_image(self, sonar_msg):
        """The polar uint8 image of a ping.  Compressed pings (`compressed_images: True`, a sensor_msgs/CompressedImage
        whose `.data` holds the PNG/JPEG bytes) are decoded on the host exactly like the reference does
        (feature_extraction.py:210-213: cv2.imdecode as BGR, then BGR -> gray); anything else is taken as the decoded
        [num_ranges, num_beams] array (the reference's ros_numpy.image_to_numpy branch, :217)."""
        img = getattr(sonar_msg, "image", None)
        if img is None:
            img = sonar_msg.ping
        data = getattr(img, "data", None)
        if self.compressed_images and isinstance(data, (bytes, bytearray, memoryview)):
            import cv2  # the reference's own decoder; only needed for compressed bags
            buf = np.frombuffer(data, np.uint8)
            dec = cv2.imdecode(buf, cv2.IMREAD_COLOR)
            if dec is None:
                raise ValueError("FeatureExtraction: the compressed ping could not be decoded")
            img = cv2.cvtColor(np.array(dec).astype(np.uint8), cv2.COLOR_BGR2GRAY)
        return np.ascontiguousarray(img)

    def callback(self, sonar_msg):
        if sonar_msg.ping_id % self.skip != 0:
            self.feature_img = None
            nan = np.array([[np.nan, np.nan]])
            return self.publish_features(sonar_msg, nan)

        img = self.ping_image(sonar_msg)
        if img.dtype != np.uint8 or img.ndim != 2:
            raise TypeError("FeatureExtraction.callback: the ping image must be uint8 [num_ranges, num_beams]")
        self.generate_map_xy(sonar_msg)

        # CFAR + amplitude gate (feature_extraction.py:223-224), on the device
        if self.detector is None:
            self.configure()
        prm = self.detector.params[self.alg]
        train_hs, guard_hs, tau = prm[0], prm[1], prm[-1]
        rank = prm[2] if self.alg == "OS" else 0
        ctx = _lib.default_context()
        R, B = img.shape
        mask = np.empty((R, B), np.uint8)
        _lib.check(ctx.lib.sfe_cfar_host(ctx.handle, _lib.ptr(img), 0, 1, R, B, _ALG[self.alg], int(train_hs),
                                         int(guard_hs), int(rank), float(tau), 1, float(self.threshold),
                                         _lib.ptr(mask), None), "FeatureExtraction: cfar")
        # remap + nonzero + metres (:231-238)
        maps = self.device_maps(ctx, B)
        cap = self.rows * self.cols
        ij = np.empty((cap, 2), np.int32)
        xy = np.empty((cap, 2), np.float32)
        cnt = np.zeros(1, np.int32)
        _lib.check(ctx.lib.sfe_cart_points_host(ctx.handle, maps.handle, _lib.ptr(mask), 1, cap, _lib.ptr(ij),
                                                _lib.ptr(xy), _lib.ptr(cnt)), "FeatureExtraction: cart_points")
        k = int(cnt[0])
        self.locs = ij[:k].astype(np.int64)
        points = xy[:k].copy()

        # filters (:241-249)
        if len(points) and self.resolution > 0:
            points = pcl.downsample(points, self.resolution)
        if self.outlier_filter_min_points > 1 and len(points) > 0:
            points = pcl.remove_outlier(points, self.outlier_filter_radius, self.outlier_filter_min_points)
        return self.publish_features(sonar_msg, points)
