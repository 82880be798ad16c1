/* libsonarfe -- B200 (sm_100a) sonar front end: CFAR detector, polar->Cartesian
 * feature cloud, cloud filters, ICP scan matcher.
 *
 * C ABI that replaces the two pybind11 modules of jake3991/sonar-SLAM's
 * bruce_slam package on its per-keyframe hot path:
 *
 *     bruce_slam.cfar   bruce_slam/src/bruce_slam/cpp/cfar.cpp:194-204
 *     bruce_slam.pcl    bruce_slam/src/bruce_slam/cpp/pcl.cpp:176-214
 *
 * plus the numeric body of FeatureExtraction.callback that sits between them
 * (bruce_slam/src/bruce_slam/feature_extraction.py:223-249).  Plain pointers and
 * sizes only.  Every function returns SFE_OK (0) or a negative SFE_ERR_* code and
 * never throws; sfe_last_error() gives the thread's last message.
 *
 * Pointer suffixes:  *_dev  = device memory on the context's GPU (asynchronous on
 * the context's stream; call sfe_sync() or synchronise the stream yourself),
 * *_host = ordinary host memory (the call copies in/out and returns when done).
 * There is no CPU fallback: every entry point runs CUDA kernels.
 */
#ifndef SONARFE_H_
#define SONARFE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFE_VERSION 201 /* round 2: + sfe_fov_select_*, sfe_frontend_set_carry, sfe_frontend_params.flip_lateral, sfe_icp_params.flags bit 1;
                            201: + sfe_icp_params.minimizer / normals_knn (point-to-plane, icp.yaml:18-19) */

#if defined(__GNUC__)
#define SFE_API __attribute__((visibility("default")))
#else
#define SFE_API
#endif

enum {
  SFE_OK = 0,
  SFE_ERR_ARG = -1,      /* bad argument (shape, dtype code, null pointer ...) */
  SFE_ERR_CUDA = -2,     /* a CUDA runtime/driver call failed                  */
  SFE_ERR_CAPACITY = -3, /* a caller-provided output buffer was too small      */
  SFE_ERR_UNSUPPORTED = -4
};

/* CFAR variants, in the order of cfar.cpp:194-204 */
enum { SFE_CFAR_CA = 0, SFE_CFAR_SOCA = 1, SFE_CFAR_GOCA = 2, SFE_CFAR_OS = 3 };
/* image element types */
enum { SFE_U8 = 0, SFE_F32 = 1 };

typedef struct sfe_ctx sfe_ctx; /* one GPU + one stream + scratch memory; not thread-safe */

SFE_API int sfe_version(void);
SFE_API const char *sfe_last_error(void);

/* Create a context on `device`.  own_stream != 0: the library creates its own
 * non-blocking stream (cuda_stream is ignored).  own_stream == 0: all work is enqueued
 * on the caller's cudaStream_t `cuda_stream` (NULL = the legacy default stream; pass
 * e.g. torch's current stream). */
SFE_API int sfe_ctx_create(int device, void *cuda_stream, int own_stream, sfe_ctx **out);
SFE_API void sfe_ctx_destroy(sfe_ctx *ctx);
SFE_API int sfe_sync(sfe_ctx *ctx);
/* blocking device -> host copy on the context's stream (for callers without a CUDA runtime binding) */
SFE_API int sfe_copy_to_host(sfe_ctx *ctx, void *dst_host, const void *src_dev, uint64_t bytes);
/* number of kernels this context has launched so far (bench.py: gpu_launches) */
SFE_API uint64_t sfe_launch_count(const sfe_ctx *ctx);

/* ------------------------------------------------------------------ CFAR
 * Replaces cfar::{ca,soca,goca,os} (cfar.cpp:10,30,53,76) and, with a non-NULL
 * thr pointer, cfar::{ca2,soca2,goca2,os2} (cfar.cpp:98,120,145,170), batched
 * over `n_frames` images [n_frames][R][B] (range bins x beams, beams contiguous).
 * The detection window runs along R for every beam, exactly as in the reference:
 * `train_hs` leading + `train_hs` lagging training cells separated from the cell
 * under test by `guard_hs` guard cells on each side; the first and last
 * train_hs+guard_hs rows of the mask are 0.
 *
 *   dtype        SFE_U8 or SFE_F32.  The reference takes float32 (pybind converts
 *                the node's uint8 image); uint8 is accepted directly and gives the
 *                same result as converting first.
 *   k, tau       OS rank (0-based, < 2*train_hs; ignored otherwise), threshold factor
 *   gate_enable / gate_threshold
 *                fuses the node's `peaks &= img > threshold`
 *                (feature_extraction.py:224) into the detector when gate_enable != 0
 *   mask         uint8 0/1, [n_frames][R][B]                     (may be NULL)
 *   thr          float32 threshold image, [n_frames][R][B]       (may be NULL)
 *   bits         the same mask as a bit plane: uint32 [n_frames][R][ceil(B/32)],
 *                bit (b & 31) of word b >> 5 is beam b             (may be NULL)
 * Results are bit-identical to the reference for every float32 input, integer
 * valued or not (non-integer tiles are re-done with the reference's sequential
 * float32 accumulation order).
 */
SFE_API int sfe_cfar_dev(sfe_ctx *ctx, const void *img_dev, int dtype, int n_frames, int R, int B, int alg,
                 int train_hs, int guard_hs, int k, double tau, int gate_enable, double gate_threshold,
                 uint8_t *mask_dev, float *thr_dev, uint32_t *bits_dev);
SFE_API int sfe_cfar_host(sfe_ctx *ctx, const void *img_host, int dtype, int n_frames, int R, int B, int alg,
                  int train_hs, int guard_hs, int k, double tau, int gate_enable, double gate_threshold,
                  uint8_t *mask_host, float *thr_host);

/* ------------------------------------------------------------------ polar -> Cartesian cloud
 * Replaces feature_extraction.py:231-238: cv2.remap(mask, map_x, map_y, INTER_LINEAR),
 * np.nonzero, pixel -> metres.  `sfe_maps` holds, on the device, the per-geometry sampling
 * table derived from the float32 maps that generate_map_xy builds (feature_extraction.py:134-173;
 * the maps themselves stay host-side Python).  R x B = polar image (range bins x beams),
 * rows x cols = Cartesian image, width/height = its extent in metres (self.width, self.height).
 */
typedef struct sfe_maps sfe_maps;
SFE_API int sfe_maps_create(sfe_ctx *ctx, const float *map_x_host, const float *map_y_host, int rows, int cols, int R,
                            int B, double width, double height, sfe_maps **out);
/* Diagnostic, host only (no context, no GPU): the inverse lists sfe_maps_create uploads -- for every polar cell
 * the Cartesian pixels the detection-driven kernel tests when that cell is a detection -- as a CSR structure:
 * off_out [R*B + 1], idx_out [*n_entries] pixel indices (row * cols + col).  A pixel is listed under the smallest
 * set of its taps that every firing configuration must light (see featx.cu: build_inverse_lists).  Returns
 * SFE_ERR_CAPACITY (with *n_entries set) when idx_capacity is too small. */
SFE_API int sfe_maps_inverse_lists_host(const float *map_x_host, const float *map_y_host, int rows, int cols, int R,
                                        int B, int32_t *off_out, int32_t *idx_out, int64_t idx_capacity,
                                        int64_t *n_entries);
SFE_API void sfe_maps_destroy(sfe_maps *maps);

/* For each of n_frames polar 0/1 masks (bytes [n_frames][R][B], or the bit plane written by
 * sfe_cfar_dev -- pass exactly one, the other NULL) emit the non-zero Cartesian pixels in row-major
 * order: ij[f][i] = (row, col) int32 (np.nonzero order), xy[f][i] = (y_forward_m, x_lateral_m)
 * float32 = float32(points) of feature_extraction.py:238, count[f] = number found.  At most
 * `capacity` points per frame are written (count still reports the true number). */
SFE_API int sfe_cart_points_dev(sfe_ctx *ctx, const sfe_maps *maps, const uint8_t *mask_dev, const uint32_t *bits_dev,
                                int n_frames, int capacity, int32_t *ij_dev, float *xy_dev, int32_t *count_dev);
/* Host-buffer flavour (byte masks); returns SFE_ERR_CAPACITY if a frame overflowed `capacity`. */
SFE_API int sfe_cart_points_host(sfe_ctx *ctx, const sfe_maps *maps, const uint8_t *mask_host, int n_frames,
                                 int capacity, int32_t *ij_host, float *xy_host, int32_t *count_host);

/* ------------------------------------------------------------------ point-cloud filters
 * Clouds are packed: pts [total][dim] float32 (dim = 2: x,y; remove_outlier also takes 3: x,y,z),
 * off [n_clouds + 1] int32 offsets (cloud c = rows off[c] .. off[c+1]-1), n_max = largest cloud.
 * Results are compacted in place of each cloud: out_pts rows off[c] .. off[c]+out_count[c]-1, and
 * out_idx gives, for every surviving point, its row inside its input cloud.
 *
 * sfe_downsample_*      bruce_slam.pcl.downsample (pcl.cpp:128-159): libpointmatcher
 *                       OctreeGridDataPointsFilter{maxSizeByNode = resolution, samplingMethod = 3}:
 *                       quadtree on the cloud's bounding square, one medoid per leaf, leaves in
 *                       depth-first order.  The two-argument overload (points + descriptors,
 *                       pcl.cpp:143) is served by gathering descriptors with out_idx.
 * sfe_remove_outlier_*  bruce_slam.pcl.remove_outlier (pcl.cpp:54-74): PCL RadiusOutlierRemoval:
 *                       keep a point iff at least min_points OTHER points lie within `radius`;
 *                       input order is preserved.
 */
SFE_API int sfe_downsample_dev(sfe_ctx *ctx, const float *pts_dev, const int32_t *off_dev, int n_clouds, int dim,
                               int n_max, float resolution, float *out_pts_dev, int32_t *out_idx_dev,
                               int32_t *out_count_dev);
SFE_API int sfe_remove_outlier_dev(sfe_ctx *ctx, const float *pts_dev, const int32_t *off_dev, int n_clouds, int dim,
                                   int n_max, double radius, int min_points, float *out_pts_dev,
                                   int32_t *out_idx_dev, int32_t *out_count_dev);
/* single-cloud host flavours: n points in, *n_out points out (out buffers hold n entries) */
SFE_API int sfe_downsample_host(sfe_ctx *ctx, const float *pts_host, int n, int dim, float resolution,
                                float *out_pts_host, int32_t *out_idx_host, int32_t *n_out);
SFE_API int sfe_remove_outlier_host(sfe_ctx *ctx, const float *pts_host, int n, int dim, double radius,
                                    int min_points, float *out_pts_host, int32_t *out_idx_host, int32_t *n_out);

/* ------------------------------------------------------------------ nearest-neighbour match
 * bruce_slam.pcl.match(ref, in, knn = 1, max_dist) (pcl.cpp:161-174; libpointmatcher KDTreeMatcher):
 * for every query point the index of the nearest reference point and the SQUARED float32 distance;
 * no reference point within max_dist -> id -1, dist +inf; ties -> lowest reference index.
 * Batched over n_pairs (ref cloud p, query cloud p), both packed [total][2] with CSR offsets;
 * ids / dists are indexed like the query points.  Only knn = 1 (what the reference uses). */
SFE_API int sfe_match_dev(sfe_ctx *ctx, const float *ref_pts_dev, const int32_t *ref_off_dev, const float *in_pts_dev,
                          const int32_t *in_off_dev, int n_pairs, int n_ref_max, float max_dist, int32_t *ids_dev,
                          float *dists_dev);
SFE_API int sfe_match_host(sfe_ctx *ctx, const float *ref_host, int n_ref, const float *in_host, int n_in,
                           float max_dist, int32_t *ids_host, float *dists_host);

/* ------------------------------------------------------------------ ICP scan matcher
 * bruce_slam.pcl.ICP (pcl.cpp:185-213): libpointmatcher PointMatcher<float>::ICP configured by
 * bruce_slam/config/icp.yaml.  The struct carries what that YAML sets; sfe_icp_params_default()
 * fills in the shipped values (the host side parses the YAML, see bruce_slam/pcl.py). */
typedef struct {
  float matcher_max_dist; /* KDTreeMatcher maxDist                      icp.yaml:9   10.0 */
  float outlier_max_dist; /* MaxDistOutlierFilter maxDist (<= 0: off)   icp.yaml:13   3.0 */
  float trim_ratio;       /* TrimmedDistOutlierFilter ratio (< 0: off)  icp.yaml:15   0.8 */
  int max_iterations;     /* CounterTransformationChecker               icp.yaml:24    40 */
  float min_diff_rot;     /* DifferentialTransformationChecker          icp.yaml:26  0.01 */
  float min_diff_trans;   /*                                            icp.yaml:27   0.1 */
  int smooth_length;      /* (0: differential checker off; <= 15)        icp.yaml:28     4 */
  int flags;              /* bit 0: MaxDist filter compares squared distance with maxDist itself
                             bit 1: accumulate the sums over points (reference mean, pair means, cross-covariance) in
                                    float64, order-independent.  DEFAULT (bit clear): every such sum is a SEQUENTIAL
                                    float32 sum in point order, the accumulation order of the CPU oracle
                                    (oracle/icp_ref.c) -- results are bit-identical to it.  The float64 mode is a few
                                    per cent faster and closer to exact arithmetic, but differs from the float32 chain by
                                    up to ~2e-3 m on ill-conditioned scans (tests/test_icp_parity_gpu.py). */
  int minimizer;          /* errorMinimizer: 0 PointToPointErrorMinimizer (shipped, icp.yaml:20)
                             1 PointToPlaneErrorMinimizer force2D (the alternative the YAML keeps commented out,
                               icp.yaml:18-19): per iteration the 3x3 normal equations of sum ((R q + t - r).n)^2
                               linearised in (theta, tx, ty), solved by Cholesky; the step is Rotation2D(theta), (tx, ty) */
  int normals_knn;        /* point-to-plane only: the reference's normals, as a SurfaceNormalDataPointsFilter{knn}
                             in referenceDataPointsFilters would attach them (libpointmatcher default 5; 3..16):
                             smaller-eigenvalue direction of the covariance of each point's knn nearest reference
                             points (itself included) */
} sfe_icp_params;
SFE_API void sfe_icp_params_default(sfe_icp_params *p);

/* per-problem status (ICP.compute returns the text as its message, pcl.cpp:203-211) */
enum {
  SFE_ICP_SUCCESS = 0,      /* "success" */
  SFE_ICP_NO_OUTLIER = 1,   /* ConvergenceError "no outlier to filter" (no source point has a match) */
  SFE_ICP_NO_POINT = 2,     /* ConvergenceError "ErrorMnimizer: no point to minimize" */
  SFE_ICP_NAN_ROT = 3,      /* ConvergenceError "abs rotation norm not a number" */
  SFE_ICP_NAN_TRANS = 4,    /* ConvergenceError "abs translation norm not a number" */
  SFE_ICP_NOT_RIGID = 5,    /* TransformationError: the initial guess is not a rigid transform */
  SFE_ICP_EMPTY_REF = 6,    /* the target cloud is empty */
  SFE_ICP_SKIPPED = 7,      /* front-end pipeline only: fewer than min_points source/target points (slam.py:654-663) */
  SFE_ICP_TOO_LARGE = 8     /* front-end pipeline only: a cloud exceeded the configured capacity */
};
SFE_API const char *sfe_icp_status_message(int status);

/* Solve n_problems scan matches.  Problem p aligns source cloud src_id[p] to target cloud tgt_id[p]
 * (ids NULL: cloud p) from the initial guess guess[p] (3x3 row-major float32, as ICP.compute's third
 * argument); clouds packed [total][2] with CSR offsets, ns_max / nt_max = largest source / target.
 * Outputs per problem: T (3x3 row-major; = guess when status != 0, like pcl.cpp:207-210), iterations
 * made, inlier count of the last iteration (pairs with non-zero weight), status. */
SFE_API int sfe_icp_dev(sfe_ctx *ctx, const sfe_icp_params *prm, const float *src_pts_dev, const int32_t *src_off_dev,
                        const float *tgt_pts_dev, const int32_t *tgt_off_dev, const int32_t *src_id_dev,
                        const int32_t *tgt_id_dev, int n_problems, int ns_max, int nt_max, const float *guess_dev,
                        float *T_dev, int32_t *iters_dev, int32_t *inliers_dev, int32_t *status_dev);
/* One source / target pair from host memory, n_guesses initial guesses (1 = ICP.compute; >1 = the
 * loop of SLAM.compute_icp_with_cov, slam.py:346-358, run as one batch). */
SFE_API int sfe_icp_host(sfe_ctx *ctx, const sfe_icp_params *prm, const float *src_host, int ns, const float *tgt_host,
                         int nt, const float *guess_host, int n_guesses, float *T_host, int32_t *iters_host,
                         int32_t *inliers_host, int32_t *status_host);

/* ------------------------------------------------------------------ batched per-keyframe front end
 * One call for a backlog of frames: FeatureExtraction.callback (feature_extraction.py:196-252)
 * for every frame, then for every frame i the sequential scan match of SLAM
 * (slam.py:607-633,718-776): source = cloud i, target = clouds of frames i-window .. i-1 moved into
 * frame i-1 with the odometry poses and voxel-down-sampled (get_points, slam.py:229-292),
 * guess = between(pose[i-1], pose[i]), then ICP.  Frame 0 of a batch has no target (status
 * SFE_ICP_SKIPPED), frames 1 .. window-1 use the frames available.  Pose-graph optimisation is not
 * part of this library: the SE(2) results feed the caller's (CPU) ISAM2.
 */
typedef struct {
  int R, B;                 /* polar image: range bins x beams (must equal the maps') */
  int cfar_alg, train_hs, guard_hs, rank; /* feature.yaml CFAR/{alg, Ntc/2, Ngc/2, rank} */
  double tau;               /* threshold factor of that variant (CFAR.py:71-121, computed by the host) */
  int gate_enable;          /* feature.yaml filter/threshold: `peaks &= img > threshold` */
  double gate_threshold;
  float resolution;         /* feature.yaml filter/resolution (<= 0: no down-sampling) */
  double outlier_radius;    /* feature.yaml filter/radius */
  int outlier_min_points;   /* feature.yaml filter/min_points (<= 1: no outlier removal) */
  int window;               /* slam.yaml ssm/target_frames (3) */
  float submap_resolution;  /* slam.yaml point_resolution (0.5; <= 0: no down-sampling of the target) */
  int min_points;           /* slam.yaml ssm/min_points (50) */
  sfe_icp_params icp;
  int cap_points;           /* capacity (rows) reserved per frame for its Cartesian cloud; a frame with more
                               detections is truncated and reported as SFE_ICP_TOO_LARGE */
  int cap_source, cap_target; /* largest source / target cloud the scan matcher accepts */
  int flip_lateral;         /* != 0 (default 1, the reference's behaviour): the cloud SLAM matches on is
                               (forward, -lateral).  FeatureExtraction publishes points as xyz = [p0, 0, p1]
                               (feature_extraction.py:182) and the SLAM node reads them back as
                               np.c_[x, -z] = (p0, -p1) (slam_ros.py:169-170), so every keyframe cloud -- scan-match
                               source, window submap, and the clouds sfe_frontend_results_dev exposes -- has its
                               lateral coordinate negated AFTER the feature filters.  0 keeps
                               FeatureExtraction.callback's own (p0, p1) convention. */
} sfe_frontend_params;
SFE_API void sfe_frontend_params_default(sfe_frontend_params *p);

typedef struct sfe_frontend sfe_frontend;
SFE_API int sfe_frontend_create(sfe_ctx *ctx, const sfe_maps *maps, const sfe_frontend_params *params,
                                int max_frames, sfe_frontend **out);
SFE_API void sfe_frontend_destroy(sfe_frontend *fe);
/* frames: uint8 [n_frames][R][B]; poses: float64 [n_frames][3] = odometry (x, y, theta) of every
 * frame in a common frame, always in HOST memory (24 B per frame).  Device flavour: frames already in
 * device memory, asynchronous, results stay on the device
 * (sfe_frontend_results_dev).  Host flavour: copies frames in chunks of `chunk_frames` (<= 0: 256)
 * overlapped with the kernels, copies the results back and returns when they are there:
 * T [n][9] (row-major 3x3, source -> previous frame), iterations, inliers, status (SFE_ICP_*),
 * npoints (size of every frame's filtered cloud; may be NULL). */
SFE_API int sfe_frontend_run_dev(sfe_frontend *fe, const uint8_t *frames_dev, const double *poses_host, int n_frames);
SFE_API int sfe_frontend_results_dev(const sfe_frontend *fe, const float **T, const int32_t **iters,
                                     const int32_t **inliers, const int32_t **status, const float **cloud_xy,
                                     const int32_t **cloud_count, int32_t *cloud_stride);
/* Continuation across calls (the window submap stays resident on the device).  With carry enabled,
 * every call keeps the clouds and odometry poses of its last `window` frames on the device, and the next call on
 * this handle continues the sequence: its first frames are matched against the carried frames (frame 0 included),
 * exactly as if the batches had been one call -- a replay can be fed in pieces of any size without a cold window at
 * every seam.  enable = 0 (the default) drops the carried frames: every call starts cold (frame 0 SFE_ICP_SKIPPED). */
SFE_API int sfe_frontend_set_carry(sfe_frontend *fe, int enable);
/* Optional per-stage device timing (CUDA events on the launch stream around every stage's kernels).
 * get_timing synchronises, adds the intervals recorded since the last call to running totals and
 * returns the totals: stage_ms[SFE_FE_STAGES], stage_calls[SFE_FE_STAGES] (may be NULL). */
enum { SFE_FE_CFAR = 0, SFE_FE_CART = 1, SFE_FE_DOWNSAMPLE = 2, SFE_FE_OUTLIER = 3, SFE_FE_SUBMAP = 4,
       SFE_FE_ICP = 5, SFE_FE_STAGES = 6 };
SFE_API int sfe_frontend_set_timing(sfe_frontend *fe, int enable);
SFE_API int sfe_frontend_get_timing(sfe_frontend *fe, double *stage_ms, int64_t *stage_calls);
SFE_API int sfe_frontend_run_host(sfe_frontend *fe, const uint8_t *frames_host, const double *poses_host,
                                  int n_frames, int chunk_frames, float *T_host, int32_t *iters_host,
                                  int32_t *inliers_host, int32_t *status_host, int32_t *npoints_host);

/* ------------------------------------------------------------------ global-initialisation cost (next row N2)
 * SLAM.get_matching_cost_subroutine1 (slam.py:461-570), the function scipy.shgo minimises in
 * initialize_sequential_scan_matching (slam.py:683-701) and initialize_nonsequential_scan_matching
 * (slam.py:943-961): an occupancy grid of the target cloud at `resolution` (= point_noise / 10), dilated by an
 * elliptical structuring element, and for a candidate transform the NEGATED number of transformed source points
 * that land on an occupied cell.
 *
 * The grid geometry is computed by the caller exactly like slam.py:507-512 (xmin/ymin = cloud minimum minus
 * 2 * point_noise as float32; rows/cols = lengths of the two np.arange calls).  Cell of a point (slam.py:516-517,
 * 556-557), all in float32 like numpy:  r = int(rint((y - ymin) / resolution)), c = int(rint((x - xmin) / resolution)),
 * target cells clipped into the grid, source cells outside the grid not counted.  The structuring element is
 * passed as one column span [se_lo[j], se_hi[j]) per row j of the (2 * dilate_hs + 1)^2 kernel (empty row: lo >= hi)
 * -- for cv2.getStructuringElement(MORPH_ELLIPSE, ...) the caller evaluates OpenCV's span formula (slam.py:523-527).
 * A candidate transform is 6 floats {r00, r01, r10, r11, tx, ty} = sample_transform.matrix().astype(float32)
 * (slam_objects.py:195); points move as x' = fma(y, r01, x * r00) + tx (what numpy's float32 dot does).  */
typedef struct sfe_costmap sfe_costmap;
SFE_API int sfe_costmap_create(sfe_ctx *ctx, const float *target_xy_host, int n_target, float xmin, float ymin,
                               float resolution, int rows, int cols, int dilate_hs, const int32_t *se_lo,
                               const int32_t *se_hi, sfe_costmap **out);
SFE_API void sfe_costmap_destroy(sfe_costmap *cm);
/* the dilated grid as the reference holds it: uint8 [rows][cols], 0 or 255 */
SFE_API int sfe_costmap_grid_host(sfe_ctx *ctx, const sfe_costmap *cm, uint8_t *grid_host);
/* source cloud of the closure (kept on the device between evaluations) */
SFE_API int sfe_costmap_set_source_host(sfe_ctx *ctx, sfe_costmap *cm, const float *source_xy_host, int n_source);
/* cost[k] = -(number of source points on occupied cells under transform k); K candidates in one launch */
SFE_API int sfe_costmap_score_host(sfe_ctx *ctx, const sfe_costmap *cm, const float *transforms_host, int n_candidates,
                                   int32_t *cost_host);
/* device flavour: source cloud, transforms and costs in device memory, asynchronous on the context's stream */
SFE_API int sfe_costmap_score_dev(sfe_ctx *ctx, const sfe_costmap *cm, const float *source_xy_dev, int n_source,
                                  const float *transforms_dev, int n_candidates, int32_t *cost_dev);

/* ------------------------------------------------------------------ loop-closure target pre-filter (next row N3)
 * SLAM.initialize_nonsequential_scan_matching, slam.py:878-899: keep the points of the accumulated target cloud that
 * lie inside the field of view of at least one of the n_frames source keyframes, range and aperture inflated by the
 * keyframe's pose uncertainty.  Per source keyframe k the caller passes (computed on the host exactly like
 * slam.py:883-888):
 *   inv_T[k]          6 floats {r00, r01, r10, r11, tx, ty} of pose.inverse().matrix().astype(float32)
 *                     (Keyframe.transform_points, slam_objects.py:195)
 *   range_bound[k]    translation_std * 5.0 + oculus.max_range                       (float64)
 *   bearing_bound[k]  rotation_std * 5.0 + oculus.horizontal_aperture * 0.5          (float64)
 * pts: float32 [n][2] (global frame); sel[i] = 1 iff for some k the point moved into keyframe k's frame has
 * float32 norm < range_bound[k] and |float32 atan2| < bearing_bound[k]  (the `sel |= sel_i` loop, slam.py:879-895). */
SFE_API int sfe_fov_select_dev(sfe_ctx *ctx, const float *pts_dev, int n, const float *inv_T_dev,
                               const double *range_bound_dev, const double *bearing_bound_dev, int n_frames,
                               uint8_t *sel_dev);
SFE_API int sfe_fov_select_host(sfe_ctx *ctx, const float *pts_host, int n, const float *inv_T_host,
                                const double *range_bound_host, const double *bearing_bound_host, int n_frames,
                                uint8_t *sel_host);

#ifdef __cplusplus
}
#endif
#endif /* SONARFE_H_ */
