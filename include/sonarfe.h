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

#define SFE_VERSION 100

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

#ifdef __cplusplus
}
#endif
#endif /* SONARFE_H_ */
