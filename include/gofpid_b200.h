/*
 * gofpid_b200.h -- C ABI of libgofpid_b200.so: PyGOFPID's per-frame foreground-extraction
 * hot path (blur -> foreground detection -> morphology -> blob creation) as hand-written
 * CUDA for sm_100a.
 *
 * The reference (qbarthelemy/PyGOFPID v1.1.0) is pure Python and has no FFI of its own; the
 * plug points this ABI replaces are the OpenCV calls of GOFPID.detect
 * (pygofpid/pipeline.py:461-479), each cited on the entry point that replaces it.
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, opaque handles, int status returns (0 = OK, negative = error); the message
 *     of the last error on the calling thread is returned by gofpid_last_error();
 *   - nothing throws or aborts; there is NO CPU fallback: without a CUDA device every
 *     compute entry point returns GOFPID_ERR_CUDA;
 *   - "d_" pointers are device pointers owned by the caller (e.g. torch tensors),
 *     "h_" pointers are host pointers; pitches are in bytes;
 *   - `stream` arguments are cudaStream_t passed as void* (NULL = legacy default stream);
 *     device entry points are asynchronous on that stream unless stated otherwise;
 *   - one handle is single-threaded; distinct handles may be driven from distinct host
 *     threads.
 */
#ifndef GOFPID_B200_H
#define GOFPID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(GOFPID_BUILD)
#define GOFPID_API __attribute__((visibility("default")))
#else
#define GOFPID_API
#endif

/* ---- status codes --------------------------------------------------------------- */
#define GOFPID_OK 0
#define GOFPID_ERR_INVALID (-1)     /* bad argument (maps to ValueError / KeyError upstream)  */
#define GOFPID_ERR_CUDA (-2)        /* CUDA runtime error, or no CUDA device                  */
#define GOFPID_ERR_UNSUPPORTED (-3) /* valid for OpenCV but outside what the kernels cover    */
#define GOFPID_ERR_NOMEM (-4)
#define GOFPID_ERR_OVERFLOW (-5)    /* more blobs / contour points than the handle's capacity */

/* ---- enumerations --------------------------------------------------------------- */
enum { GOFPID_BLUR_NONE = 0, GOFPID_BLUR_GAUSSIAN = 1, GOFPID_BLUR_BOX = 2 };
enum { GOFPID_DETECT_FD = 0, GOFPID_DETECT_MOG2 = 1, GOFPID_DETECT_KNN = 2 };
enum { GOFPID_MORPH_ERODE = 0, GOFPID_MORPH_DILATE = 1 };

#define GOFPID_CONVERT_NONE 0
#define GOFPID_CONVERT_BGR2GRAY 1 /* cv.COLOR_BGR2GRAY */
#define GOFPID_CONVERT_RGB2GRAY 2 /* cv.COLOR_RGB2GRAY */

#define GOFPID_MAX_MORPH_OPS 8
#define GOFPID_MAX_KSIZE 31 /* largest blur kernel side the fused kernel stages in smem */
#define GOFPID_MAX_SE 31    /* largest structuring-element side                          */
#define GOFPID_MOG2_NMIX 5
#define GOFPID_KNN_NSAMPLES 7

/* One morphology step: d['fun'](mask, d['kernel']) of pipeline.py:473-475.  `kernel` is a
 * host pointer to kh*kw bytes, row-major; non-zero = member of the structuring element. */
typedef struct gofpid_morph_op {
    int32_t op; /* GOFPID_MORPH_ERODE | GOFPID_MORPH_DILATE */
    int32_t kh, kw;
    const uint8_t *kernel;
} gofpid_morph_op;

/* Parameterisation of one pipeline.  gofpid_config_default() fills the values the reference
 * uses: GOFPID.__init__ defaults (pipeline.py:168-212) and OpenCV's subtractor defaults
 * (pipeline.py:232,234). */
typedef struct gofpid_config {
    /* blur: pipeline.py:462-467.  NB the reference passes blur['borderType'] positionally, so
     * for cv.GaussianBlur it is sigmaX (default 4.0); the border is always REFLECT_101.  */
    int32_t blur_kind;
    int32_t blur_kw, blur_kh;
    double blur_sigma; /* Gaussian sigma; <= 0: derived from the kernel size as OpenCV does */

    int32_t detector; /* pipeline.py:231-240 */

    /* frame differencing: segmentation.py:36 */
    double fd_threshold;

    /* MOG2: cv.createBackgroundSubtractorMOG2() defaults */
    int32_t mog2_history;
    float mog2_var_threshold;     /* Tb  = 16  */
    float mog2_var_threshold_gen; /* Tg  = 9   */
    float mog2_background_ratio;  /* TB  = 0.9 */
    float mog2_var_init, mog2_var_min, mog2_var_max; /* 15, 4, 75 */
    float mog2_ct;                /* complexity reduction threshold 0.05 */

    /* KNN: cv.createBackgroundSubtractorKNN() defaults */
    int32_t knn_history;      /* 500 */
    int32_t knn_k;            /* 2   */
    float knn_dist2_threshold; /* 400 */
    uint64_t knn_rng_seed;    /* state of the cv::RNG replica; 0 -> 0xffffffff (OpenCV default) */

    /* shared by MOG2 and KNN */
    int32_t detect_shadows; /* 1 */
    int32_t shadow_value;   /* 127 */
    float shadow_tau;       /* 0.5 */

    /* morphology: pipeline.py:473-476 */
    int32_t n_morph;
    gofpid_morph_op morph[GOFPID_MAX_MORPH_OPS];

    /* blob creation: pipeline.py:555-570 (area_min is hard-coded to 100 there) */
    int32_t find_blobs;    /* 0: stop after the mask */
    double blob_area_min;  /* keep contours with cv.contourArea >= this */
    int32_t max_blobs;     /* capacity: connected components per frame  */
    int32_t max_kept;      /* capacity: blobs that pass the area filter */
    int32_t max_points;    /* capacity: contour points per frame (0: 32 (H + W), within 64 Ki .. 1 Mi) */

    /* colour conversion before the blur: pipeline.py:458-459.  Frames have 3 interleaved channels,
     * the path works on their gray value (cv.cvtColor's 15-bit fixed point), computed while the
     * frame is loaded.  The pipeline is created with C = 3. */
    int32_t convert;       /* GOFPID_CONVERT_* */

    /* pipelining; 0 = default.  A synchronous caller (GOFPID.detect: one frame per call) never has
     * more than one frame in flight and may ask for a shallow handle to save memory. */
    int32_t depth;         /* frame slots = frames that can be in flight (1 .. 16; default 12, 4 - 8 with many streams) */
    int32_t mog2_group;    /* frames per MOG2 launch in the clip entry points (1 .. 8; default 3)   */
} gofpid_config;

/* One blob that passed the area filter (the device-side half of GOFPID._create_blob,
 * pipeline.py:572-588).  The host forms m00 = a00/2, m10 = a10/6, m01 = a01/6 as
 * cv.moments does (helpers.py:156,165); a00 keeps its sign (orientation of the trace). */
typedef struct gofpid_blob {
    int32_t first;              /* raster index y*W+x of the component's first pixel   */
    int32_t x0, y0, x1, y1;     /* inclusive bounding box (== cv.boundingRect)         */
    int32_t area_px;            /* pixel count of the 8-connected component            */
    int32_t n_points;           /* contour length (CHAIN_APPROX_NONE)                  */
    int32_t point_offset;       /* index of the first point in the frame's point pool  */
    int32_t nested;             /* device-side flag: enclosed by another blob (such    */
    int32_t reserved;           /* blobs are removed before the list is returned)      */
    int64_t a00, a10, a01;      /* Green sums over the contour polygon (exact)         */
} gofpid_blob;

/* Per-frame result header of one camera stream. */
typedef struct gofpid_frame_info {
    int32_t n_components; /* 8-connected foreground components                         */
    int32_t n_candidates; /* external components passing the bbox pre-filter           */
    int32_t n_kept;       /* contours with contourArea >= area_min (incl. enclosed ones, see reserved[0]) */
    int32_t n_points;     /* contour points in the pool                                */
    int32_t overflow;     /* the frame did not fit the handle's per-slot capacities (bit0: components,
                             bit1: kept blobs, bit2: points; gofpid_pipeline_blobs then re-runs blob
                             creation on the frame's mask with buffers of the size it needs), bit3: internal */
    int32_t frame_index;  /* 1-based count of frames this stream has processed         */
    int32_t reserved[2];  /* [0]: how many of n_kept are enclosed by another blob (not returned)     */
} gofpid_frame_info;

typedef struct gofpid_pipeline gofpid_pipeline; /* opaque */

/* ---- library ------------------------------------------------------------------- */
GOFPID_API const char *gofpid_version(void);
GOFPID_API const char *gofpid_last_error(void);
GOFPID_API int gofpid_device_count(void);
GOFPID_API void gofpid_config_default(gofpid_config *cfg);

/* Page-locked host memory for frames handed to the host entry points (submit_host, process_host,
 * process_host_sequence): copies from it are asynchronous and run at the link's speed.  With
 * write_combined != 0 the buffer is write-combined (fast for the device to read over PCIe, slow for
 * the CPU to read back: fill it, do not read it).  Free with gofpid_host_free. */
GOFPID_API int gofpid_host_alloc(size_t bytes, int write_combined, void **out);
GOFPID_API int gofpid_host_free(void *ptr);
/* Copies n_rows rows of row_bytes bytes between HOST buffers with the library's staging threads (the ones
 * that move pageable frames into page-locked memory inside the host entry points; streaming stores): for
 * capture loops that fill a gofpid_host_alloc buffer themselves -- one thread moves a 1080p frame in
 * 170-240 us, the link needs 40.  n_threads: worker threads beside the caller (0: the caller alone,
 * < 0: default).  Pure host code, no CUDA call; one copy at a time per process. */
GOFPID_API int gofpid_host_copy2d(void *dst, size_t dst_pitch, const void *src, size_t src_pitch,
                                  size_t row_bytes, size_t n_rows, int n_threads);

/* 8.8 fixed-point taps cv.GaussianBlur uses for u8 images (n odd, <= GOFPID_MAX_KSIZE). */
GOFPID_API int gofpid_gaussian_taps_q8(int n, double sigma, uint16_t *h_taps);

/* ---- pipeline: n_streams independent camera streams of identical shape ---------------- */
/* Replaces GOFPID.initialize()'s detector factory (pipeline.py:231-240) plus the state that
 * cv2 keeps inside the subtractor objects.  Allocates the structure-of-arrays model in HBM
 * on `device`.  Frames are u8, C interleaved channels (1 or 3). */
GOFPID_API int gofpid_pipeline_create(const gofpid_config *cfg, int H, int W, int C,
                                      int n_streams, int device, gofpid_pipeline **out);
GOFPID_API int gofpid_pipeline_destroy(gofpid_pipeline *p);
/* Forget all temporal state (as re-creating the cv2 subtractor would). */
GOFPID_API int gofpid_pipeline_reset(gofpid_pipeline *p);

/* One time step for every stream: blur -> detect -> morphology -> blobs
 * (GOFPID.detect, pipeline.py:461-479).  d_frames holds n_streams frames, `stream_stride`
 * bytes apart, rows `pitch` bytes apart.  Asynchronous on `stream`; results stay in HBM
 * until fetched. */
GOFPID_API int gofpid_pipeline_process_device(gofpid_pipeline *p, const uint8_t *d_frames,
                                              size_t pitch, size_t stream_stride, void *stream);
/* Same, from host memory: copies the frames host->device (pinned staging when h_frames is
 * pageable), runs the step, copies result headers + blob lists back and synchronises. */
GOFPID_API int gofpid_pipeline_process_host(gofpid_pipeline *p, const uint8_t *h_frames,
                                            size_t pitch, size_t stream_stride);

/* A clip of n_frames time steps in host memory (frame t at h_frames + t * frame_stride; examples/
 * video.py's read-detect loop over a decoded file) in ONE call: frames are submitted ahead of the
 * one whose blobs are being collected, as a caller of submit_host / select / blobs would do.
 * Blob records and contour points of all frames are appended to h_blobs / h_points (point_offset
 * restarts at 0 for every frame and stream); h_counts[(t * n_streams + z) * 2 + {0, 1}] receives
 * the blob and point count of frame t, stream z.  h_blobs and h_points may be NULL (counts only). */
GOFPID_API int gofpid_pipeline_process_host_sequence(gofpid_pipeline *p, const uint8_t *h_frames,
                                                     int n_frames, size_t frame_stride, size_t pitch,
                                                     size_t stream_stride, gofpid_blob *h_blobs,
                                                     int blob_capacity, int32_t *h_points,
                                                     int point_capacity, int32_t *h_counts);

/* A clip call whose h_blobs / h_points were too small returns GOFPID_ERR_INVALID *after* the whole clip
 * went through (the temporal model is where it should be, h_counts is complete) and keeps the blob
 * lists of every frame of that clip in the handle: size the buffers from h_counts and fetch them here.
 * The lists stay valid until the next clip call on the handle. */
GOFPID_API int gofpid_pipeline_sequence_results(gofpid_pipeline *p, gofpid_blob *h_blobs, int blob_capacity,
                                                int32_t *h_points, int point_capacity);

/* The same for a clip that already is in device memory (frame t at d_frames + t * frame_stride);
 * the detector launches are enqueued on `stream`.  With h_counts == NULL nothing is collected: the
 * call only enqueues the work and the newest `depth` frames stay selectable (gofpid_pipeline_select).
 * Both clip entry points carry MOG2 through gofpid_pipeline_group() consecutive frames per detector
 * launch (a tile's modes stay in shared memory meanwhile: the model is read and written once per
 * group); results are identical to frame-by-frame calls. */
GOFPID_API int gofpid_pipeline_process_device_sequence(gofpid_pipeline *p, const uint8_t *d_frames,
                                                       int n_frames, size_t frame_stride, size_t pitch,
                                                       size_t stream_stride, void *stream,
                                                       gofpid_blob *h_blobs, int blob_capacity,
                                                       int32_t *h_points, int point_capacity,
                                                       int32_t *h_counts);
/* Orders `stream` after everything the handle has enqueued so far -- the detector launches AND the
 * morphology / blob kernels that run on the handle's internal streams -- without blocking the host
 * (cudaStreamWaitEvent on every frame slot in flight).  An event recorded on `stream` afterwards marks
 * the end of all that work: what a caller timing or chaining device-resident clips needs. */
GOFPID_API int gofpid_pipeline_join(gofpid_pipeline *p, void *stream);
/* Frames per detector launch in the clip entry points (1 unless the detector is MOG2). */
GOFPID_API int gofpid_pipeline_group(gofpid_pipeline *p);

/* Asynchronous variant of process_host for pinned host frames: enqueues the host->device copy
 * and the step and returns; several frames are in flight (morphology / labelling / tracing of
 * frames t, t-1, ... overlap the model update of frame t+1).  Results are collected later with
 * the getters below, after gofpid_pipeline_select(p, age) for an earlier frame. */
GOFPID_API int gofpid_pipeline_submit_host(gofpid_pipeline *p, const uint8_t *h_frames,
                                           size_t pitch, size_t stream_stride);
/* Which frame the result getters refer to: age 0 = newest submitted frame (default after every
 * process/submit call), 1 = the one before it, ... up to gofpid_pipeline_depth() - 1. */
GOFPID_API int gofpid_pipeline_select(gofpid_pipeline *p, int age);
/* Frame slots of the handle = frames that can be in flight (submit_host) at once. */
GOFPID_API int gofpid_pipeline_depth(gofpid_pipeline *p);

/* Results of the selected frame (host copies; these wait for that frame's kernels). */
GOFPID_API int gofpid_pipeline_frame_info(gofpid_pipeline *p, int stream_idx,
                                          gofpid_frame_info *h_info);
/* Blobs in the reference's order (descending raster order of the first pixel) and their
 * contour points as (x, y) int32 pairs.  Capacities are in records / points; when they are too
 * small the call returns GOFPID_ERR_INVALID with the sizes it needs in *n_blobs / *n_points.
 * A frame that exceeded the handle's max_blobs / max_kept / max_points is not an error: blob
 * creation is run again on the frame's (still resident) final mask with larger buffers. */
GOFPID_API int gofpid_pipeline_blobs(gofpid_pipeline *p, int stream_idx, gofpid_blob *h_blobs,
                                     int blob_capacity, int32_t *h_points, int point_capacity,
                                     int *n_blobs, int *n_points);
/* foreground_mask_ of pipeline.py:476, (H, W) u8, to host or device memory. */
GOFPID_API int gofpid_pipeline_mask(gofpid_pipeline *p, int stream_idx, uint8_t *dst,
                                    size_t dst_pitch, int dst_is_device);
/* Canonical labels: 0 = background, else 1 + raster index of the component's first pixel. */
GOFPID_API int gofpid_pipeline_labels(gofpid_pipeline *p, int stream_idx, int32_t *dst,
                                      size_t dst_pitch, int dst_is_device);
/* Per-component statistics of the last frame (unordered): first, x0, y0, x1, y1, area_px
 * as 6 int32 per component. */
GOFPID_API int gofpid_pipeline_components(gofpid_pipeline *p, int stream_idx, int32_t *h_stats,
                                          int capacity, int *n_components);

/* Model state, for parity checks and checkpointing (cv2 hides it).  MOG2: dense arrays
 * weight[H][W][5], var[H][W][5], mean[H][W][5][C] (f32) and nmodes[H][W] (u8), host memory. */
GOFPID_API int gofpid_pipeline_mog2_model(gofpid_pipeline *p, int stream_idx, float *h_weight,
                                          float *h_var, float *h_mean, uint8_t *h_nmodes);
/* KNN: samples[H][W][21][C+1], idx[3][H][W], nxt[3][H][W] (u8), counters[3], RNG state. */
GOFPID_API int gofpid_pipeline_knn_model(gofpid_pipeline *p, int stream_idx, uint8_t *h_samples,
                                         uint8_t *h_idx, uint8_t *h_nxt, int32_t *h_counters,
                                         uint64_t *rng_state);
/* FD: the previous (blurred) frame, [H][W][C] u8; returns 1 through *valid once a frame has
 * been seen. */
GOFPID_API int gofpid_pipeline_fd_model(gofpid_pipeline *p, int stream_idx, uint8_t *h_prev,
                                        int *valid);
/* Mean number of live MOG2 modes per pixel of the last frame (feeds the algorithmic-bytes
 * figure of the roofline). */
GOFPID_API int gofpid_pipeline_mog2_mean_modes(gofpid_pipeline *p, double *mean_modes);

/* Timing aid for bench.py: device time (ms, CUDA events on the pipeline's launch stream)
 * spent in each stage kernel group during the last process call:
 * [0] fused blur+detect kernel(s), [1] morphology, [2] labelling+stats+contours,
 * [3] frame tick, [4] whole step.  After a clip call that grouped frames (MOG2): [0] the group's
 * ONE detector launch (gofpid_pipeline_group() frames), [1] morphology of the group's first frame,
 * [2] the rest of the group's stage B, [4] the whole group. */
GOFPID_API int gofpid_pipeline_set_profiling(gofpid_pipeline *p, int enabled);
GOFPID_API int gofpid_pipeline_stage_ms(gofpid_pipeline *p, float *h_ms5);
/* Number of kernels the last process call launched. */
GOFPID_API int gofpid_pipeline_launch_count(gofpid_pipeline *p);
/* Number of kernels launched since creation / the last reset (all entry points). */
GOFPID_API int gofpid_pipeline_launch_total(gofpid_pipeline *p, long long *n_launches);

/* ---- host tail: blob tracking, post-filtering, alarm (GOFPID.detect steps 5-7) ---------------
 * Replaces pipeline.py:590-799 (_track_blob .. _maintain_alarm) for one camera stream.  Pure host code,
 * no CUDA call: usable without a device.  The arithmetic keeps the reference's data types (int16 centres
 * whose differences and squares wrap, int32 ellipse radii, float64 quotients). */
typedef struct gofpid_tracker gofpid_tracker; /* opaque */
typedef struct gofpid_tracker_config {
    int32_t H;                    /* frame rows = length of the two perspective tables                     */
    const double *persp_height;   /* GOFPID._perspective['height'] (pipeline.py:509-547), H values          */
    const double *persp_area;     /* GOFPID._perspective['area'],   H values                                */
    const int32_t *perimeter;     /* post_filter['perimeter'] in pixels: n_perimeter (x, y) pairs           */
    int32_t n_perimeter;
    double tracking_factor;       /* tracking['factor']                                                     */
    int32_t absence_max;          /* tracking['absence_max']                                                */
    int32_t anchor_center;        /* post_filter['anchor'] == 'center' (else 'bottom')                      */
    double perspective_coeff, presence_min, distance_min; /* post_filter                                    */
    int32_t keep_alarm;           /* alarm_def                                                              */
    double duration_min;
} gofpid_tracker_config;
/* One tracked blob (an entry of the reference's tracked_blobs_, pipeline.py:630-645). */
#define GOFPID_FILTER_PRESENCE 1
#define GOFPID_FILTER_DISTANCE 2
#define GOFPID_FILTER_PERIMETER 4
#define GOFPID_FILTER_PERSPECTIVE 8
typedef struct gofpid_track {
    int16_t anchor[2], bottom[2], center[2], center_first[2];
    int32_t presence, absence;
    int32_t filtered;             /* GOFPID_FILTER_* bits: the reference's set of filter names            */
    int32_t already_in_alarm;
    int32_t x0, y0, x1, y1;       /* cv.boundingRect of the contour (inclusive corners)                     */
    int32_t n_points;             /* contour length (gofpid_tracker_contour)                                */
} gofpid_track;
GOFPID_API int gofpid_tracker_create(const gofpid_tracker_config *cfg, gofpid_tracker **out);
GOFPID_API int gofpid_tracker_destroy(gofpid_tracker *t);
GOFPID_API int gofpid_tracker_reset(gofpid_tracker *t);
/* Steps 5-7 for one frame from the blob records of gofpid_pipeline_blobs / a clip call (centre and bottom
 * are formed from the Green sums as GOFPID._create_blob does); *alarm = what GOFPID.detect returns. */
GOFPID_API int gofpid_tracker_step(gofpid_tracker *t, const gofpid_blob *h_blobs, int n_blobs,
                                   const int32_t *h_points, int *alarm);
/* The same for blobs given as the reference's blob dictionaries hold them: anchor / bottom / center as
 * n_blobs (x, y) int16 pairs, h_boxes = (x0, y0, x1, y1) int32 per blob, contour i = points
 * h_offsets[i] .. h_offsets[i + 1] of h_points. */
GOFPID_API int gofpid_tracker_step_prepared(gofpid_tracker *t, int n_blobs, const int16_t *h_anchor,
                                            const int16_t *h_bottom, const int16_t *h_center,
                                            const int32_t *h_boxes, const int32_t *h_points,
                                            const int32_t *h_offsets, int *alarm);
/* A whole clip: the outputs of gofpid_pipeline_process_host_sequence / _device_sequence (h_counts: blob and
 * point count per frame and stream) for camera stream `stream_idx`; h_alarms receives one byte per frame. */
GOFPID_API int gofpid_tracker_run(gofpid_tracker *t, const gofpid_blob *h_blobs, const int32_t *h_points,
                                  const int32_t *h_counts, int n_frames, int n_streams, int stream_idx,
                                  uint8_t *h_alarms);
GOFPID_API int gofpid_tracker_count(gofpid_tracker *t);
GOFPID_API int gofpid_tracker_get(gofpid_tracker *t, int i, gofpid_track *out);
GOFPID_API int gofpid_tracker_contour(gofpid_tracker *t, int i, int32_t *h_points, int point_capacity);
GOFPID_API int gofpid_tracker_alarm_state(gofpid_tracker *t, int *already_in_alarm, double *duration);

/* ---- stage-level entry points (device buffers, used by the parity tests) ------------- */
/* cv.GaussianBlur / cv.blur on u8, BORDER_REFLECT_101 (pipeline.py:463). */
GOFPID_API int gofpid_blur_u8(const uint8_t *d_src, size_t src_pitch, uint8_t *d_dst,
                              size_t dst_pitch, int H, int W, int C, int kind, int kw, int kh,
                              double sigma, void *stream);
/* Chain of cv.erode / cv.dilate, fused in shared memory (pipeline.py:473-475). */
GOFPID_API int gofpid_morph_u8(const uint8_t *d_src, size_t src_pitch, uint8_t *d_dst,
                               size_t dst_pitch, int H, int W, const gofpid_morph_op *ops,
                               int n_ops, void *stream);
/* Same chain on the bit-plane kernel the pipeline uses: the mask may only hold 0 and the given
 * `levels` (1..3 ascending values, e.g. {127, 255}); other values are an input contract
 * violation (they are read as the largest level not above them). */
GOFPID_API int gofpid_morph_levels_u8(const uint8_t *d_src, size_t src_pitch, uint8_t *d_dst,
                                      size_t dst_pitch, int H, int W, const gofpid_morph_op *ops,
                                      int n_ops, const int *levels, int n_levels, void *stream);
/* 8-connected labelling of a u8 mask (non-zero = foreground) into canonical int32 labels. */
GOFPID_API int gofpid_ccl_u8(const uint8_t *d_mask, size_t mask_pitch, int32_t *d_labels,
                             size_t labels_pitch, int H, int W, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GOFPID_B200_H */
