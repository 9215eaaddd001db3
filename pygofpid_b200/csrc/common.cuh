// common.cuh -- shared declarations of libgofpid_b200 (sm_100a only).
//
// Internal interface between the translation units:
//   detect.cu   fused blur + {frame differencing, MOG2, KNN} kernels, frame tick, cv::RNG fill
//   morph.cu    fused erode/dilate chain
//   blobs.cu    union-find labelling, component statistics, outer-border tracing
//   pipeline.cu the C ABI of include/gofpid_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gofpid_b200.h"

namespace gofpid {

// ------------------------------------------------------------------ error plumbing
void set_error(const char *fmt, ...);
#define GOFPID_CUDA_TRY(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            gofpid::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                              __FILE__, __LINE__);                                         \
            return GOFPID_ERR_CUDA;                                                        \
        }                                                                                  \
    } while (0)
#define GOFPID_TRY(expr)                                                                   \
    do {                                                                                   \
        int _s = (expr);                                                                   \
        if (_s != GOFPID_OK) return _s;                                                    \
    } while (0)

static inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// Experiment / test knobs (GOFPID_* environment variables), read ONCE per handle in
// gofpid_pipeline_create -- never on the per-frame path.
struct Tuning {
    int depth = 0;        // GOFPID_DEPTH: frame slots (0: default)
    int group = 0;        // GOFPID_MOG2_GROUP: frames per MOG2 launch in the clip entry points (0: default)
    int b_prio_low = 0;   // GOFPID_B_PRIO=0: stage-B streams at the lowest priority
    int no_graph = 0;     // GOFPID_NO_GRAPH=1: eager launches only
    int no_cond = 0;      // GOFPID_NO_COND=1: tiled fallback launched unconditionally (ncu cannot profile conditional nodes)
    int morph_bytes = 0;  // GOFPID_MORPH_BYTES=1: byte morphology kernel inside the pipeline
    int knn_scalar = 0;   // GOFPID_KNN_SCALAR=1: general KNN kernel on gray frames too (tests)
    int runs_cap = -1;    // GOFPID_RUNS_CAP: run capacity of the shared-memory labeller (tests: force the fallback)
    int copier_threads = -1; // GOFPID_COPIER_THREADS: worker threads staging pageable host frames (-1: default, by host threads per GPU)
    int copier_stream = 1;   // GOFPID_COPIER_STREAM=0: plain memcpy instead of streaming stores into the staging buffer
    int copier_spin_us = 500; // GOFPID_COPIER_SPIN_US: how long a worker polls for the next frame before it sleeps
    int h2d_chunk = 0;    // GOFPID_H2D_CHUNK: frames per host->device copy of a host clip (0: the frames of one detector launch)
    int mog2_smem_pad = 0; // GOFPID_MOG2_SMEM_PAD: extra dynamic shared memory per MOG2 CTA (caps the CTAs per SM: A/B measurements)
    int mog2_generic = 0; // GOFPID_MOG2_GENERIC=1: the generic MOG2 kernel even for ordinary parameters (tests)
    int event_timing = 0; // GOFPID_EVENT_TIMING=1: slot events keep timestamps (debug build tools)
};
Tuning read_tuning();

// ------------------------------------------------------------------ geometry
// Fused blur+detect kernels: a CTA owns a tile of (32*PX) x DET_TH pixels; thread (tx, ty)
// owns PX consecutive pixels of one row.
constexpr int DET_TH = 8;
constexpr int DET_PX = 4;

// BORDER_REFLECT_101, iterated for tiny images (cv2 BORDER_DEFAULT).
__host__ __device__ inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// ------------------------------------------------------------------ blur
struct BlurParams {
    int kind;       // GOFPID_BLUR_*
    int kw, kh;     // kernel size (1 when kind == NONE)
    int rx, ry;     // anchors kw/2, kh/2
    int box_d, box_delta, box_scale; // box normalisation (see detect.cu)
    uint16_t qx[GOFPID_MAX_KSIZE + 1];
    uint16_t qy[GOFPID_MAX_KSIZE + 1];
    // colour -> gray fused into the frame load (cv.cvtColor COLOR_BGR2GRAY / COLOR_RGB2GRAY of
    // pipeline.py:458-459): the source has 3 interleaved channels, the kernels work on 1;
    // gray = (c0*cw[0] + c1*cw[1] + c2*cw[2] + 2^14) >> 15, OpenCV's 15-bit fixed point
    int conv;        // 0: frames already have the model's channel count
    uint32_t cw[3];
    // 3x3 kernels whose taps fit a byte: the taps packed for dp4a (0: not available)
    uint32_t qx3, qy3;
    uint32_t q3x[3], q3y[3]; // ... and one per word (operands straight from the constant bank)
};
int make_blur_params(int kind, int kw, int kh, double sigma, BlurParams *bp);
int gaussian_taps_q8(int n, double sigma, uint16_t *q);

struct FrameSrc {
    const uint8_t *base;
    size_t pitch;         // bytes between rows
    size_t stream_stride; // bytes between streams
};
struct ImageU8 {
    uint8_t *base;
    size_t pitch;
    size_t stream_stride;
};

// ------------------------------------------------------------------ per-frame scalars
// Written on the device by tick_kernel at the start of every step so that a captured
// CUDA graph can be replayed unchanged.
struct FrameTick {
    int32_t nframes;  // 1-based index of the frame being processed
    // MOG2: alphaT, alpha1, prune of the next MOG2_MAX_GROUP frames (a launch consumes the first
    // `nf` of them, its last tile prepares the following ones)
    alignas(16) float mog2_lr[8][4];
    // KNN
    int32_t knn_u[3];       // update periods uS, uM, uL
    int32_t knn_counter[3]; // counters cS, cM, cL valid DURING the frame
    int32_t knn_refresh[3]; // tier draws fresh next-update ticks this frame
    uint64_t knn_rng_start;    // RNG state before this frame's fills
    uint64_t knn_draw_off[3];  // draws before each refreshed tier's fill, within one stream
    uint64_t knn_draws;        // draws one stream consumes this frame (all refreshed tiers)
    uint64_t knn_rng;          // RNG state after this frame's fills (all streams)
    int32_t mog2_done;      // tiles of the running mog2_kernel launch that have finished
    // FD
    int32_t fd_valid; // a previous frame exists
    // where this step's frames are (set by the tick kernel, the only node of the captured
    // step graph whose parameters change from frame to frame)
    FrameSrc src;
};

// ------------------------------------------------------------------ detector state
struct Mog2Params {
    float Tb, TB, Tg, varInit, varMin, varMax, tau;
    int detectShadows, shadowVal, history;
    float ct;
    int force_generic; // Tuning::mog2_generic: never take the lean kernel
    int smem_pad;      // Tuning::mog2_smem_pad
};
// MOG2 model, tiled: per stream tiles_x * tiles_y tiles of MOG2_TILE_W x MOG2_TILE_H pixels; a
// tile is one contiguous block [mode][plane: weight, var, mean c..][512 px] of floats, so its
// live modes are a single run of bytes for the bulk-copy engine.
#ifndef GOFPID_MOG2_TILE_W
#define GOFPID_MOG2_TILE_W 64
#endif
constexpr int MOG2_TILE_W = GOFPID_MOG2_TILE_W;
constexpr int MOG2_TILE_H = 8;
// Temporal blocking: one launch carries a tile through up to MOG2_MAX_GROUP consecutive frames of
// its stream while the tile's modes stay in shared memory (model read and written once per group).
constexpr int MOG2_MAX_GROUP = 8;
struct Mog2Group {
    const uint8_t *frame[MOG2_MAX_GROUP]; // frame f of the group (stream 0; FrameSrc::pitch / stream_stride apply)
    uint8_t *mask[MOG2_MAX_GROUP];        // detector mask of frame f (stream 0; ImageU8::pitch / stream_stride apply)
    int nf;                               // frames in this launch (1 .. MOG2_MAX_GROUP)
    int aligned4;                         // pitch, stream stride and every frame pointer are multiples of 4 (set by the launcher)
};
int mog2_group_aligned4(const Mog2Group &g, const FrameSrc &src);
struct Mog2Model {
    float *planes;      // [stream][tile][5][2 + C][512]
    uint8_t *nmodes;    // [stream][tile][512]
    uint8_t *tile_nmax; // [stream][tile]: most live modes of any pixel of the tile
    int tiles_x, tiles_y;
};
struct KnnParams {
    float Tb, tau;
    int knn, detectShadows, shadowVal, history;
};
struct KnnModel {
    uint8_t *samples;      // per stream: 21 planes of PW*H*(C+1) bytes; gray: four pixels = (v0 v1 v2 v3 | f0 f1 f2 f3), colour: (c0 c1 c2 flag) per pixel
    uint32_t *nxt;         // per stream: PW*H words (nxtS | nxtM<<8 | nxtL<<16)
    uint16_t *idx;         // per stream: PW*H halfwords (idxS | idxM<<3 | idxL<<6)
    uint8_t *fresh;        // per stream: 3 dense arrays of H*W bytes (RNG fills S, M, L)
    size_t plane_elems;    // PW*H
    int PW;
};
struct FdModel {
    uint8_t *prev; // per stream: previous blurred frame, pitch PB bytes
    size_t pitch, stream_stride;
};

// ------------------------------------------------------------------ launchers
struct DetectLaunch {
    int H, W, C, n_streams;
    FrameSrc src;
    BlurParams blur;
    ImageU8 mask; // (H, W) u8 per stream, padded pitch (multiple of 128 bytes)
    const FrameTick *tick;
    cudaStream_t stream;
};
struct TickArgs { // by-value kernel arguments of tick_kernel, kept by the pipeline for graph updates
    FrameTick *tick;
    int detector;
    Mog2Params mp;
    KnnParams kp;
    const int32_t *knn_u_table;
    int knn_table_len;
    const uint64_t *pow2tab;
    int total, n_streams;
    FrameSrc src;
};
int launch_tick(const TickArgs &a, cudaStream_t s);
const void *tick_kernel_func();
void tick_kernel_params(TickArgs &a, void **params); // params[10]
int launch_blur_only(const DetectLaunch &L, ImageU8 dst);
int launch_fd(const DetectLaunch &L, FdModel model, int thr);
// MOG2 derives its per-frame scalars itself (no tick kernel): by-value kernel arguments kept by the
// pipeline so that a captured step can be replayed with a new frame pointer
struct Mog2Args {
    BlurParams bp;
    ImageU8 mask;   // pitch / stream stride of the masks (bases: grp.mask)
    Mog2Model M;
    Mog2Params P;
    FrameTick *tick;
    FrameSrc src;   // pitch / stream stride of the frames (bases: grp.frame)
    Mog2Group grp;
    int H, W;
};
// one frame (L.src / L.mask) when grp == nullptr, else the frames and masks of the group
int launch_mog2(const DetectLaunch &L, Mog2Model model, const Mog2Params &mp, FrameTick *tick,
                const Mog2Group *grp = nullptr);
void mog2_prepare_lr(const Mog2Params &mp, long n_done, float (*lr)[4]); // rates of frames n_done+1 ..
const void *mog2_kernel_func(int C, int H, int W, const Mog2Params &mp);
void mog2_kernel_params(Mog2Args &a, void **params); // params[9]
int launch_knn(const DetectLaunch &L, KnnModel model, const KnnParams &kp,
               const uint64_t *d_jump_table, bool force_scalar = false);
int detect_launches_per_step(int detector);

// morphology ------------------------------------------------------------------
struct MorphRun { int8_t dy, dx; uint8_t len; uint8_t pad; }; // horizontal run of an SE row
struct MorphOpDev {
    int op;          // GOFPID_MORPH_*
    int kh, kw, ay, ax;
    int is_rect;     // every element set -> separable path
    int run_begin, run_end; // slice of the run table
};
struct MorphPlan {
    int n_ops;
    MorphOpDev ops[GOFPID_MAX_MORPH_OPS];
    int halo_l, halo_r, halo_t, halo_b; // accumulated over the chain
    MorphRun *d_runs;                   // device copy of the run table
    int n_runs;
    int anchored;                       // every SE contains its anchor (uniform tiles stay uniform)
    int tile_w, tile_h;                 // output tile per CTA
    size_t smem_bytes;
    // bit-plane path (morph_bits_kernel): the mask only takes the values {0} + levels[0..n_levels)
    // (ascending), so the chain runs on n_levels one-bit threshold planes, 32 pixels per word
    int n_levels;                       // 0: byte path
    int levels[3];
    int bit_sel[3];                     // bit of the byte that IS threshold plane i (-1: compare bytes)
    int bit_tile_h, bit_hw_l, bit_hw_r; // tile rows, halo words left / right (tile width = 256 px)
    int bit_pad;                        // spare rows above / below the staged rows in shared memory
    size_t bit_smem_bytes;
};
// Declares the set of non-zero values the chain's input can hold (at most 3); selects the bit-plane
// kernel.  Called by the pipeline, which knows its detector's output alphabet.
void morph_plan_set_levels(MorphPlan *plan, const int *levels, int n_levels);
int morph_plan_create(const gofpid_morph_op *ops, int n_ops, MorphPlan *plan);
void morph_plan_destroy(MorphPlan *plan);
// bitimg (optional, bit-plane path only): also writes the foreground bit image the blob stage reads
int launch_morph(const MorphPlan &plan, ImageU8 src, ImageU8 dst, int H, int W, int n_streams,
                 cudaStream_t s, uint32_t *bitimg = nullptr, int bit_pitch = 0, size_t bit_stream = 0);

// blobs ------------------------------------------------------------------------
// per-stream list of horizontal foreground runs (runs_extract_kernel -> runs_ccl_kernel)
struct RunList {
    uint16_t *xs, *xe, *y; // [stream][max_runs]
    int *rowstart, *rowend; // [stream][row_pitch]: slice of the list holding each image row's runs
    int *count;             // [stream]: runs of the current frame (reset by runs_ccl_kernel)
    int *n_labelled;        // [stream]: runs of the frame runs_ccl_kernel labelled last
    int32_t *label;         // [stream][max_runs]: label value (1 + first pixel index) of every run
    uint2 *byc;             // [stream][max_runs]: the runs grouped by component (y << 16 | xs, xe)
    int max_runs;           // capacity per stream (multiple of 8; 0: frame shape not supported)
    int row_pitch;
};
int run_list_alloc(RunList *R, int H, int bit_pitch, int n_streams);
void run_list_free(RunList *R);

struct BlobBuffers {
    RunList runs;
    int32_t *labels;      // per stream: LP*H int32 (signed: >0 fg root value, <=0 bg class)
    size_t labels_pitch;  // elements
    size_t labels_stream; // elements between streams
    int32_t *slot;        // per stream: LP*H int32, component slot of root pixels (sparse)
    uint32_t *bitimg;     // per stream: (H+2) rows of bit_pitch words; pixel (x,y) = bit x+32 of row y+1
    int bit_pitch;        // words per row (>= W/32 + 2 guard words)
    size_t bit_stream;    // words between streams
    int32_t *comp;        // per stream: max_blobs * 8 int32 component records
    gofpid_frame_info *info; // per stream
    int32_t *cand;        // per stream: max_blobs candidate slots
    gofpid_blob *kept;    // per stream: max_kept records
    int32_t *points;      // per stream: max_points * 2 int32
    int max_blobs, max_kept, max_points;
    double area_min;
    int runs_cap;         // < 0: the run list's own capacity (Tuning::runs_cap)
};
// bits_ready: the morphology kernel already wrote B.bitimg for this mask
int launch_blobs(const BlobBuffers &B, ImageU8 mask, int H, int W, int n_streams,
                 const FrameTick *tick, cudaStream_t s, int *n_launches, bool bits_ready,
                 const cudaGraphConditionalHandle *cond = nullptr, int nest_ctas = 1);
// h_*: device-side addresses of the pinned host mirrors (cudaHostGetDevicePointer)
int launch_results_to_host(const BlobBuffers &B, gofpid_frame_info *h_info, gofpid_blob *h_kept, int32_t *h_points,
                           int n_streams, cudaStream_t s);
int launch_fill_labels(const RunList &R, int z, const gofpid_frame_info *info, int32_t *labels, size_t LP,
                       cudaStream_t s);
int launch_canonical_labels(const uint8_t *mask, size_t mask_pitch, const int32_t *labels, size_t LP,
                            int32_t *out, size_t out_pitch_elems, int H, int W, cudaStream_t s);
int launch_ccl_only(const uint8_t *d_mask, size_t mask_pitch, int32_t *d_labels,
                    size_t labels_pitch_elems, int H, int W, cudaStream_t s);

} // namespace gofpid
