// pipeline.cu -- the C ABI of include/gofpid_b200.h.
//
// A gofpid_pipeline owns, in HBM, everything cv2 keeps inside GOFPID's per-stream objects
// (pipeline.py:231-240: the MOG2 / KNN subtractor or FrameDifferencing._model) for n_streams
// independent camera streams of one shape, plus the intermediate mask / label / blob buffers
// of the step.  One step = GOFPID.detect lines pipeline.py:461-479 for every stream.
#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace gofpid {

static thread_local char g_error[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

Tuning read_tuning()
{
    Tuning t;
    auto num = [](const char *name, int lo, int hi, int dflt) {
        const char *e = getenv(name);
        if (!e || !e[0]) return dflt;
        const int v = atoi(e);
        return v >= lo && v <= hi ? v : dflt;
    };
    auto flag = [](const char *name, char on) { const char *e = getenv(name); return e && e[0] == on ? 1 : 0; };
    t.depth = num("GOFPID_DEPTH", 1, 16, 0);
    t.group = num("GOFPID_MOG2_GROUP", 1, MOG2_MAX_GROUP, 0);
    t.b_prio_low = flag("GOFPID_B_PRIO", '0');
    t.no_graph = flag("GOFPID_NO_GRAPH", '1');
    t.no_cond = flag("GOFPID_NO_COND", '1');
    t.morph_bytes = flag("GOFPID_MORPH_BYTES", '1');
    t.knn_scalar = flag("GOFPID_KNN_SCALAR", '1');
    t.runs_cap = num("GOFPID_RUNS_CAP", 0, 1 << 30, -1);
    t.event_timing = flag("GOFPID_EVENT_TIMING", '1');
    t.mog2_generic = flag("GOFPID_MOG2_GENERIC", '1');
    t.mog2_smem_pad = num("GOFPID_MOG2_SMEM_PAD", 0, 65536, 0);
    t.h2d_chunk = num("GOFPID_H2D_CHUNK", 0, 16, 0);
    t.copier_threads = num("GOFPID_COPIER_THREADS", 0, 64, -1);
    t.copier_spin_us = num("GOFPID_COPIER_SPIN_US", 0, 1000000, 500);
    t.copier_stream = num("GOFPID_COPIER_STREAM", 0, 1, 1);
    return t;
}

} // namespace gofpid

using namespace gofpid;

// NVTX ranges around the host side of a step (visible in nsys / `ncu --nvtx`; a few nanoseconds each when no
// tool is attached): uploads, the enqueue of stage A + stage B, result collection, re-runs of overflowed frames.
namespace {
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
} // namespace

// Up to DEPTH frames are in flight: stage A (tick, blur+detect) of frame t+1 runs on the caller's
// stream while stage B (morphology, run extraction, labelling, border tracing, results) of the
// frames before it runs on the internal streams of their slots.  The temporal model makes stage A strictly
// sequential; stage B only needs the finished mask of its own frame and occupies a handful of
// SMs (one CTA per stream or per blob), so consecutive frames of ONE camera stream overlap and
// the frame rate is set by stage A alone.
constexpr int MAX_DEPTH = 16;
constexpr int DEFAULT_DEPTH = 12;      // few camera streams per handle: deep, stage B latency is several stage As
                                       // (measured, 6 -> 12 slots: 1080p KNN 13.1 k -> 13.9 k frames/s, 320x240 FD end to end 80 k -> 88 k)
constexpr int DEFAULT_DEPTH_MANY = 4;  // many streams per handle: every kernel is already wide, slots cost memory
// MOG2 handles: the clip entry points carry a tile through DEFAULT_GROUP consecutive frames per
// launch (temporal blocking, see mog2_kernel), which wants two groups' worth of slots and a few more
constexpr int DEFAULT_GROUP = 3; // (measured: 2 / 3 / 4 / 8 frames per launch = 30.7 / 28.2 / 33.8 / 32.0 us per 1080p frame with blobs)
constexpr int DEFAULT_DEPTH_MOG2 = 12, DEFAULT_DEPTH_MOG2_MANY = 8;
// (GOFPID_MOG2_GROUP overrides: 1 .. MOG2_MAX_GROUP, at most half the slots)
// (GOFPID_DEPTH overrides: 1 .. MAX_DEPTH)

// Pageable host frames (what cv.VideoCapture.read() hands to GOFPID.detect) have to pass through page-locked
// memory before the copy engine can take them.  One thread moves 2 MB in about 200 us -- five times what the
// frame then needs over the link -- so the rows of a frame are split over a few persistent worker threads
// (fork / join per frame; the workers keep polling for a while after a job so that a per-frame caller finds
// them awake), a frame is uploaded strip by strip while it is still being copied, and every frame slot has
// its own page-locked buffer so that filling the next one overlaps the upload of the previous one.
class HostCopier {
public:
    static constexpr int MAX_STRIPS = 16;
    HostCopier(int workers, int spin_us, bool streaming) : spin_us_(spin_us), streaming_(streaming)
    {
        for (int i = 0; i < workers; i++) th_.emplace_back([this] { worker(); });
    }
    ~HostCopier()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            gen_pub_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    // nrows rows of `row` bytes, source rows spitch apart, destination rows dpitch apart.  The rows are cut
    // into chunks that the caller and whichever workers are awake take one after the other: a worker that
    // wakes up late simply finds nothing left, so the call is never slower than the caller copying alone.
    // The chunks form `nstrips` consecutive strips of rows; on_strip(first_row, n_rows) is called by the
    // CALLER's thread, strip by strip in order, as soon as every chunk of the strip has been copied -- the
    // caller starts the strip's upload while the rest of the frame is still being copied.
    template <class OnStrip>
    void copy2d(uint8_t *dst, size_t dpitch, const uint8_t *src, size_t spitch, size_t row, size_t nrows, int nstrips,
                OnStrip &&on_strip)
    {
        bool wake;
        {
            // a job is described under the mutex, and only once no worker is still inside the previous one
            // (a worker that wakes up late must find either nothing to do or a completely described job)
            std::unique_lock<std::mutex> lk(m_);
            while (active_ > 0) { lk.unlock(); std::this_thread::yield(); lk.lock(); }
            dst_ = dst; src_ = src; dpitch_ = dpitch; spitch_ = spitch; row_ = row; nrows_ = nrows;
            rows_per_chunk_ = std::max<size_t>(1, (size_t)32768 / std::max<size_t>(row, 1));
            nchunks_ = (int)((nrows + rows_per_chunk_ - 1) / rows_per_chunk_);
            nstrips = std::max(1, std::min(std::min(nstrips, (int)MAX_STRIPS), nchunks_));
            chunks_per_strip_ = (nchunks_ + nstrips - 1) / nstrips;
            nstrips_ = (nchunks_ + chunks_per_strip_ - 1) / chunks_per_strip_;
            for (int k = 0; k < nstrips_; k++) strip_done_[k].store(0, std::memory_order_relaxed);
            next_.store(0, std::memory_order_release);
            wake = !th_.empty() && nchunks_ > 4;
            if (wake) { gen_++; gen_pub_.fetch_add(1, std::memory_order_release); }
        }
        if (wake) cv_.notify_all();
        int issued = 0;
        while (issued < nstrips_) {
            const bool worked = take_chunk();
            while (issued < nstrips_) {
                const int c0 = issued * chunks_per_strip_, c1 = std::min(nchunks_, c0 + chunks_per_strip_);
                if (strip_done_[issued].load(std::memory_order_acquire) < c1 - c0) break;
                const size_t r0 = (size_t)c0 * rows_per_chunk_, r1 = std::min(nrows_, (size_t)c1 * rows_per_chunk_);
                on_strip(r0, r1 - r0);
                issued++;
            }
            if (!worked && issued < nstrips_) cpu_pause();
        }
    }
    void copy2d(uint8_t *dst, size_t dpitch, const uint8_t *src, size_t spitch, size_t row, size_t nrows)
    {
        copy2d(dst, dpitch, src, spitch, row, nrows, 1, [](size_t, size_t) {});
    }

private:
    static void cpu_pause()
    {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    // The staging buffer is written once and then read by the copy engine: streaming stores spare the
    // read-for-ownership of every destination line (a third of the memory traffic of a plain memcpy).
    void copy_bytes(uint8_t *d, const uint8_t *s, size_t n) const
    {
#if defined(__x86_64__)
        if (streaming_ && n >= 256) {
            size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
            memcpy(d, s, head);
            d += head; s += head; n -= head;
            size_t i = 0;
            for (; i + 64 <= n; i += 64) {
                const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i));
                const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 16));
                const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 32));
                const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 48));
                _mm_stream_si128(reinterpret_cast<__m128i *>(d + i), a);
                _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 16), b);
                _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 32), c);
                _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 48), e);
            }
            memcpy(d + i, s + i, n - i);
            _mm_sfence(); // (streaming stores are weakly ordered: before the chunk is reported done)
            return;
        }
#endif
        memcpy(d, s, n);
    }
    bool take_chunk()
    {
        if (next_.load(std::memory_order_acquire) >= nchunks_) return false; // (a waiting caller polls this: no unbounded increments)
        const int c = next_.fetch_add(1, std::memory_order_acq_rel);
        if (c >= nchunks_) return false;
        const size_t r0 = (size_t)c * rows_per_chunk_, r1 = std::min(nrows_, r0 + rows_per_chunk_);
        if (dpitch_ == row_ && spitch_ == row_) copy_bytes(dst_ + r0 * row_, src_ + r0 * row_, (r1 - r0) * row_);
        else
            for (size_t r = r0; r < r1; r++) copy_bytes(dst_ + r * dpitch_, src_ + r * spitch_, row_);
        strip_done_[c / chunks_per_strip_].fetch_add(1, std::memory_order_acq_rel);
        return true;
    }
    void worker()
    {
        unsigned seen = 0, seen_pub = 0;
        for (;;) {
            // A synchronous per-frame caller comes back every few hundred microseconds; waking a sleeping
            // thread costs a good part of that.  So a worker keeps looking for the next job for spin_us_
            // after its last one before it goes to sleep on the condition variable.
            if (spin_us_ > 0) {
                const auto t0 = std::chrono::steady_clock::now();
                while (gen_pub_.load(std::memory_order_acquire) == seen_pub) {
                    cpu_pause();
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_)) break;
                }
            }
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                seen_pub = gen_pub_.load(std::memory_order_relaxed);
                active_++;
            }
            while (take_chunk()) {}
            {
                std::lock_guard<std::mutex> lk(m_);
                active_--;
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_;
    const int spin_us_;
    const bool streaming_;
    unsigned gen_ = 0;
    std::atomic<unsigned> gen_pub_{0}; // gen_ (and stop_) for the spinning workers, outside the mutex
    int active_ = 0; // workers inside a job
    bool stop_ = false;
    std::atomic<int> next_{0};
    std::atomic<int> strip_done_[MAX_STRIPS];
    int nchunks_ = 0, nstrips_ = 1, chunks_per_strip_ = 1;
    size_t rows_per_chunk_ = 1;
    uint8_t *dst_ = nullptr;
    const uint8_t *src_ = nullptr;
    size_t dpitch_ = 0, spitch_ = 0, row_ = 0, nrows_ = 0;
};

struct FrameSlot {
    uint8_t *h_stage = nullptr;   // page-locked staging of pageable host frames (allocated on first use)
    uint8_t *d_final = nullptr;   // final mask of the frame in this slot
    uint8_t *d_mask_det = nullptr; // detector output of the frame when a morphology chain follows
    BlobBuffers blobs = {};
    cudaEvent_t ev_mask = nullptr, ev_blobs = nullptr; // stage A / stage B of this slot finished
    gofpid_frame_info *h_info = nullptr;               // pinned host copies of the results
    gofpid_blob *h_kept = nullptr;
    int32_t *h_points = nullptr;
    gofpid_frame_info *dh_info = nullptr;              // the same buffers as the device addresses them
    gofpid_blob *dh_kept = nullptr;
    int32_t *dh_points = nullptr;
    uint8_t *d_frames = nullptr;  // staging of host frames
    cudaEvent_t ev_h2d = nullptr;
    bool used = false, fetched = false;
    int frame_index = 0;
    // stage graphs of this slot (buffers are baked into the kernel arguments)
    cudaGraph_t graph_a = nullptr, graph_b = nullptr, graph_b2 = nullptr;
    cudaGraphExec_t exec_a = nullptr, exec_b = nullptr;
    // stage B without the conditional node (the four fallback kernels launched, exiting early): used
    // for frames that come from the host -- next to the frame uploads a conditional node costs the
    // end-to-end rate 3 us per 1080p frame, while device-resident clips lose 3 us to the four launches
    cudaGraphExec_t exec_b2 = nullptr;
    cudaGraphNode_t tick_node = nullptr;   // the node of stage A that takes the frame pointer (tick or MOG2 kernel)
    cudaKernelNodeParams src_node_params = {}; // its launch geometry, as captured
    Mog2Args mog2_args;                    // by-value MOG2 kernel arguments of this slot
    cudaGraphConditionalHandle cond_b = 0; // raised by runs_ccl_kernel: run the tiled labelling fallback
};

struct gofpid_pipeline {
    gofpid_config cfg;
    Tuning tune; // GOFPID_* knobs as they were when the handle was created
    std::vector<std::vector<uint8_t>> se_copies;
    int H = 0, W = 0, C = 0, n_streams = 0, device = 0; // C: channels the kernels work on
    int Csrc = 0;                                        // channels of the frames (3 with a gray conversion)
    cudaStream_t stream = nullptr;   // own stage-A stream (host entry points)
    cudaStream_t stream_b[MAX_DEPTH] = {}; // stage B, one per frame slot
    int depth = DEFAULT_DEPTH;            // frame slots in use
    int group = 1;                        // frames per MOG2 launch in the clip entry points
    int launches_b = 0;                   // kernels of one captured stage B
    int launches_b2 = 0;                  // ... of its unconditional variant (frames from the host)
    bool from_host = false;               // the step being enqueued takes its frames from the host
    long long launches_total = 0;         // kernels launched since creation / reset
    cudaStream_t stream_c = nullptr; // result fetches
    cudaStream_t stream_h = nullptr; // host->device frame copies (a second copy stream was measured: no gain)
    BlurParams blur;
    Mog2Params mog2p;
    KnnParams knnp;
    int fd_thr = 0;

    FrameTick *d_tick = nullptr;
    size_t mask_pitch = 0, mask_stream = 0;

    Mog2Model mog2 = {};
    KnnModel knn = {};
    FdModel fd = {};
    int32_t *d_knn_table = nullptr;
    int knn_table_len = 0;
    uint64_t *d_pow2tab = nullptr;

    MorphPlan morph = {};
    FrameSlot slots[MAX_DEPTH];
    long frames = 0;  // frames submitted so far
    int selected = 0; // age of the frame the result getters refer to (0 newest, 1 previous)

    size_t frames_pitch = 0, frames_stream = 0;
    const uint8_t *pinned_lo = nullptr, *pinned_hi = nullptr; // host range last found to be page-locked
    HostCopier *copier = nullptr; // worker threads that move pageable frames into page-locked staging
    uint8_t *d_frames_base = nullptr; // device staging of all slots (slot i at i * frames_stream * n_streams)

    // a frame that exceeded the slot capacities, re-run with buffers of the size it needs (rerun_big)
    struct {
        const FrameSlot *slot = nullptr;
        int frame_index = -1, stream = -1;
        gofpid_frame_info info = {};
        std::vector<gofpid_blob> kept;
        std::vector<int32_t> points;
    } big;
    // blob lists of the last clip call when the caller's buffers were too small (gofpid_pipeline_sequence_results)
    std::vector<gofpid_blob> seq_blobs;
    std::vector<int32_t> seq_points;

    bool profiling = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int launch_count = 0;
    bool use_graph = true;
    TickArgs tick_args;
    int graph_launches = 0;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define GOFPID_GUARD(p)                                                             \
    if (!(p)) { set_error("null pipeline handle"); return GOFPID_ERR_INVALID; }    \
    DeviceGuard _guard((p)->device);                                                \
    if (!_guard.ok) { set_error("cannot select CUDA device %d", (p)->device); return GOFPID_ERR_CUDA; }

int check_stream_index(const gofpid_pipeline *p, int idx)
{
    if (idx < 0 || idx >= p->n_streams) {
        set_error("stream index %d out of range [0, %d)", idx, p->n_streams);
        return GOFPID_ERR_INVALID;
    }
    return GOFPID_OK;
}

uint64_t host_mulmod(uint64_t a, uint64_t b)
{
    const uint64_t m = (4164903690ULL << 32) - 1ULL;
    return (uint64_t)(((unsigned __int128)a * b) % m);
}

// KNN update periods for frame n (cv::BackgroundSubtractorKNNImpl::apply)
void knn_periods(long n, int history, int nsamples, int32_t *u)
{
    const long n2 = 2 * n;
    const double lr = 1.0 / (double)(n2 < history ? n2 : history);
    const int Ks = (int)(log(0.7) / log(1 - lr)) + 1;
    const int Km = (int)(log(0.4) / log(1 - lr)) - Ks + 1;
    const int Kl = (int)(log(0.1) / log(1 - lr)) - Ks - Km + 1;
    u[0] = Ks / nsamples + 1;
    u[1] = Km / nsamples + 1;
    u[2] = Kl / nsamples + 1;
}

int alloc_zero(void **ptr, size_t bytes)
{
    GOFPID_CUDA_TRY(cudaMalloc(ptr, bytes ? bytes : 1));
    GOFPID_CUDA_TRY(cudaMemset(*ptr, 0, bytes ? bytes : 1));
    return GOFPID_OK;
}

// the slot of the frame the getters refer to (nullptr: no such frame yet)
FrameSlot *selected_slot(gofpid_pipeline *p)
{
    const long idx = p->frames - 1 - p->selected;
    if (idx < 0) return nullptr;
    FrameSlot *S = &p->slots[idx % p->depth];
    return S->used ? S : nullptr;
}

int reset_state(gofpid_pipeline *p)
{
    FrameTick t;
    memset(&t, 0, sizeof(t));
    t.knn_rng = p->cfg.knn_rng_seed ? p->cfg.knn_rng_seed : 0xffffffffULL;
    if (p->cfg.detector == GOFPID_DETECT_MOG2) // learning rates of the first frames (mog2_kernel prepares the following ones)
        mog2_prepare_lr(p->mog2p, 0, t.mog2_lr);
    GOFPID_CUDA_TRY(cudaMemcpy(p->d_tick, &t, sizeof(t), cudaMemcpyHostToDevice));
    if (p->cfg.detector == GOFPID_DETECT_MOG2) {
        const size_t nt = (size_t)p->mog2.tiles_x * p->mog2.tiles_y * p->n_streams;
        GOFPID_CUDA_TRY(cudaMemset(p->mog2.nmodes, 0, nt * MOG2_TILE_W * MOG2_TILE_H));
        GOFPID_CUDA_TRY(cudaMemset(p->mog2.tile_nmax, 0, nt));
    } else if (p->cfg.detector == GOFPID_DETECT_KNN) {
        const size_t pe = p->knn.plane_elems;
        GOFPID_CUDA_TRY(cudaMemset(p->knn.samples, 0, pe * 21 * (p->C + 1) * p->n_streams));
        GOFPID_CUDA_TRY(cudaMemset(p->knn.nxt, 0, pe * sizeof(uint32_t) * p->n_streams));
        GOFPID_CUDA_TRY(cudaMemset(p->knn.idx, 0, pe * sizeof(uint16_t) * p->n_streams));
    }
    for (auto &S : p->slots) { S.used = false; S.fetched = false; S.frame_index = 0; }
    p->frames = 0;
    p->selected = 0;
    p->launches_total = 0;
    return GOFPID_OK;
}

void free_all(gofpid_pipeline *p)
{
    cudaFree(p->d_tick);
    cudaFree(p->mog2.planes); cudaFree(p->mog2.nmodes); cudaFree(p->mog2.tile_nmax);
    cudaFree(p->knn.samples); cudaFree(p->knn.nxt); cudaFree(p->knn.idx); cudaFree(p->knn.fresh);
    cudaFree(p->fd.prev);
    cudaFree(p->d_knn_table); cudaFree(p->d_pow2tab);
    morph_plan_destroy(&p->morph);
    for (auto &S : p->slots) {
        cudaFree(S.d_final);
        cudaFree(S.d_mask_det);
        run_list_free(&S.blobs.runs);
        cudaFree(S.blobs.labels); cudaFree(S.blobs.slot); cudaFree(S.blobs.comp); cudaFree(S.blobs.bitimg);
        cudaFree(S.blobs.info); cudaFree(S.blobs.kept); cudaFree(S.blobs.points);
        if (S.h_info) cudaFreeHost(S.h_info);
        if (S.h_kept) cudaFreeHost(S.h_kept);
        if (S.h_points) cudaFreeHost(S.h_points);
        if (S.ev_mask) cudaEventDestroy(S.ev_mask);
        if (S.ev_blobs) cudaEventDestroy(S.ev_blobs);
        if (S.ev_h2d) cudaEventDestroy(S.ev_h2d);
        if (S.exec_a) cudaGraphExecDestroy(S.exec_a);
        if (S.exec_b) cudaGraphExecDestroy(S.exec_b);
        if (S.exec_b2) cudaGraphExecDestroy(S.exec_b2);
        if (S.graph_b2) cudaGraphDestroy(S.graph_b2);
        if (S.graph_a) cudaGraphDestroy(S.graph_a);
        if (S.graph_b) cudaGraphDestroy(S.graph_b);
    }
    cudaFree(p->d_frames_base);
    delete p->copier;
    p->copier = nullptr;
    for (auto &S : p->slots)
        if (S.h_stage) cudaFreeHost(S.h_stage);
    for (auto &e : p->ev) if (e) cudaEventDestroy(e);
    for (cudaStream_t s : {p->stream, p->stream_c, p->stream_h})
        if (s) cudaStreamDestroy(s);
    for (cudaStream_t s : p->stream_b)
        if (s) cudaStreamDestroy(s);
}

int build(gofpid_pipeline *p)
{
    const gofpid_config &cfg = p->cfg;
    const int H = p->H, W = p->W, C = p->C, S = p->n_streams;
    for (cudaStream_t *s : {&p->stream, &p->stream_c, &p->stream_h})
        GOFPID_CUDA_TRY(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking));
    {
        const bool mog2 = cfg.detector == GOFPID_DETECT_MOG2;
        if (mog2) p->depth = p->n_streams <= 8 ? DEFAULT_DEPTH_MOG2 : DEFAULT_DEPTH_MOG2_MANY;
        else p->depth = p->n_streams <= 8 ? DEFAULT_DEPTH : DEFAULT_DEPTH_MANY;
        if (cfg.depth >= 1 && cfg.depth <= MAX_DEPTH) p->depth = cfg.depth;
        if (p->tune.depth >= 1 && p->tune.depth <= MAX_DEPTH) p->depth = p->tune.depth;
        p->group = mog2 ? DEFAULT_GROUP : 1;
        if (mog2 && cfg.mog2_group >= 1 && cfg.mog2_group <= MOG2_MAX_GROUP) p->group = cfg.mog2_group;
        if (mog2 && p->tune.group >= 1) p->group = p->tune.group;
        p->group = std::max(1, std::min(p->group, p->depth / 2));
        // stage B runs at a higher priority than stage A: its few CTAs take the SMs they need as soon
        // as the detector's CTAs drain instead of queueing behind the whole grid
        int lo = 0, hi = 0;
        GOFPID_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const int prio = p->tune.b_prio_low ? lo : hi;
        for (int i = 0; i < p->depth; i++)
            GOFPID_CUDA_TRY(cudaStreamCreateWithPriority(&p->stream_b[i], cudaStreamNonBlocking, prio));
    }
    for (auto &e : p->ev) GOFPID_CUDA_TRY(cudaEventCreate(&e));
    GOFPID_TRY(make_blur_params(cfg.blur_kind, cfg.blur_kw, cfg.blur_kh, cfg.blur_sigma, &p->blur));
    if (cfg.convert != GOFPID_CONVERT_NONE) { // cv.cvtColor u8: B*3735 + G*19235 + R*9798, 15-bit fixed point
        const bool bgr = cfg.convert == GOFPID_CONVERT_BGR2GRAY;
        p->blur.conv = 1;
        p->blur.cw[0] = bgr ? 3735 : 9798; p->blur.cw[1] = 19235; p->blur.cw[2] = bgr ? 9798 : 3735;
    }

    GOFPID_TRY(alloc_zero((void **)&p->d_tick, sizeof(FrameTick)));
    p->mask_pitch = round_up((size_t)W, 128);
    p->mask_stream = p->mask_pitch * H;

    const int PW = (int)round_up((size_t)W, 32);
    const size_t pe = (size_t)PW * H;
    p->mog2.tiles_x = (W + MOG2_TILE_W - 1) / MOG2_TILE_W;
    p->mog2.tiles_y = (H + MOG2_TILE_H - 1) / MOG2_TILE_H;
    p->knn.PW = PW; p->knn.plane_elems = pe;
    Mog2Params &mp = p->mog2p;
    mp.Tb = cfg.mog2_var_threshold; mp.TB = cfg.mog2_background_ratio; mp.Tg = cfg.mog2_var_threshold_gen;
    mp.varInit = cfg.mog2_var_init; mp.varMin = cfg.mog2_var_min; mp.varMax = cfg.mog2_var_max;
    mp.tau = cfg.shadow_tau; mp.detectShadows = cfg.detect_shadows; mp.shadowVal = cfg.shadow_value;
    mp.history = cfg.mog2_history; mp.ct = cfg.mog2_ct;
    mp.force_generic = p->tune.mog2_generic;
    mp.smem_pad = p->tune.mog2_smem_pad;
    KnnParams &kp = p->knnp;
    kp.Tb = cfg.knn_dist2_threshold; kp.tau = cfg.shadow_tau; kp.knn = cfg.knn_k;
    kp.detectShadows = cfg.detect_shadows; kp.shadowVal = cfg.shadow_value; kp.history = cfg.knn_history;

    if (cfg.detector == GOFPID_DETECT_MOG2) {
        if (cfg.mog2_history < 1) { set_error("MOG2 history must be >= 1"); return GOFPID_ERR_INVALID; }
        const size_t nt = (size_t)p->mog2.tiles_x * p->mog2.tiles_y * S;
        const size_t tpx = (size_t)MOG2_TILE_W * MOG2_TILE_H;
        GOFPID_TRY(alloc_zero((void **)&p->mog2.planes, nt * GOFPID_MOG2_NMIX * (2 + C) * tpx * sizeof(float)));
        GOFPID_TRY(alloc_zero((void **)&p->mog2.nmodes, nt * tpx));
        GOFPID_TRY(alloc_zero((void **)&p->mog2.tile_nmax, nt));
    } else if (cfg.detector == GOFPID_DETECT_KNN) {
        if (cfg.knn_history < 1) { set_error("KNN history must be >= 1"); return GOFPID_ERR_INVALID; }
        GOFPID_TRY(alloc_zero((void **)&p->knn.samples, pe * 21 * (C + 1) * S));
        GOFPID_TRY(alloc_zero((void **)&p->knn.nxt, pe * sizeof(uint32_t) * S));
        GOFPID_TRY(alloc_zero((void **)&p->knn.idx, pe * sizeof(uint16_t) * S));
        GOFPID_TRY(alloc_zero((void **)&p->knn.fresh, (size_t)H * W * 3 * S));
        p->knn_table_len = (cfg.knn_history + 1) / 2 + 1;
        std::vector<int32_t> tab((size_t)p->knn_table_len * 3, 1);
        for (int n = 1; n < p->knn_table_len; n++)
            knn_periods(n, cfg.knn_history, GOFPID_KNN_NSAMPLES, &tab[(size_t)n * 3]);
        for (size_t i = 0; i < tab.size(); i++)
            if (tab[i] < 1 || tab[i] > 255) {
                set_error("KNN history %d gives an update period outside [1, 255]", cfg.knn_history);
                return GOFPID_ERR_UNSUPPORTED;
            }
        GOFPID_CUDA_TRY(cudaMalloc(&p->d_knn_table, tab.size() * sizeof(int32_t)));
        GOFPID_CUDA_TRY(cudaMemcpy(p->d_knn_table, tab.data(), tab.size() * sizeof(int32_t),
                                   cudaMemcpyHostToDevice));
        uint64_t pw[64];
        pw[0] = 4164903690ULL;
        for (int j = 1; j < 64; j++) pw[j] = host_mulmod(pw[j - 1], pw[j - 1]);
        GOFPID_CUDA_TRY(cudaMalloc(&p->d_pow2tab, sizeof(pw)));
        GOFPID_CUDA_TRY(cudaMemcpy(p->d_pow2tab, pw, sizeof(pw), cudaMemcpyHostToDevice));
    } else if (cfg.detector == GOFPID_DETECT_FD) {
        p->fd.pitch = round_up(round_up((size_t)W, 4) * C, 128);
        p->fd.stream_stride = p->fd.pitch * H;
        GOFPID_TRY(alloc_zero((void **)&p->fd.prev, p->fd.stream_stride * S));
        p->fd_thr = (int)floor(cfg.fd_threshold);
    } else {
        set_error("Unknown method for foreground detection (%d)", cfg.detector);
        return GOFPID_ERR_INVALID;
    }

    p->tick_args.tick = p->d_tick; p->tick_args.detector = cfg.detector;
    p->tick_args.mp = p->mog2p; p->tick_args.kp = p->knnp;
    p->tick_args.knn_u_table = p->d_knn_table; p->tick_args.knn_table_len = p->knn_table_len;
    p->tick_args.pow2tab = p->d_pow2tab; p->tick_args.total = H * W; p->tick_args.n_streams = S;
    p->tick_args.src = FrameSrc{nullptr, 0, 0};
    p->use_graph = !p->tune.no_graph;

    GOFPID_TRY(morph_plan_create(p->cfg.morph, p->cfg.n_morph, &p->morph));
    {
        // the detector's output alphabet selects the bit-plane morphology kernel
        int lv[3], n = 0;
        if (cfg.detector == GOFPID_DETECT_FD) {
            if (C == 3) { lv[n++] = 84; lv[n++] = 85; } // uint8-wrapping channel mean (segmentation.py:64)
            else lv[n++] = 255;
        } else {
            const int sv = cfg.shadow_value & 255;
            if (cfg.detect_shadows && sv > 0 && sv < 255) lv[n++] = sv;
            lv[n++] = 255;
        }
        if (!p->tune.morph_bytes) morph_plan_set_levels(&p->morph, lv, n);
    }

    if (cfg.find_blobs) {
        if (cfg.max_blobs < 1 || cfg.max_kept < 1 || cfg.max_points < 1) {
            set_error("blob capacities must be positive");
            return GOFPID_ERR_INVALID;
        }
        if (W > 8191 || H > 32767) { set_error("frames beyond 8191 x 32767 are not supported by the border tracer"); return GOFPID_ERR_UNSUPPORTED; }
        if (round_up((size_t)W, 4) * H >= (size_t)0x7fffffff) { set_error("frame too large for int32 labels"); return GOFPID_ERR_UNSUPPORTED; }
    }
    for (int si = 0; si < p->depth; si++) {
        FrameSlot &F = p->slots[si];
        GOFPID_TRY(alloc_zero((void **)&F.d_final, p->mask_stream * S));
        if (cfg.n_morph > 0) GOFPID_TRY(alloc_zero((void **)&F.d_mask_det, p->mask_stream * S));
        const unsigned evf = p->tune.event_timing ? cudaEventDefault : cudaEventDisableTiming; // (tools/timeline_debug.py)
        GOFPID_CUDA_TRY(cudaEventCreateWithFlags(&F.ev_mask, evf));
        GOFPID_CUDA_TRY(cudaEventCreateWithFlags(&F.ev_blobs, evf));
        GOFPID_CUDA_TRY(cudaEventCreateWithFlags(&F.ev_h2d, cudaEventDisableTiming));
        GOFPID_CUDA_TRY(cudaHostAlloc(&F.h_info, sizeof(gofpid_frame_info) * S, cudaHostAllocMapped));
        memset(F.h_info, 0, sizeof(gofpid_frame_info) * S);
        if (!cfg.find_blobs) continue;
        BlobBuffers &B = F.blobs;
        B.max_blobs = cfg.max_blobs; B.max_kept = cfg.max_kept; B.max_points = cfg.max_points;
        B.area_min = cfg.blob_area_min;
        B.runs_cap = p->tune.runs_cap;
        B.labels_pitch = round_up((size_t)W, 4);
        B.labels_stream = B.labels_pitch * H;
        GOFPID_TRY(alloc_zero((void **)&B.labels, B.labels_stream * sizeof(int32_t) * S));
        GOFPID_TRY(alloc_zero((void **)&B.slot, B.labels_stream * sizeof(int32_t) * S));
        B.bit_pitch = (int)round_up((size_t)(W + 31) / 32 + 2, 4);
        B.bit_stream = (size_t)B.bit_pitch * (H + 2);
        GOFPID_TRY(alloc_zero((void **)&B.bitimg, B.bit_stream * sizeof(uint32_t) * S));
        GOFPID_TRY(run_list_alloc(&B.runs, H, B.bit_pitch, S));
        GOFPID_TRY(alloc_zero((void **)&B.comp, (size_t)B.max_blobs * 8 * sizeof(int32_t) * S));
        GOFPID_TRY(alloc_zero((void **)&B.info, sizeof(gofpid_frame_info) * S));
        GOFPID_TRY(alloc_zero((void **)&B.kept, sizeof(gofpid_blob) * B.max_kept * S));
        GOFPID_TRY(alloc_zero((void **)&B.points, sizeof(int32_t) * 2 * B.max_points * S));
        GOFPID_CUDA_TRY(cudaHostAlloc(&F.h_kept, sizeof(gofpid_blob) * B.max_kept * S, cudaHostAllocMapped));
        GOFPID_CUDA_TRY(cudaHostAlloc(&F.h_points, sizeof(int32_t) * 2 * B.max_points * S, cudaHostAllocMapped));
        GOFPID_CUDA_TRY(cudaHostGetDevicePointer((void **)&F.dh_info, F.h_info, 0));
        GOFPID_CUDA_TRY(cudaHostGetDevicePointer((void **)&F.dh_kept, F.h_kept, 0));
        GOFPID_CUDA_TRY(cudaHostGetDevicePointer((void **)&F.dh_points, F.h_points, 0));
    }
    return reset_state(p);
}

// ---- stage A: tick, blur + detect -> the slot's detector mask -------------------------------------
// (the only part of a step that is sequential in time: the background model)
int enqueue_stage_a(gofpid_pipeline *p, FrameSlot &F, const FrameSrc &src, cudaStream_t s, bool prof,
                    int *n_launches)
{
    int launches = 0;
    if (prof) GOFPID_CUDA_TRY(cudaEventRecord(p->ev[0], s));
    const bool mog2 = p->cfg.detector == GOFPID_DETECT_MOG2;
    if (!mog2) { // (MOG2 derives its per-frame scalars inside the fused kernel)
        p->tick_args.src = src;
        GOFPID_TRY(launch_tick(p->tick_args, s));
    }
    if (prof) GOFPID_CUDA_TRY(cudaEventRecord(p->ev[4], s));
    DetectLaunch L;
    L.H = p->H; L.W = p->W; L.C = p->C; L.n_streams = p->n_streams;
    L.src = src;
    L.blur = p->blur;
    L.mask = ImageU8{p->cfg.n_morph > 0 ? F.d_mask_det : F.d_final, p->mask_pitch, p->mask_stream};
    L.tick = p->d_tick;
    L.stream = s;
    if (mog2) {
        GOFPID_TRY(launch_mog2(L, p->mog2, p->mog2p, p->d_tick));
        Mog2Group one;
        memset(&one, 0, sizeof(one));
        one.frame[0] = src.base; one.mask[0] = L.mask.base; one.nf = 1;
        one.aligned4 = mog2_group_aligned4(one, src);
        F.mog2_args = Mog2Args{L.blur, L.mask, p->mog2, p->mog2p, p->d_tick, src, one, p->H, p->W};
    }
    else if (p->cfg.detector == GOFPID_DETECT_KNN) GOFPID_TRY(launch_knn(L, p->knn, p->knnp, p->d_pow2tab, p->tune.knn_scalar != 0));
    else GOFPID_TRY(launch_fd(L, p->fd, p->fd_thr));
    launches += detect_launches_per_step(p->cfg.detector);
    if (prof) GOFPID_CUDA_TRY(cudaEventRecord(p->ev[1], s));
    if (n_launches) *n_launches = launches;
    return GOFPID_OK;
}

// ---- stage B: morphology -> the slot's final mask, then run extraction, labelling, border tracing ---
// (depends only on its own frame's detector mask, so the stage Bs of consecutive frames overlap
// each other and the next frames' stage A)
int enqueue_stage_b(gofpid_pipeline *p, FrameSlot &F, cudaStream_t s, bool prof, int *n_launches,
                    const cudaGraphConditionalHandle *cond = nullptr)
{
    int nb = 0, nm = 0;
    if (p->cfg.n_morph > 0) {
        const bool bits = p->morph.n_levels > 0 && p->cfg.find_blobs;
        GOFPID_TRY(launch_morph(p->morph, ImageU8{F.d_mask_det, p->mask_pitch, p->mask_stream},
                                ImageU8{F.d_final, p->mask_pitch, p->mask_stream},
                                p->H, p->W, p->n_streams, s, bits ? F.blobs.bitimg : nullptr,
                                F.blobs.bit_pitch, F.blobs.bit_stream));
        nm = 1;
    }
    if (prof) GOFPID_CUDA_TRY(cudaEventRecord(p->ev[2], s));
    if (p->cfg.find_blobs)
        GOFPID_TRY(launch_blobs(F.blobs, ImageU8{F.d_final, p->mask_pitch, p->mask_stream}, p->H, p->W,
                                p->n_streams, nullptr, s, &nb, p->cfg.n_morph > 0 && p->morph.n_levels > 0, cond));
    if (p->cfg.find_blobs) {
        // results travel to pinned host memory as part of the step (no host round trip to learn the
        // sizes first), written by a kernel: exactly what the frame produced, and the copy engine that
        // brings the frames in is left alone (three fixed-size copy nodes here cost the end-to-end
        // frame rate 6 us per 1080p frame)
        GOFPID_TRY(launch_results_to_host(F.blobs, F.dh_info, F.dh_kept, F.dh_points, p->n_streams, s));
        nb++;
    }
    if (n_launches) *n_launches = nb + nm;
    return GOFPID_OK;
}

// Captures the two stage graphs of a slot (on the pipeline's own streams; nothing executes).
int build_graphs(gofpid_pipeline *p, FrameSlot &F, const FrameSrc &src)
{
    int la = 0, lb = 0;
    GOFPID_CUDA_TRY(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_stage_a(p, F, src, p->stream, false, &la);
    cudaError_t e = cudaStreamEndCapture(p->stream, &F.graph_a);
    if (rc != GOFPID_OK) return rc;
    if (e != cudaSuccess) { set_error("graph capture (stage A) failed: %s", cudaGetErrorString(e)); return GOFPID_ERR_CUDA; }
    size_t n = 0;
    GOFPID_CUDA_TRY(cudaGraphGetNodes(F.graph_a, nullptr, &n));
    std::vector<cudaGraphNode_t> nodes(n);
    GOFPID_CUDA_TRY(cudaGraphGetNodes(F.graph_a, nodes.data(), &n));
    F.tick_node = nullptr;
    for (size_t i = 0; i < n; i++) {
        cudaGraphNodeType t;
        GOFPID_CUDA_TRY(cudaGraphNodeGetType(nodes[i], &t));
        if (t != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp;
        GOFPID_CUDA_TRY(cudaGraphKernelNodeGetParams(nodes[i], &kp));
        const void *want = p->cfg.detector == GOFPID_DETECT_MOG2 ? mog2_kernel_func(p->C, p->H, p->W, p->mog2p) : tick_kernel_func();
        if (kp.func == want) { F.tick_node = nodes[i]; F.src_node_params = kp; break; }
    }
    if (!F.tick_node) { set_error("graph capture: frame-source node not found"); return GOFPID_ERR_CUDA; }
    GOFPID_CUDA_TRY(cudaGraphInstantiate(&F.exec_a, F.graph_a, 0));
    if (p->cfg.find_blobs || p->cfg.n_morph > 0) {
        const bool use_cond = p->cfg.find_blobs && !p->tune.no_cond;
        GOFPID_CUDA_TRY(cudaGraphCreate(&F.graph_b, 0));
        if (use_cond)
            GOFPID_CUDA_TRY(cudaGraphConditionalHandleCreate(&F.cond_b, F.graph_b, 0, cudaGraphCondAssignDefault));
        cudaStream_t sb = p->stream_b[&F - p->slots];
        GOFPID_CUDA_TRY(cudaStreamBeginCaptureToGraph(sb, F.graph_b, nullptr, nullptr, 0,
                                                      cudaStreamCaptureModeThreadLocal));
        rc = enqueue_stage_b(p, F, sb, false, &lb, use_cond ? &F.cond_b : nullptr);
        cudaGraph_t captured = nullptr;
        e = cudaStreamEndCapture(sb, &captured);
        if (rc != GOFPID_OK) return rc;
        if (e != cudaSuccess) { set_error("graph capture (stage B) failed: %s", cudaGetErrorString(e)); return GOFPID_ERR_CUDA; }
        GOFPID_CUDA_TRY(cudaGraphInstantiate(&F.exec_b, F.graph_b, 0));
        if (use_cond) {
            int lb2 = 0;
            GOFPID_CUDA_TRY(cudaGraphCreate(&F.graph_b2, 0));
            GOFPID_CUDA_TRY(cudaStreamBeginCaptureToGraph(sb, F.graph_b2, nullptr, nullptr, 0,
                                                          cudaStreamCaptureModeThreadLocal));
            rc = enqueue_stage_b(p, F, sb, false, &lb2, nullptr);
            captured = nullptr;
            e = cudaStreamEndCapture(sb, &captured);
            if (rc != GOFPID_OK) return rc;
            if (e != cudaSuccess) { set_error("graph capture (stage B, unconditional) failed: %s", cudaGetErrorString(e)); return GOFPID_ERR_CUDA; }
            GOFPID_CUDA_TRY(cudaGraphInstantiate(&F.exec_b2, F.graph_b2, 0));
            p->launches_b2 = lb2;
        }
    }
    p->graph_launches = la + lb;
    p->launches_b = lb;
    return GOFPID_OK;
}

// One time step: stage A on `s`, stage B on the internal stream.
int run_step(gofpid_pipeline *p, const uint8_t *d_frames, size_t pitch, size_t stream_stride,
             cudaStream_t s)
{
    NvtxRange nvtx_range("gofpid: step (stage A + stage B enqueue)");
    const FrameSrc src{d_frames, pitch, stream_stride};
    FrameSlot &F = p->slots[p->frames % p->depth];
    int la = 0, lb = 0;
    // the slot's previous frame (t - DEPTH) must be done with its mask before it is overwritten
    if (F.used) GOFPID_CUDA_TRY(cudaStreamWaitEvent(s, F.ev_blobs, 0));
    if (p->profiling) {
        // everything on one stream, with events between the stages
        GOFPID_TRY(enqueue_stage_a(p, F, src, s, true, &la));
        GOFPID_TRY(enqueue_stage_b(p, F, s, true, &lb));
        GOFPID_CUDA_TRY(cudaEventRecord(p->ev[3], s));
        GOFPID_CUDA_TRY(cudaEventRecord(F.ev_mask, s));
        GOFPID_CUDA_TRY(cudaEventRecord(F.ev_blobs, s));
    } else {
        // (the first DEPTH steps run eagerly: they also perform the one-time function setup)
        const bool graph = p->use_graph && p->frames >= p->depth;
        if (graph && !F.exec_a) GOFPID_TRY(build_graphs(p, F, src));
        if (graph) {
            void *params[10];
            if (p->cfg.detector == GOFPID_DETECT_MOG2) {
                F.mog2_args.src = src;
                F.mog2_args.grp.frame[0] = src.base;
                F.mog2_args.grp.aligned4 = mog2_group_aligned4(F.mog2_args.grp, src);
                mog2_kernel_params(F.mog2_args, params);
            } else {
                p->tick_args.src = src;
                tick_kernel_params(p->tick_args, params);
            }
            cudaKernelNodeParams kp = F.src_node_params; // launch geometry as captured, new arguments
            kp.kernelParams = params; kp.extra = nullptr;
            GOFPID_CUDA_TRY(cudaGraphExecKernelNodeSetParams(F.exec_a, F.tick_node, &kp));
            GOFPID_CUDA_TRY(cudaGraphLaunch(F.exec_a, s));
            la = p->graph_launches;
        } else {
            GOFPID_TRY(enqueue_stage_a(p, F, src, s, false, &la));
        }
        GOFPID_CUDA_TRY(cudaEventRecord(F.ev_mask, s));
        if (p->cfg.find_blobs || p->cfg.n_morph > 0) {
            cudaStream_t sb = p->stream_b[&F - p->slots];
            GOFPID_CUDA_TRY(cudaStreamWaitEvent(sb, F.ev_mask, 0));
            const bool uncond = p->from_host && F.exec_b2;
            if (graph) {
                GOFPID_CUDA_TRY(cudaGraphLaunch(uncond ? F.exec_b2 : F.exec_b, sb));
                if (uncond) la += p->launches_b2 - p->launches_b;
            } else {
                GOFPID_TRY(enqueue_stage_b(p, F, sb, false, &lb));
            }
            GOFPID_CUDA_TRY(cudaEventRecord(F.ev_blobs, sb));
        } else {
            GOFPID_CUDA_TRY(cudaEventRecord(F.ev_blobs, s));
        }
    }
    p->launch_count = la + lb;
    p->launches_total += la + lb;
    F.used = true;
    F.fetched = false;
    F.frame_index = (int)std::min<long>(++p->frames, 0x7fffffffL);
    p->selected = 0;
    return GOFPID_OK;
}

// g consecutive time steps (MOG2) through ONE stage-A launch: the fused kernel keeps every tile's modes
// in shared memory while it walks the tile through the g frames, so the model is read and written once
// per group instead of once per frame and the launch / ramp / tail of the grid is paid once.  The g
// stage Bs then start together, each on its slot's stream.  frames[f]: frame f of stream 0.
int run_group(gofpid_pipeline *p, const uint8_t *const *frames, int g, size_t pitch, size_t stream_stride,
              cudaStream_t s)
{
    NvtxRange nvtx_range("gofpid: frame group (MOG2 launch + stage Bs)");
    if (g == 1 || p->cfg.detector != GOFPID_DETECT_MOG2 || !p->use_graph || p->frames < p->depth ||
        g > p->depth || g > MOG2_MAX_GROUP) {
        for (int f = 0; f < g; f++) GOFPID_TRY(run_step(p, frames[f], pitch, stream_stride, s));
        return GOFPID_OK;
    }
    const bool stage_b = p->cfg.find_blobs || p->cfg.n_morph > 0;
    Mog2Group grp;
    memset(&grp, 0, sizeof(grp));
    grp.nf = g;
    FrameSlot *slot[MOG2_MAX_GROUP];
    for (int f = 0; f < g; f++) {
        FrameSlot &F = p->slots[(p->frames + f) % p->depth];
        slot[f] = &F;
        // the slot's previous frame must be done with its masks before they are overwritten
        if (F.used) GOFPID_CUDA_TRY(cudaStreamWaitEvent(s, F.ev_blobs, 0));
        if (!F.exec_a) GOFPID_TRY(build_graphs(p, F, FrameSrc{frames[f], pitch, stream_stride}));
        grp.frame[f] = frames[f];
        grp.mask[f] = p->cfg.n_morph > 0 ? F.d_mask_det : F.d_final;
    }
    DetectLaunch L;
    L.H = p->H; L.W = p->W; L.C = p->C; L.n_streams = p->n_streams;
    L.src = FrameSrc{frames[0], pitch, stream_stride};
    L.blur = p->blur;
    L.mask = ImageU8{grp.mask[0], p->mask_pitch, p->mask_stream};
    L.tick = p->d_tick;
    L.stream = s;
    if (p->profiling) { // everything on one stream, events around the group's detector launch
        GOFPID_CUDA_TRY(cudaEventRecord(p->ev[0], s));
        GOFPID_CUDA_TRY(cudaEventRecord(p->ev[4], s));
        GOFPID_TRY(launch_mog2(L, p->mog2, p->mog2p, p->d_tick, &grp));
        GOFPID_CUDA_TRY(cudaEventRecord(p->ev[1], s));
        int lb = 0;
        for (int f = 0; f < g; f++) {
            FrameSlot &F = *slot[f];
            if (stage_b) GOFPID_TRY(enqueue_stage_b(p, F, s, f == 0, &lb));
            else if (f == 0) GOFPID_CUDA_TRY(cudaEventRecord(p->ev[2], s));
            GOFPID_CUDA_TRY(cudaEventRecord(F.ev_mask, s));
            GOFPID_CUDA_TRY(cudaEventRecord(F.ev_blobs, s));
            F.used = true;
            F.fetched = false;
            F.frame_index = (int)std::min<long>(++p->frames, 0x7fffffffL);
        }
        GOFPID_CUDA_TRY(cudaEventRecord(p->ev[3], s));
        p->launch_count = 1 + lb;
        p->launches_total += 1 + (long long)g * lb;
        p->selected = 0;
        return GOFPID_OK;
    }
    GOFPID_TRY(launch_mog2(L, p->mog2, p->mog2p, p->d_tick, &grp));
    GOFPID_CUDA_TRY(cudaEventRecord(slot[0]->ev_mask, s));
    for (int f = 0; f < g; f++) {
        FrameSlot &F = *slot[f];
        if (f > 0) GOFPID_CUDA_TRY(cudaEventRecord(F.ev_mask, s)); // (same point of the stream: every slot has its own event)
        if (stage_b) {
            cudaStream_t sb = p->stream_b[&F - p->slots];
            GOFPID_CUDA_TRY(cudaStreamWaitEvent(sb, slot[0]->ev_mask, 0));
            GOFPID_CUDA_TRY(cudaGraphLaunch(p->from_host && F.exec_b2 ? F.exec_b2 : F.exec_b, sb));
            GOFPID_CUDA_TRY(cudaEventRecord(F.ev_blobs, sb));
        } else {
            GOFPID_CUDA_TRY(cudaEventRecord(F.ev_blobs, s));
        }
        F.used = true;
        F.fetched = false;
        F.frame_index = (int)std::min<long>(++p->frames, 0x7fffffffL);
    }
    const int lbg = p->from_host && slot[0]->exec_b2 ? p->launches_b2 : p->launches_b;
    p->launch_count = 1 + lbg; // (kernels of the newest frame: its share of the group launch + stage B)
    p->launches_total += 1 + (long long)g * lbg;
    p->selected = 0;
    return GOFPID_OK;
}

// Brings the per-stream headers, kept blobs and contour points of a slot's frame to the host.
int fetch_results(gofpid_pipeline *p, FrameSlot *F)
{
    NvtxRange nvtx_range("gofpid: wait for a frame's results");
    if (!F) { set_error("no frame has been processed yet"); return GOFPID_ERR_INVALID; }
    if (F->fetched) return GOFPID_OK;
    const int S = p->n_streams;
    GOFPID_CUDA_TRY(cudaEventSynchronize(F->ev_blobs));
    if (!p->cfg.find_blobs) {
        for (int z = 0; z < S; z++) { memset(&F->h_info[z], 0, sizeof(gofpid_frame_info)); F->h_info[z].frame_index = F->frame_index; }
        F->fetched = true;
        return GOFPID_OK;
    }
    // headers, records and points arrived with the step (results_to_host_kernel, enqueue_stage_b)
    for (int z = 0; z < S; z++) F->h_info[z].frame_index = F->frame_index;
    F->fetched = true;
    return GOFPID_OK;
}

// A frame that exceeded the handle's per-slot blob capacities (noise without morphology, an illumination
// step, a fractal outline ...) is not an error for the reference, so it is not one here: blob creation
// runs once more on the slot's still-resident final mask (and bit image) with buffers of the size the
// frame needs -- the overflowed pass counted the components, the next one counts kept blobs and contour
// points exactly -- while the temporal model is untouched.  Rare and synchronous.
int rerun_big(gofpid_pipeline *p, FrameSlot *F, int z)
{
    NvtxRange nvtx_range("gofpid: re-run blob creation with larger buffers");
    if (p->big.slot == F && p->big.frame_index == F->frame_index && p->big.stream == z) return GOFPID_OK;
    p->big.slot = nullptr;
    const int H = p->H, W = p->W;
    const gofpid_frame_info first = F->h_info[z];
    BlobBuffers B = F->blobs; // the slot's run list, label / slot images and bit image, stream z only
    B.labels += (size_t)z * B.labels_stream;
    B.slot += (size_t)z * B.labels_pitch * H;
    B.bitimg += (size_t)z * B.bit_stream;
    {
        RunList &R = B.runs;
        const size_t mr = (size_t)(R.max_runs ? R.max_runs : 8);
        R.xs += z * mr; R.xe += z * mr; R.y += z * mr; R.label += z * mr; R.byc += z * mr;
        R.rowstart += (size_t)z * R.row_pitch; R.rowend += (size_t)z * R.row_pitch;
        R.count += z; R.n_labelled += z;
    }
    const ImageU8 mask{F->d_final + (size_t)z * p->mask_stream, p->mask_pitch, p->mask_stream};
    const long worst_comp = (long)((H + 1) / 2) * ((W + 1) / 2);
    long need_blobs = std::max<long>(first.n_components, B.max_blobs);
    long need_kept = std::max<long>(2L * first.n_kept, B.max_kept);
    long need_points = std::max<long>(2L * first.n_points, 4L * B.max_points);
    cudaStream_t s = p->stream_c;
    int rc = GOFPID_OK;
    for (int pass = 0; pass < 4; pass++) {
        need_blobs = std::min(need_blobs, worst_comp);
        need_kept = std::min(need_kept, need_blobs);
        need_points = std::min<long>(need_points, 0x3fffffffL);
        B.max_blobs = (int)need_blobs; B.max_kept = (int)need_kept; B.max_points = (int)need_points;
        B.comp = nullptr; B.kept = nullptr; B.points = nullptr; B.info = nullptr;
        gofpid_frame_info fi;
        memset(&fi, 0, sizeof(fi));
        if (cudaMalloc(&B.comp, (size_t)B.max_blobs * 8 * sizeof(int32_t)) != cudaSuccess ||
            cudaMalloc(&B.kept, (size_t)B.max_kept * sizeof(gofpid_blob)) != cudaSuccess ||
            cudaMalloc(&B.points, (size_t)B.max_points * 2 * sizeof(int32_t)) != cudaSuccess ||
            cudaMalloc(&B.info, sizeof(gofpid_frame_info)) != cudaSuccess) {
            cudaGetLastError();
            set_error("out of device memory re-running a frame with %ld components / %ld blobs / %ld contour points",
                      need_blobs, need_kept, need_points);
            rc = GOFPID_ERR_NOMEM;
        }
        if (rc == GOFPID_OK && cudaMemsetAsync(B.info, 0, sizeof(gofpid_frame_info), s) != cudaSuccess) rc = GOFPID_ERR_CUDA;
        int nl = 0;
        if (rc == GOFPID_OK) rc = launch_blobs(B, mask, H, W, 1, nullptr, s, &nl, true, nullptr, 64);
        if (rc == GOFPID_OK && (cudaMemcpyAsync(&fi, B.info, sizeof(fi), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                                cudaStreamSynchronize(s) != cudaSuccess)) {
            set_error("re-running blob creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = GOFPID_ERR_CUDA;
        }
        p->launches_total += nl;
        const bool fits = rc == GOFPID_OK && (fi.overflow & 7) == 0;
        if (fits) {
            const int nk = std::min(fi.n_kept, B.max_kept), np = std::min(fi.n_points, B.max_points);
            p->big.kept.resize((size_t)std::max(nk, 1));
            p->big.points.resize((size_t)std::max(np, 1) * 2);
            if ((nk && cudaMemcpy(p->big.kept.data(), B.kept, (size_t)nk * sizeof(gofpid_blob), cudaMemcpyDeviceToHost) != cudaSuccess) ||
                (np && cudaMemcpy(p->big.points.data(), B.points, (size_t)np * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess)) {
                set_error("fetching the re-run blob lists failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = GOFPID_ERR_CUDA;
            }
            fi.frame_index = F->frame_index;
            p->big.info = fi;
        }
        cudaFree(B.comp); cudaFree(B.kept); cudaFree(B.points); cudaFree(B.info);
        if (rc != GOFPID_OK) return rc;
        if (fits) {
            p->big.slot = F; p->big.frame_index = F->frame_index; p->big.stream = z;
            return GOFPID_OK;
        }
        // what the pass counted is exact for whatever fitted before it; grow and go again
        if (fi.overflow & 1) need_blobs = std::max<long>(fi.n_components, 2 * need_blobs);
        if (fi.overflow & 2) need_kept = std::max<long>(fi.n_kept, 2 * need_kept);
        if (fi.overflow & 4) need_points = std::max<long>(fi.n_points, 2 * need_points);
        if (fi.overflow & 1) { need_kept = std::max(need_kept, need_blobs / 8); need_points = std::max(need_points, 4 * need_points); }
    }
    set_error("blob creation did not fit after four passes (components %ld, blobs %ld, points %ld)", need_blobs, need_kept, need_points);
    return GOFPID_ERR_OVERFLOW;
}

// Where the selected frame's blob records and contour points of stream z are (the slot's pinned mirrors,
// or the re-run's vectors for a frame that exceeded the slot capacities).
struct BlobView {
    gofpid_frame_info fi;
    const gofpid_blob *kb;
    const int32_t *pts;
    int nk;
};
int blob_view(gofpid_pipeline *p, FrameSlot *F, int z, BlobView *v)
{
    GOFPID_TRY(fetch_results(p, F));
    const BlobBuffers &B = F->blobs;
    v->fi = F->h_info[z];
    if (v->fi.overflow & 8) {
        set_error("internal error: a border did not close (frame %d, stream %d)", F->frame_index, z);
        return GOFPID_ERR_OVERFLOW;
    }
    if (v->fi.overflow) {
        GOFPID_TRY(rerun_big(p, F, z));
        v->fi = p->big.info;
        v->kb = p->big.kept.data();
        v->pts = p->big.points.data();
        v->nk = p->big.info.n_kept;
        return GOFPID_OK;
    }
    v->nk = std::min(v->fi.n_kept, B.max_kept);
    v->kb = F->h_kept + (size_t)z * B.max_kept;
    v->pts = F->h_points + (size_t)z * B.max_points * 2;
    return GOFPID_OK;
}

// Blobs of the selected frame in the reference's order (descending raster order of each component's
// first pixel; blobs enclosed by another blob were flagged on the device: RETR_EXTERNAL drops them).
// n_blobs / n_points always receive the sizes; GOFPID_ERR_INVALID when they exceed the capacities.
int ordered_blobs(gofpid_pipeline *p, FrameSlot *F, int z, gofpid_blob *h_blobs, int blob_capacity, int32_t *h_points,
                  int point_capacity, int *n_blobs, int *n_points)
{
    BlobView v;
    GOFPID_TRY(blob_view(p, F, z, &v));
    std::vector<int> order;
    order.reserve(v.nk);
    for (int i = 0; i < v.nk; i++)
        if (!v.kb[i].nested) order.push_back(i);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return v.kb[a].first > v.kb[b].first; });
    const int nout = (int)order.size();
    long total = 0;
    for (int i = 0; i < nout; i++) total += v.kb[order[i]].n_points;
    if (n_blobs) *n_blobs = nout;
    if (n_points) *n_points = (int)total;
    if (!h_blobs && !h_points) return GOFPID_OK; // size query
    if (nout > blob_capacity || total > point_capacity) {
        set_error("output capacity too small: need %d blobs, %ld points", nout, total);
        return GOFPID_ERR_INVALID;
    }
    int off = 0;
    for (int i = 0; i < nout; i++) {
        gofpid_blob b = v.kb[order[i]];
        if (h_points) memcpy(h_points + (size_t)off * 2, v.pts + (size_t)b.point_offset * 2, sizeof(int32_t) * 2 * b.n_points);
        b.point_offset = off;
        off += b.n_points;
        if (h_blobs) h_blobs[i] = b;
    }
    return GOFPID_OK;
}

// Host frames -> the staging buffer of the slot `ahead` steps after the next one (asynchronous when
// the source is pinned).
int stage_host_frames(gofpid_pipeline *p, int ahead, const uint8_t *h_frames, size_t pitch, size_t stream_stride)
{
    NvtxRange nvtx_range("gofpid: upload frame");
    const size_t row = (size_t)p->W * p->Csrc;
    if (pitch < row) { set_error("pitch %zu smaller than a row (%zu bytes)", pitch, row); return GOFPID_ERR_INVALID; }
    FrameSlot &F = p->slots[(p->frames + ahead) % p->depth];
    if (!p->frames_pitch) {
        // rows packed back to back whenever the kernels' 4-byte loads allow it: a dense host frame then arrives
        // with ONE linear copy.  (A pitch rounded up to 128 made every 720-pixel BGR or 320-pixel gray frame a
        // 2-D copy, which the copy engine describes row by row: 13 k instead of 40 k stream-frames/s end to end
        // for 256 streams of 720 x 576 x 3.)
        p->frames_pitch = round_up(row, 4);
        p->frames_stream = p->frames_pitch * p->H;
    }
    if (!F.d_frames) {
        // all slots at once (no allocation inside a running pipeline), one block: consecutive slots are
        // consecutive in memory, so a group of frames can arrive with one copy (stage_host_group)
        const size_t per_slot = p->frames_stream * p->n_streams;
        GOFPID_CUDA_TRY(cudaMalloc(&p->d_frames_base, per_slot * p->depth));
        for (int si = 0; si < p->depth; si++) p->slots[si].d_frames = p->d_frames_base + per_slot * si;
    }
    bool pinned = h_frames >= p->pinned_lo && h_frames < p->pinned_hi;
    if (!pinned) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, h_frames) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
        else cudaGetLastError();
        if (pinned) { // remember the frame's own bytes: clips are usually one allocation walked frame by frame
            p->pinned_lo = h_frames;
            p->pinned_hi = h_frames + (p->n_streams > 1 ? stream_stride * (size_t)p->n_streams : pitch * (size_t)p->H);
        }
    }
    const uint8_t *src = h_frames;
    size_t spitch = pitch, sstride = stream_stride;
    cudaStream_t sh = p->stream_h;
    // stage A of the slot's previous frame must have finished reading the staging buffer
    if (F.used) GOFPID_CUDA_TRY(cudaStreamWaitEvent(sh, F.ev_mask, 0));
    if (!pinned) {
        // pageable memory: through the slot's own page-locked buffer, filled by the caller and the copier
        // threads; the upload of this frame then overlaps the filling of the next slot's buffer
        if (!F.h_stage) GOFPID_CUDA_TRY(cudaMallocHost(&F.h_stage, row * p->H * p->n_streams));
        if (!p->copier) {
            const unsigned hc = std::thread::hardware_concurrency();
            // (default: the host threads that fall to this GPU, half of them left to the caller, at most 7)
            int ngpu = 1;
            if (cudaGetDeviceCount(&ngpu) != cudaSuccess || ngpu < 1) { cudaGetLastError(); ngpu = 1; }
            const int dflt = std::max(1, std::min(7, (int)hc / (2 * ngpu) - 1));
            const int want = p->tune.copier_threads >= 0 ? p->tune.copier_threads : dflt;
            p->copier = new (std::nothrow) HostCopier(std::min(want, hc > 1 ? (int)hc - 1 : 0), p->tune.copier_spin_us,
                                                      p->tune.copier_stream != 0);
            if (!p->copier) { set_error("out of host memory"); return GOFPID_ERR_NOMEM; }
        }
        GOFPID_CUDA_TRY(cudaEventSynchronize(F.ev_h2d)); // the slot's previous upload out of h_stage is done
        if (p->n_streams == 1 && p->frames_pitch == row) {
            // one dense frame: cut into strips, and a strip travels while the next ones are still being copied
            // (a synchronous caller -- GOFPID.detect -- waits for copy + upload + step of every frame)
            const int nstrips = row * p->H >= (1u << 20) ? 4 : 1;
            cudaError_t err = cudaSuccess;
            p->copier->copy2d(F.h_stage, row, h_frames, pitch, row, (size_t)p->H, nstrips, [&](size_t r0, size_t nr) {
                if (err == cudaSuccess)
                    err = cudaMemcpyAsync(F.d_frames + r0 * row, F.h_stage + r0 * row, nr * row, cudaMemcpyHostToDevice, sh);
            });
            GOFPID_CUDA_TRY(err);
            GOFPID_CUDA_TRY(cudaEventRecord(F.ev_h2d, sh));
            return GOFPID_OK;
        }
        for (int z = 0; z < p->n_streams; z++)
            p->copier->copy2d(F.h_stage + (size_t)z * p->H * row, row, h_frames + z * stream_stride, pitch, row, (size_t)p->H);
        src = F.h_stage; spitch = row; sstride = row * p->H;
    }
    if ((sstride == spitch * p->H || p->n_streams == 1) && spitch == row && p->frames_pitch == row) {
        // dense on both sides: one linear copy (a 2-D copy of the same bytes is described row by row)
        GOFPID_CUDA_TRY(cudaMemcpyAsync(F.d_frames, src, row * p->H * p->n_streams, cudaMemcpyHostToDevice, sh));
    } else if (sstride == spitch * p->H || p->n_streams == 1) {
        GOFPID_CUDA_TRY(cudaMemcpy2DAsync(F.d_frames, p->frames_pitch, src, spitch, row,
                                          (size_t)p->H * p->n_streams, cudaMemcpyHostToDevice, sh));
    } else {
        for (int z = 0; z < p->n_streams; z++)
            GOFPID_CUDA_TRY(cudaMemcpy2DAsync(F.d_frames + z * p->frames_stream, p->frames_pitch,
                                              src + z * sstride, spitch, row, p->H, cudaMemcpyHostToDevice, sh));
    }
    GOFPID_CUDA_TRY(cudaEventRecord(F.ev_h2d, sh));
    return GOFPID_OK; // (the caller makes the detector's stream wait for F.ev_h2d when the frame's turn comes)
}

// g consecutive frames of a dense pinned clip -> g consecutive slots (starting `ahead` slots after the
// next one) with ONE copy: copies of a frame or three do not reach the link's bandwidth, half a ring of
// slots does (tools/pcie_bw.py); anything else goes frame by frame.
int stage_host_group(gofpid_pipeline *p, int ahead, int g, const uint8_t *h_frames, size_t frame_stride, size_t pitch,
                     size_t stream_stride)
{
    NvtxRange nvtx_range("gofpid: upload frame group");
    const size_t row = (size_t)p->W * p->Csrc;
    const int s0 = (int)((p->frames + ahead) % p->depth);
    bool one = g > 1 && s0 + g <= p->depth && p->d_frames_base && pitch >= row &&
               (p->n_streams == 1 || stream_stride == pitch * (size_t)p->H) &&
               frame_stride == pitch * (size_t)p->H * p->n_streams;
    if (one) {
        cudaPointerAttributes a0, a1;
        one = cudaPointerGetAttributes(&a0, h_frames) == cudaSuccess && a0.type == cudaMemoryTypeHost &&
              cudaPointerGetAttributes(&a1, h_frames + frame_stride * (size_t)g - 1) == cudaSuccess &&
              a1.type == cudaMemoryTypeHost;
        if (!one) cudaGetLastError();
    }
    if (!one) {
        for (int f = 0; f < g; f++)
            GOFPID_TRY(stage_host_frames(p, ahead + f, h_frames + (size_t)f * frame_stride, pitch, stream_stride));
        return GOFPID_OK;
    }
    cudaStream_t sh = p->stream_h;
    for (int f = 0; f < g; f++) { // stage A of the slots' previous frames must have finished reading the staging
        FrameSlot &F = p->slots[s0 + f];
        if (F.used) GOFPID_CUDA_TRY(cudaStreamWaitEvent(sh, F.ev_mask, 0));
    }
    if (pitch == row && p->frames_pitch == row)
        GOFPID_CUDA_TRY(cudaMemcpyAsync(p->slots[s0].d_frames, h_frames, row * p->H * p->n_streams * g,
                                        cudaMemcpyHostToDevice, sh));
    else
        GOFPID_CUDA_TRY(cudaMemcpy2DAsync(p->slots[s0].d_frames, p->frames_pitch, h_frames, pitch, row,
                                          (size_t)p->H * p->n_streams * g, cudaMemcpyHostToDevice, sh));
    for (int f = 0; f < g; f++) GOFPID_CUDA_TRY(cudaEventRecord(p->slots[s0 + f].ev_h2d, sh)); // (every slot: groups and copies need not align)
    return GOFPID_OK;
}

// Host frames -> the slot's staging buffer, then the step.
int submit_host(gofpid_pipeline *p, const uint8_t *h_frames, size_t pitch, size_t stream_stride)
{
    GOFPID_TRY(stage_host_frames(p, 0, h_frames, pitch, stream_stride));
    FrameSlot &F = p->slots[p->frames % p->depth];
    GOFPID_CUDA_TRY(cudaStreamWaitEvent(p->stream, F.ev_h2d, 0));
    p->from_host = true;
    const int rc = run_step(p, F.d_frames, p->frames_pitch, p->frames_stream, p->stream);
    p->from_host = false;
    return rc;
}

// A clip, in device or host memory: the pipelined loop (submit frames, collect the frame whose slot
// is needed next) that a caller of submit / select / blobs would write, without the per-call
// overhead -- and, for MOG2, with `group` frames per detector launch (run_group).
int run_sequence(gofpid_pipeline *p, const uint8_t *frames, bool host, int n_frames, size_t frame_stride,
                 size_t pitch, size_t stream_stride, cudaStream_t s, gofpid_blob *h_blobs, int blob_capacity,
                 int32_t *h_points, int point_capacity, int32_t *h_counts)
{
    NvtxRange nvtx_range("gofpid: clip");
    const int S = p->n_streams;
    long nb_total = 0, np_total = 0;
    int t = 0, c = 0; // frames of this clip submitted / collected
    // Blob lists go straight into the caller's buffers while they fit.  From the first frame that does
    // not, everything moves to vectors of the handle: the clip still runs to its end (the temporal
    // model and h_counts are complete), the call reports the sizes, and the caller fetches the lists
    // with gofpid_pipeline_sequence_results.
    bool spilled = false;
    p->seq_blobs.clear(); p->seq_points.clear();
    auto collect = [&](int upto) -> int { // every frame of the clip up to `upto` (inclusive)
        for (; c <= upto; c++) {
            FrameSlot *F = &p->slots[(p->frames - (t - c)) % p->depth];
            for (int z = 0; z < S; z++) {
                int nb = 0, np = 0;
                int rc = GOFPID_OK;
                if ((h_blobs || h_points) && !spilled) {
                    rc = ordered_blobs(p, F, z, h_blobs ? h_blobs + nb_total : nullptr, (int)(blob_capacity - nb_total),
                                       h_points ? h_points + (size_t)np_total * 2 : nullptr,
                                       (int)(point_capacity - np_total), &nb, &np);
                    if (rc == GOFPID_ERR_INVALID && (nb > blob_capacity - nb_total || np > point_capacity - np_total)) {
                        spilled = true;
                        if (h_blobs) p->seq_blobs.assign(h_blobs, h_blobs + nb_total);
                        if (h_points) p->seq_points.assign(h_points, h_points + (size_t)np_total * 2);
                    }
                } else if (!spilled) {
                    rc = ordered_blobs(p, F, z, nullptr, 0, nullptr, 0, &nb, &np); // counts only
                }
                if (spilled) {
                    GOFPID_TRY(ordered_blobs(p, F, z, nullptr, 0, nullptr, 0, &nb, &np));
                    p->seq_blobs.resize((size_t)nb_total + nb);
                    p->seq_points.resize(((size_t)np_total + np) * 2);
                    rc = ordered_blobs(p, F, z, p->seq_blobs.data() + nb_total, nb, p->seq_points.data() + (size_t)np_total * 2, np,
                                       &nb, &np);
                }
                if (rc != GOFPID_OK) return rc;
                h_counts[((size_t)c * S + z) * 2] = nb;
                h_counts[((size_t)c * S + z) * 2 + 1] = np;
                if (h_blobs || h_points) { nb_total += nb; np_total += np; }
            }
        }
        return GOFPID_OK;
    };
    int staged = 0; // frames of this clip whose host->device copy has been issued
    // Uploads run AHEAD of the step: the copy of a frame only needs its slot's staging buffer, which is free
    // once stage A of the slot's previous frame has run (a wait on the copy stream), not that frame's blob
    // lists -- so the copies of the next group(s) are queued BEFORE the host blocks collecting old results,
    // and the copy engine, which bounds the end-to-end rate, never runs dry (measured: 42.4 -> 39.x us per
    // 1080p frame against 39.0 for the copies alone).  GOFPID_H2D_CHUNK frames per copy; default: the frames
    // of one detector launch (half a ring of slots per copy was measured 8 % SLOWER: the first detector
    // launch of a chunk then waits for all of it).
    const int chunk_cfg = p->tune.h2d_chunk;
    const int per_copy = std::max(1, chunk_cfg > 0 ? std::min(chunk_cfg, p->depth / 2) : p->group);
    const int lead = std::min(2 * per_copy, p->depth / 2); // frames staged beyond the group about to run
    while (t < n_frames) {
        int g = 1;
        if (p->group > 1 && p->frames >= p->depth && p->use_graph) g = std::min(p->group, n_frames - t);
        const uint8_t *ptr[MOG2_MAX_GROUP];
        // (frame u goes into the slot of frame u - depth, whose stage A has been submitted when u - depth < t)
        while (host && staged < std::min(n_frames, std::min(t + g + lead, t + p->depth))) {
            const int s0 = (int)((p->frames + (staged - t)) % p->depth);
            int n = std::min(std::min(per_copy, n_frames - staged), p->depth - s0); // (a copy does not wrap around the ring)
            n = std::min(n, t + p->depth - staged);
            if (staged < t + g) n = std::min(n, t + g - staged); // (the group about to run: no more than it needs)
            GOFPID_TRY(stage_host_group(p, staged - t, n, frames + (size_t)staged * frame_stride, frame_stride, pitch, stream_stride));
            staged += n;
        }
        // frames t .. t+g-1 take the slots of frames t-depth .. t+g-1-depth: collect those first
        if (h_counts) GOFPID_TRY(collect(t + g - 1 - p->depth));
        if (host) GOFPID_CUDA_TRY(cudaStreamWaitEvent(p->stream, p->slots[(p->frames + g - 1) % p->depth].ev_h2d, 0));
        for (int f = 0; f < g; f++)
            ptr[f] = host ? p->slots[(p->frames + f) % p->depth].d_frames : frames + (size_t)(t + f) * frame_stride;
        p->from_host = host;
        const int rc = host ? run_group(p, ptr, g, p->frames_pitch, p->frames_stream, p->stream)
                            : run_group(p, ptr, g, pitch, stream_stride, s);
        p->from_host = false;
        GOFPID_TRY(rc);
        t += g;
    }
    if (h_counts) GOFPID_TRY(collect(n_frames - 1));
    if (spilled) {
        set_error("output capacity too small: the clip needs %ld blobs, %ld points (the lists are kept: "
                  "gofpid_pipeline_sequence_results)", nb_total, np_total);
        return GOFPID_ERR_INVALID;
    }
    return GOFPID_OK;
}

} // namespace

// =====================================================================================
// library
// =====================================================================================
extern "C" {

GOFPID_API const char *gofpid_version(void) { return "gofpid_b200 0.1.0 (sm_100a)"; }
GOFPID_API const char *gofpid_last_error(void) { return g_error; }

GOFPID_API int gofpid_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

GOFPID_API void gofpid_config_default(gofpid_config *cfg)
{
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->blur_kind = GOFPID_BLUR_GAUSSIAN; // pipeline.py:171-175
    cfg->blur_kw = cfg->blur_kh = 3;
    cfg->blur_sigma = 4.0; // cv.BORDER_DEFAULT (4) lands in sigmaX: pipeline.py:463-467
    cfg->detector = GOFPID_DETECT_MOG2;
    cfg->fd_threshold = 5.0;
    cfg->mog2_history = 500;
    cfg->mog2_var_threshold = 16.0f;
    cfg->mog2_var_threshold_gen = 9.0f;
    cfg->mog2_background_ratio = 0.9f;
    cfg->mog2_var_init = 15.0f;
    cfg->mog2_var_min = 4.0f;
    cfg->mog2_var_max = 75.0f;
    cfg->mog2_ct = 0.05f;
    cfg->knn_history = 500;
    cfg->knn_k = 2;
    cfg->knn_dist2_threshold = 400.0f;
    cfg->knn_rng_seed = 0;
    cfg->detect_shadows = 1;
    cfg->shadow_value = 127;
    cfg->shadow_tau = 0.5f;
    cfg->n_morph = 0; // the caller supplies structuring elements (pipeline.py:177-186)
    cfg->find_blobs = 1;
    cfg->blob_area_min = 100.0; // pipeline.py:555
    cfg->max_blobs = 1 << 16;
    cfg->max_kept = 4096;
    cfg->max_points = 0; // 0: sized from the frame (see gofpid_pipeline_create)
}

GOFPID_API int gofpid_host_alloc(size_t bytes, int write_combined, void **out)
{
    if (!out) { set_error("null output"); return GOFPID_ERR_INVALID; }
    *out = nullptr;
    if (gofpid_device_count() == 0) { set_error("no CUDA device available (no CPU fallback)"); return GOFPID_ERR_CUDA; }
    const unsigned flags = cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0u);
    if (cudaHostAlloc(out, bytes ? bytes : 1, flags) != cudaSuccess) {
        cudaGetLastError();
        set_error("cannot page-lock %zu bytes of host memory", bytes);
        return GOFPID_ERR_NOMEM;
    }
    return GOFPID_OK;
}
GOFPID_API int gofpid_host_copy2d(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t row_bytes,
                                  size_t n_rows, int n_threads)
{
    if ((!dst || !src) && row_bytes && n_rows) { set_error("null buffer"); return GOFPID_ERR_INVALID; }
    if (n_rows > 1 && (dst_pitch < row_bytes || src_pitch < row_bytes)) {
        set_error("pitch smaller than a row (%zu / %zu < %zu)", dst_pitch, src_pitch, row_bytes);
        return GOFPID_ERR_INVALID;
    }
    if (!row_bytes || !n_rows) return GOFPID_OK;
    static std::mutex mu;
    static HostCopier *copier = nullptr;
    static int copier_threads = -2;
    std::lock_guard<std::mutex> lk(mu);
    const unsigned hc = std::thread::hardware_concurrency();
    int want = n_threads >= 0 ? n_threads : std::max(1, std::min(7, (int)hc / 2 - 1));
    want = std::min(want, hc > 1 ? (int)hc - 1 : 0);
    if (!copier || copier_threads != want) {
        delete copier;
        copier = new (std::nothrow) HostCopier(want, 500, true);
        copier_threads = want;
        if (!copier) { set_error("out of host memory"); return GOFPID_ERR_NOMEM; }
    }
    // (four strips, as the host entry points use them: a strip is reported once all its chunks are in place)
    size_t next_row = 0;
    bool in_order = true;
    copier->copy2d(static_cast<uint8_t *>(dst), dst_pitch, static_cast<const uint8_t *>(src), src_pitch, row_bytes, n_rows, 4,
                   [&](size_t r0, size_t nr) { in_order &= r0 == next_row; next_row = r0 + nr; });
    if (!in_order || next_row != n_rows) { set_error("internal: strips reported out of order"); return GOFPID_ERR_INVALID; }
    return GOFPID_OK;
}

GOFPID_API int gofpid_host_free(void *ptr)
{
    if (ptr && cudaFreeHost(ptr) != cudaSuccess) { cudaGetLastError(); set_error("not a gofpid_host_alloc pointer"); return GOFPID_ERR_INVALID; }
    return GOFPID_OK;
}

GOFPID_API int gofpid_gaussian_taps_q8(int n, double sigma, uint16_t *h_taps)
{
    if (!h_taps) { set_error("null output"); return GOFPID_ERR_INVALID; }
    const int rc = gaussian_taps_q8(n, sigma, h_taps);
    if (rc) set_error("Gaussian kernel size must be odd and in [1, %d], got %d", GOFPID_MAX_KSIZE, n);
    return rc;
}

GOFPID_API int gofpid_pipeline_create(const gofpid_config *cfg, int H, int W, int C, int n_streams,
                                      int device, gofpid_pipeline **out)
{
    if (!cfg || !out) { set_error("null argument"); return GOFPID_ERR_INVALID; }
    *out = nullptr;
    if (H <= 0 || W <= 0 || n_streams <= 0) {
        set_error("invalid geometry H=%d W=%d n_streams=%d", H, W, n_streams);
        return GOFPID_ERR_INVALID;
    }
    if (C != 1 && C != 3) {
        set_error("frames must have 1 or 3 channels, got %d", C);
        return GOFPID_ERR_UNSUPPORTED;
    }
    if (cfg->n_morph < 0 || cfg->n_morph > GOFPID_MAX_MORPH_OPS) {
        set_error("mat_morph has %d steps; at most %d are supported", cfg->n_morph, GOFPID_MAX_MORPH_OPS);
        return GOFPID_ERR_UNSUPPORTED;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (libgofpid_b200 has no CPU fallback)");
        return GOFPID_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        set_error("device %d out of range [0, %d)", device, ndev);
        return GOFPID_ERR_INVALID;
    }
    gofpid_pipeline *p = new (std::nothrow) gofpid_pipeline();
    if (!p) { set_error("out of host memory"); return GOFPID_ERR_NOMEM; }
    p->cfg = *cfg;
    p->tune = read_tuning();
    if (p->cfg.max_points <= 0) { // contour points per frame: 32 frame perimeters' worth, 64 Ki .. 1 Mi
        const long want = 32L * (H + W);
        p->cfg.max_points = (int)(want < 65536 ? 65536 : (want > (1L << 20) ? (1L << 20) : want));
    }
    p->H = H; p->W = W; p->C = C; p->Csrc = C; p->n_streams = n_streams; p->device = device;
    if (cfg->convert != GOFPID_CONVERT_NONE) {
        if ((cfg->convert != GOFPID_CONVERT_BGR2GRAY && cfg->convert != GOFPID_CONVERT_RGB2GRAY) || C != 3) {
            delete p;
            set_error("convert must be BGR2GRAY or RGB2GRAY on 3-channel frames (got code %d, C = %d)", cfg->convert, C);
            return GOFPID_ERR_UNSUPPORTED;
        }
        p->C = 1; // the model, masks and blobs are those of the gray image
    }
    // own the structuring elements
    p->se_copies.resize(cfg->n_morph);
    for (int i = 0; i < cfg->n_morph; i++) {
        const gofpid_morph_op &o = cfg->morph[i];
        if (o.kernel && o.kh > 0 && o.kw > 0) {
            p->se_copies[i].assign(o.kernel, o.kernel + (size_t)o.kh * o.kw);
            p->cfg.morph[i].kernel = p->se_copies[i].data();
        }
    }
    DeviceGuard g(device);
    int rc = g.ok ? build(p) : GOFPID_ERR_CUDA;
    if (!g.ok) set_error("cannot select CUDA device %d", device);
    if (rc != GOFPID_OK) {
        free_all(p);
        delete p;
        cudaGetLastError();
        return rc;
    }
    *out = p;
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_destroy(gofpid_pipeline *p)
{
    if (!p) return GOFPID_OK;
    {
        DeviceGuard g(p->device);
        cudaDeviceSynchronize();
        free_all(p);
    }
    delete p;
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_reset(gofpid_pipeline *p)
{
    GOFPID_GUARD(p);
    GOFPID_CUDA_TRY(cudaDeviceSynchronize());
    return reset_state(p);
}

GOFPID_API int gofpid_pipeline_process_device(gofpid_pipeline *p, const uint8_t *d_frames,
                                              size_t pitch, size_t stream_stride, void *stream)
{
    GOFPID_GUARD(p);
    if (!d_frames) { set_error("null frame pointer"); return GOFPID_ERR_INVALID; }
    if (pitch < (size_t)p->W * p->Csrc) { set_error("pitch %zu smaller than a row (%d bytes)", pitch, p->W * p->Csrc); return GOFPID_ERR_INVALID; }
    return run_step(p, d_frames, pitch, stream_stride, (cudaStream_t)stream);
}

GOFPID_API int gofpid_pipeline_submit_host(gofpid_pipeline *p, const uint8_t *h_frames, size_t pitch,
                                           size_t stream_stride)
{
    GOFPID_GUARD(p);
    if (!h_frames) { set_error("null frame pointer"); return GOFPID_ERR_INVALID; }
    return submit_host(p, h_frames, pitch, stream_stride);
}

GOFPID_API int gofpid_pipeline_process_host(gofpid_pipeline *p, const uint8_t *h_frames,
                                            size_t pitch, size_t stream_stride)
{
    GOFPID_GUARD(p);
    if (!h_frames) { set_error("null frame pointer"); return GOFPID_ERR_INVALID; }
    GOFPID_TRY(submit_host(p, h_frames, pitch, stream_stride));
    return fetch_results(p, selected_slot(p));
}

// A whole clip in one call: the pipelined loop (submit frame t, collect frame t - lag) that a
// caller of submit_host / select / blobs would write, without the per-call overhead.
GOFPID_API int gofpid_pipeline_process_host_sequence(gofpid_pipeline *p, const uint8_t *h_frames, int n_frames,
                                                     size_t frame_stride, size_t pitch, size_t stream_stride,
                                                     gofpid_blob *h_blobs, int blob_capacity, int32_t *h_points,
                                                     int point_capacity, int32_t *h_counts)
{
    GOFPID_GUARD(p);
    if (!h_frames || n_frames < 0 || !h_counts) { set_error("bad argument"); return GOFPID_ERR_INVALID; }
    if (!p->cfg.find_blobs) { set_error("pipeline was created with find_blobs = 0"); return GOFPID_ERR_INVALID; }
    return run_sequence(p, h_frames, true, n_frames, frame_stride, pitch, stream_stride, p->stream, h_blobs,
                        blob_capacity, h_points, point_capacity, h_counts);
}

// The same for a clip that is already in device memory (frames frame_stride bytes apart); stage A is
// enqueued on `stream`.  h_counts == NULL: nothing is collected (the newest `depth` frames stay
// selectable), the call only enqueues work.
GOFPID_API int gofpid_pipeline_process_device_sequence(gofpid_pipeline *p, const uint8_t *d_frames, int n_frames,
                                                       size_t frame_stride, size_t pitch, size_t stream_stride,
                                                       void *stream, gofpid_blob *h_blobs, int blob_capacity,
                                                       int32_t *h_points, int point_capacity, int32_t *h_counts)
{
    GOFPID_GUARD(p);
    if (!d_frames || n_frames < 0) { set_error("bad argument"); return GOFPID_ERR_INVALID; }
    if (h_counts && !p->cfg.find_blobs) { set_error("pipeline was created with find_blobs = 0"); return GOFPID_ERR_INVALID; }
    if (pitch < (size_t)p->W * p->Csrc) { set_error("pitch %zu smaller than a row (%d bytes)", pitch, p->W * p->Csrc); return GOFPID_ERR_INVALID; }
    return run_sequence(p, d_frames, false, n_frames, frame_stride, pitch, stream_stride, (cudaStream_t)stream, h_blobs,
                        blob_capacity, h_points, point_capacity, h_counts);
}

#ifdef GOFPID_TRACE_DEBUG
// Debug aids of the instrumented build only (make debug; not part of the ABI in include/): a mark on
// `stream`, then per slot the milliseconds from the mark to the end of the slot's stage A and stage B
// (needs GOFPID_EVENT_TIMING=1).
GOFPID_API int gofpid_debug_mark(gofpid_pipeline *p, void *stream)
{
    GOFPID_GUARD(p);
    GOFPID_CUDA_TRY(cudaEventRecord(p->ev[5], (cudaStream_t)stream));
    return GOFPID_OK;
}
GOFPID_API int gofpid_debug_slot_times(gofpid_pipeline *p, float *ms2, int *frame_index)
{
    GOFPID_GUARD(p);
    GOFPID_CUDA_TRY(cudaDeviceSynchronize());
    for (int i = 0; i < p->depth; i++) {
        FrameSlot &F = p->slots[i];
        ms2[2 * i] = ms2[2 * i + 1] = -1.f;
        frame_index[i] = F.frame_index;
        if (!F.used) continue;
        if (cudaEventElapsedTime(&ms2[2 * i], p->ev[5], F.ev_mask) != cudaSuccess) cudaGetLastError();
        if (cudaEventElapsedTime(&ms2[2 * i + 1], p->ev[5], F.ev_blobs) != cudaSuccess) cudaGetLastError();
    }
    return GOFPID_OK;
}
#endif

GOFPID_API int gofpid_pipeline_join(gofpid_pipeline *p, void *stream)
{
    GOFPID_GUARD(p);
    for (int i = 0; i < p->depth; i++)
        if (p->slots[i].used) GOFPID_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, p->slots[i].ev_blobs, 0));
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_group(gofpid_pipeline *p)
{
    return p ? p->group : 0;
}

GOFPID_API int gofpid_pipeline_launch_total(gofpid_pipeline *p, long long *n_launches)
{
    GOFPID_GUARD(p);
    if (!n_launches) { set_error("null output pointer"); return GOFPID_ERR_INVALID; }
    *n_launches = p->launches_total;
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_depth(gofpid_pipeline *p)
{
    return p ? p->depth : 0;
}

GOFPID_API int gofpid_pipeline_select(gofpid_pipeline *p, int age)
{
    GOFPID_GUARD(p);
    if (age < 0 || age >= p->depth) { set_error("age must be in [0, %d): 0 is the newest frame", p->depth); return GOFPID_ERR_INVALID; }
    if (p->frames - 1 - age < 0) { set_error("no such frame yet"); return GOFPID_ERR_INVALID; }
    p->selected = age;
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_frame_info(gofpid_pipeline *p, int stream_idx, gofpid_frame_info *h_info)
{
    GOFPID_GUARD(p);
    GOFPID_TRY(check_stream_index(p, stream_idx));
    if (!h_info) { set_error("null output"); return GOFPID_ERR_INVALID; }
    FrameSlot *F = selected_slot(p);
    GOFPID_TRY(fetch_results(p, F));
    *h_info = F->h_info[stream_idx];
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_blobs(gofpid_pipeline *p, int stream_idx, gofpid_blob *h_blobs,
                                     int blob_capacity, int32_t *h_points, int point_capacity,
                                     int *n_blobs, int *n_points)
{
    GOFPID_GUARD(p);
    GOFPID_TRY(check_stream_index(p, stream_idx));
    if (!p->cfg.find_blobs) { set_error("pipeline was created with find_blobs = 0"); return GOFPID_ERR_INVALID; }
    FrameSlot *F = selected_slot(p);
    if (!F) { set_error("no frame has been processed yet"); return GOFPID_ERR_INVALID; }
    return ordered_blobs(p, F, stream_idx, h_blobs, blob_capacity, h_points, point_capacity, n_blobs, n_points);
}

GOFPID_API int gofpid_pipeline_sequence_results(gofpid_pipeline *p, gofpid_blob *h_blobs, int blob_capacity,
                                                int32_t *h_points, int point_capacity)
{
    GOFPID_GUARD(p);
    const size_t nb = p->seq_blobs.size(), np = p->seq_points.size() / 2;
    if ((h_blobs && nb > (size_t)std::max(blob_capacity, 0)) || (h_points && np > (size_t)std::max(point_capacity, 0))) {
        set_error("output capacity too small: need %zu blobs, %zu points", nb, np);
        return GOFPID_ERR_INVALID;
    }
    if (h_blobs && nb) memcpy(h_blobs, p->seq_blobs.data(), nb * sizeof(gofpid_blob));
    if (h_points && np) memcpy(h_points, p->seq_points.data(), np * 2 * sizeof(int32_t));
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_mask(gofpid_pipeline *p, int stream_idx, uint8_t *dst, size_t dst_pitch,
                                    int dst_is_device)
{
    GOFPID_GUARD(p);
    GOFPID_TRY(check_stream_index(p, stream_idx));
    if (!dst || dst_pitch < (size_t)p->W) { set_error("bad destination"); return GOFPID_ERR_INVALID; }
    FrameSlot *F = selected_slot(p);
    if (!F) { set_error("no frame has been processed yet"); return GOFPID_ERR_INVALID; }
    GOFPID_CUDA_TRY(cudaEventSynchronize(F->ev_blobs)); // (the final mask is written by stage B's first kernel)
    const uint8_t *src = F->d_final + (size_t)stream_idx * p->mask_stream;
    GOFPID_CUDA_TRY(cudaMemcpy2DAsync(dst, dst_pitch, src, p->mask_pitch, p->W, p->H,
                                      dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                      p->stream_c));
    GOFPID_CUDA_TRY(cudaStreamSynchronize(p->stream_c));
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_labels(gofpid_pipeline *p, int stream_idx, int32_t *dst, size_t dst_pitch,
                                      int dst_is_device)
{
    GOFPID_GUARD(p);
    GOFPID_TRY(check_stream_index(p, stream_idx));
    if (!p->cfg.find_blobs) { set_error("pipeline was created with find_blobs = 0"); return GOFPID_ERR_INVALID; }
    if (!dst || dst_pitch < (size_t)p->W * sizeof(int32_t) || dst_pitch % sizeof(int32_t)) { set_error("bad destination"); return GOFPID_ERR_INVALID; }
    FrameSlot *F = selected_slot(p);
    if (!F) { set_error("no frame has been processed yet"); return GOFPID_ERR_INVALID; }
    GOFPID_CUDA_TRY(cudaEventSynchronize(F->ev_blobs));
    cudaStream_t s = p->stream_c;
    int32_t *tmp = nullptr;
    int32_t *out = dst;
    size_t out_pitch = dst_pitch / sizeof(int32_t);
    if (!dst_is_device) {
        GOFPID_CUDA_TRY(cudaMalloc(&tmp, (size_t)p->W * p->H * sizeof(int32_t)));
        out = tmp; out_pitch = p->W;
    }
    const BlobBuffers &B = F->blobs;
    // (the hot path labels runs, not pixels: the label image is filled in here, on demand)
    int rc = launch_fill_labels(B.runs, stream_idx, B.info + stream_idx,
                                B.labels + (size_t)stream_idx * B.labels_stream, B.labels_pitch, s);
    if (rc == GOFPID_OK)
        rc = launch_canonical_labels(F->d_final + (size_t)stream_idx * p->mask_stream, p->mask_pitch,
                                     B.labels + (size_t)stream_idx * B.labels_stream, B.labels_pitch, out,
                                     out_pitch, p->H, p->W, s);
    if (rc == GOFPID_OK && !dst_is_device) {
        if (cudaMemcpy2DAsync(dst, dst_pitch, tmp, (size_t)p->W * sizeof(int32_t), (size_t)p->W * sizeof(int32_t),
                              p->H, cudaMemcpyDeviceToHost, s) != cudaSuccess) { set_error("label copy failed"); rc = GOFPID_ERR_CUDA; }
    }
    if (cudaStreamSynchronize(s) != cudaSuccess && rc == GOFPID_OK) { set_error("label export failed: %s", cudaGetErrorString(cudaGetLastError())); rc = GOFPID_ERR_CUDA; }
    if (tmp) cudaFree(tmp);
    return rc;
}

GOFPID_API int gofpid_pipeline_components(gofpid_pipeline *p, int stream_idx, int32_t *h_stats, int capacity,
                                          int *n_components)
{
    GOFPID_GUARD(p);
    GOFPID_TRY(check_stream_index(p, stream_idx));
    if (!p->cfg.find_blobs) { set_error("pipeline was created with find_blobs = 0"); return GOFPID_ERR_INVALID; }
    FrameSlot *F = selected_slot(p);
    GOFPID_TRY(fetch_results(p, F));
    const BlobBuffers &B = F->blobs;
    const gofpid_frame_info &fi = F->h_info[stream_idx];
    const int n = std::min(fi.n_components, B.max_blobs);
    if (n_components) *n_components = fi.n_components;
    if (!h_stats) return GOFPID_OK;
    if (n > capacity) { set_error("output capacity too small: need %d components", n); return GOFPID_ERR_INVALID; }
    std::vector<int32_t> rec((size_t)n * 8);
    if (n > 0)
        GOFPID_CUDA_TRY(cudaMemcpy(rec.data(), B.comp + (size_t)stream_idx * B.max_blobs * 8,
                                   rec.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    const int LP = (int)B.labels_pitch;
    for (int i = 0; i < n; i++) {
        const int32_t *r = &rec[(size_t)i * 8];
        const int fy = r[0] / LP, fx = r[0] - fy * LP;
        h_stats[i * 6 + 0] = fy * p->W + fx;
        h_stats[i * 6 + 1] = r[1]; h_stats[i * 6 + 2] = r[2];
        h_stats[i * 6 + 3] = r[3]; h_stats[i * 6 + 4] = r[4]; h_stats[i * 6 + 5] = r[5];
    }
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_mog2_model(gofpid_pipeline *p, int stream_idx, float *h_weight, float *h_var,
                                          float *h_mean, uint8_t *h_nmodes)
{
    GOFPID_GUARD(p);
    GOFPID_TRY(check_stream_index(p, stream_idx));
    if (p->cfg.detector != GOFPID_DETECT_MOG2) { set_error("detector is not MOG2"); return GOFPID_ERR_INVALID; }
    GOFPID_CUDA_TRY(cudaDeviceSynchronize());
    const int H = p->H, W = p->W, C = p->C;
    const size_t nt = (size_t)p->mog2.tiles_x * p->mog2.tiles_y, tpx = (size_t)MOG2_TILE_W * MOG2_TILE_H;
    const size_t tile_floats = (size_t)GOFPID_MOG2_NMIX * (2 + C) * tpx;
    std::vector<float> tiles(nt * tile_floats);
    std::vector<uint8_t> nm(nt * tpx);
    GOFPID_CUDA_TRY(cudaMemcpy(tiles.data(), p->mog2.planes + (size_t)stream_idx * nt * tile_floats,
                               tiles.size() * sizeof(float), cudaMemcpyDeviceToHost));
    GOFPID_CUDA_TRY(cudaMemcpy(nm.data(), p->mog2.nmodes + (size_t)stream_idx * nt * tpx, nm.size(),
                               cudaMemcpyDeviceToHost));
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t t = (size_t)(y / MOG2_TILE_H) * p->mog2.tiles_x + x / MOG2_TILE_W;
            const size_t l = (size_t)(y % MOG2_TILE_H) * MOG2_TILE_W + x % MOG2_TILE_W;
            const size_t d = (size_t)y * W + x;
            if (h_nmodes) h_nmodes[d] = nm[t * tpx + l];
            for (int k = 0; k < GOFPID_MOG2_NMIX; k++) {
                const float *mk = &tiles[t * tile_floats + (size_t)k * (2 + C) * tpx];
                if (h_weight) h_weight[d * 5 + k] = mk[l];
                if (h_var) h_var[d * 5 + k] = mk[tpx + l];
                if (h_mean)
                    for (int c = 0; c < C; c++) h_mean[(d * 5 + k) * C + c] = mk[(2 + c) * tpx + l];
            }
        }
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_knn_model(gofpid_pipeline *p, int stream_idx, uint8_t *h_samples, uint8_t *h_idx,
                                         uint8_t *h_nxt, int32_t *h_counters, uint64_t *rng_state)
{
    GOFPID_GUARD(p);
    GOFPID_TRY(check_stream_index(p, stream_idx));
    if (p->cfg.detector != GOFPID_DETECT_KNN) { set_error("detector is not KNN"); return GOFPID_ERR_INVALID; }
    GOFPID_CUDA_TRY(cudaDeviceSynchronize());
    const int H = p->H, W = p->W, ND = p->C + 1, PW = p->knn.PW;
    const size_t pe = p->knn.plane_elems, n = (size_t)H * W;
    std::vector<uint8_t> smp(pe * 21 * ND);
    std::vector<uint32_t> nxt(pe);
    std::vector<uint16_t> idx(pe);
    FrameTick t;
    GOFPID_CUDA_TRY(cudaMemcpy(smp.data(), p->knn.samples + (size_t)stream_idx * smp.size(), smp.size(), cudaMemcpyDeviceToHost));
    GOFPID_CUDA_TRY(cudaMemcpy(nxt.data(), p->knn.nxt + (size_t)stream_idx * pe, pe * 4, cudaMemcpyDeviceToHost));
    GOFPID_CUDA_TRY(cudaMemcpy(idx.data(), p->knn.idx + (size_t)stream_idx * pe, pe * 2, cudaMemcpyDeviceToHost));
    GOFPID_CUDA_TRY(cudaMemcpy(&t, p->d_tick, sizeof(t), cudaMemcpyDeviceToHost));
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t s = (size_t)y * PW + x, d = (size_t)y * W + x;
            if (h_samples)
                for (int k = 0; k < 21; k++)
                    for (int c = 0; c < ND; c++) {
                        // gray: every group of four pixels is (v0 v1 v2 v3 | f0 f1 f2 f3); colour: (c0 c1 c2 flag) per pixel
                        const size_t at = ND == 2 ? ((size_t)k * pe + (s & ~(size_t)3)) * 2 + 4 * c + (s & 3)
                                                  : ((size_t)k * pe + s) * ND + c;
                        h_samples[(d * 21 + k) * ND + c] = smp[at];
                    }
            if (h_idx) { h_idx[d] = idx[s] & 7; h_idx[n + d] = (idx[s] >> 3) & 7; h_idx[2 * n + d] = (idx[s] >> 6) & 7; }
            if (h_nxt) { h_nxt[d] = nxt[s] & 255; h_nxt[n + d] = (nxt[s] >> 8) & 255; h_nxt[2 * n + d] = (nxt[s] >> 16) & 255; }
        }
    if (h_counters) {
        // counters as OpenCV holds them AFTER the last frame
        for (int k = 0; k < 3; k++) {
            int c = t.knn_counter[k];
            if (t.nframes > 0) { c++; if (c >= t.knn_u[k]) c = 0; }
            h_counters[k] = c;
        }
    }
    if (rng_state) *rng_state = t.knn_rng;
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_fd_model(gofpid_pipeline *p, int stream_idx, uint8_t *h_prev, int *valid)
{
    GOFPID_GUARD(p);
    GOFPID_TRY(check_stream_index(p, stream_idx));
    if (p->cfg.detector != GOFPID_DETECT_FD) { set_error("detector is not FD"); return GOFPID_ERR_INVALID; }
    GOFPID_CUDA_TRY(cudaDeviceSynchronize());
    FrameTick t;
    GOFPID_CUDA_TRY(cudaMemcpy(&t, p->d_tick, sizeof(t), cudaMemcpyDeviceToHost));
    if (valid) *valid = t.nframes > 0;
    if (h_prev)
        GOFPID_CUDA_TRY(cudaMemcpy2D(h_prev, (size_t)p->W * p->C, p->fd.prev + (size_t)stream_idx * p->fd.stream_stride,
                                     p->fd.pitch, (size_t)p->W * p->C, p->H, cudaMemcpyDeviceToHost));
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_mog2_mean_modes(gofpid_pipeline *p, double *mean_modes)
{
    GOFPID_GUARD(p);
    if (p->cfg.detector != GOFPID_DETECT_MOG2) { set_error("detector is not MOG2"); return GOFPID_ERR_INVALID; }
    if (!mean_modes) { set_error("null output"); return GOFPID_ERR_INVALID; }
    GOFPID_CUDA_TRY(cudaDeviceSynchronize());
    const size_t nt = (size_t)p->mog2.tiles_x * p->mog2.tiles_y, tpx = (size_t)MOG2_TILE_W * MOG2_TILE_H;
    std::vector<uint8_t> nm(nt * tpx * p->n_streams);
    GOFPID_CUDA_TRY(cudaMemcpy(nm.data(), p->mog2.nmodes, nm.size(), cudaMemcpyDeviceToHost));
    double sum = 0;
    for (int z = 0; z < p->n_streams; z++)
        for (int y = 0; y < p->H; y++)
            for (int x = 0; x < p->W; x++) {
                const size_t t = (size_t)(y / MOG2_TILE_H) * p->mog2.tiles_x + x / MOG2_TILE_W;
                sum += nm[((size_t)z * nt + t) * tpx + (size_t)(y % MOG2_TILE_H) * MOG2_TILE_W + x % MOG2_TILE_W];
            }
    *mean_modes = sum / ((double)p->n_streams * p->H * p->W);
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_set_profiling(gofpid_pipeline *p, int enabled)
{
    GOFPID_GUARD(p);
    p->profiling = enabled != 0;
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_stage_ms(gofpid_pipeline *p, float *h_ms5)
{
    GOFPID_GUARD(p);
    if (!h_ms5) { set_error("null output"); return GOFPID_ERR_INVALID; }
    if (!p->profiling || p->frames == 0) { set_error("profiling is off or no frame processed"); return GOFPID_ERR_INVALID; }
    GOFPID_CUDA_TRY(cudaEventSynchronize(p->ev[3]));
    float blobs_ms = 0.f;
    GOFPID_CUDA_TRY(cudaEventElapsedTime(&h_ms5[0], p->ev[4], p->ev[1]));
    GOFPID_CUDA_TRY(cudaEventElapsedTime(&h_ms5[1], p->ev[1], p->ev[2]));
    GOFPID_CUDA_TRY(cudaEventElapsedTime(&blobs_ms, p->ev[2], p->ev[3]));
    h_ms5[2] = blobs_ms;
    GOFPID_CUDA_TRY(cudaEventElapsedTime(&h_ms5[3], p->ev[0], p->ev[4]));
    GOFPID_CUDA_TRY(cudaEventElapsedTime(&h_ms5[4], p->ev[0], p->ev[3]));
    return GOFPID_OK;
}

GOFPID_API int gofpid_pipeline_launch_count(gofpid_pipeline *p)
{
    return p ? p->launch_count : 0;
}

// =====================================================================================
// stage-level entry points
// =====================================================================================
GOFPID_API int gofpid_blur_u8(const uint8_t *d_src, size_t src_pitch, uint8_t *d_dst, size_t dst_pitch,
                              int H, int W, int C, int kind, int kw, int kh, double sigma, void *stream)
{
    if (!d_src || !d_dst || H <= 0 || W <= 0) { set_error("bad argument"); return GOFPID_ERR_INVALID; }
    if (gofpid_device_count() == 0) { set_error("no CUDA device available (no CPU fallback)"); return GOFPID_ERR_CUDA; }
    DetectLaunch L;
    L.H = H; L.W = W; L.C = C; L.n_streams = 1;
    L.src = FrameSrc{d_src, src_pitch, 0};
    GOFPID_TRY(make_blur_params(kind, kw, kh, sigma, &L.blur));
    L.mask = ImageU8{nullptr, 0, 0};
    L.tick = nullptr;
    L.stream = (cudaStream_t)stream;
    return launch_blur_only(L, ImageU8{d_dst, dst_pitch, 0});
}

GOFPID_API int gofpid_morph_u8(const uint8_t *d_src, size_t src_pitch, uint8_t *d_dst, size_t dst_pitch,
                               int H, int W, const gofpid_morph_op *ops, int n_ops, void *stream)
{
    return gofpid_morph_levels_u8(d_src, src_pitch, d_dst, dst_pitch, H, W, ops, n_ops, nullptr, 0, stream);
}

GOFPID_API int gofpid_morph_levels_u8(const uint8_t *d_src, size_t src_pitch, uint8_t *d_dst, size_t dst_pitch,
                                      int H, int W, const gofpid_morph_op *ops, int n_ops,
                                      const int *levels, int n_levels, void *stream)
{
    if (!d_src || !d_dst || H <= 0 || W <= 0 || (n_ops > 0 && !ops)) { set_error("bad argument"); return GOFPID_ERR_INVALID; }
    if (gofpid_device_count() == 0) { set_error("no CUDA device available (no CPU fallback)"); return GOFPID_ERR_CUDA; }
    MorphPlan plan;
    GOFPID_TRY(morph_plan_create(ops, n_ops, &plan));
    if (levels && n_levels > 0) {
        bool ok = n_levels <= 3;
        for (int i = 0; ok && i < n_levels; i++)
            ok = levels[i] >= 1 && levels[i] <= 255 && (i == 0 || levels[i] > levels[i - 1]);
        if (!ok) {
            morph_plan_destroy(&plan);
            set_error("levels must be 1..3 ascending values in [1, 255]");
            return GOFPID_ERR_INVALID;
        }
        // (a chain the bit-plane kernel declines -- see morph_plan_set_levels -- stays on the byte kernel)
        morph_plan_set_levels(&plan, levels, n_levels);
    }
    int rc = launch_morph(plan, ImageU8{const_cast<uint8_t *>(d_src), src_pitch, 0}, ImageU8{d_dst, dst_pitch, 0},
                          H, W, 1, (cudaStream_t)stream);
    if (rc == GOFPID_OK && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
        set_error("morphology kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = GOFPID_ERR_CUDA;
    }
    morph_plan_destroy(&plan);
    return rc;
}

GOFPID_API int gofpid_ccl_u8(const uint8_t *d_mask, size_t mask_pitch, int32_t *d_labels, size_t labels_pitch,
                             int H, int W, void *stream)
{
    if (!d_mask || !d_labels || H <= 0 || W <= 0 || labels_pitch % sizeof(int32_t)) { set_error("bad argument"); return GOFPID_ERR_INVALID; }
    if (gofpid_device_count() == 0) { set_error("no CUDA device available (no CPU fallback)"); return GOFPID_ERR_CUDA; }
    return launch_ccl_only(d_mask, mask_pitch, d_labels, labels_pitch / sizeof(int32_t), H, W, (cudaStream_t)stream);
}

} // extern "C"
