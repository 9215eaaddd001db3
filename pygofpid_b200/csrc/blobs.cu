// blobs.cu -- foreground blob creation on the device (sm_100a).
//
// Replaces GOFPID._find_blob / _create_blob (pipeline.py:555-588) and helpers.find_contours
// (helpers.py:125-150), i.e. cv.findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) + cv.contourArea
// + the sums behind cv.moments:
//
//   1. label_tiles_kernel   union-find in shared memory over 128x16 tiles.  Foreground pixels
//                           (mask != 0) are joined 8-connected, background pixels 4-connected,
//                           in the same pass; background touching the frame border is joined
//                           to a virtual OUTSIDE node (value 0).  Writes provisional labels.
//   2. merge_tiles_kernel   unions across tile borders (global atomicMin union-find).
//   3. flatten_kernel       path-compresses every pixel to its root, and gives every
//                           foreground root a component slot.
//   4. stats_kernel         per-component area / bounding box with warp-aggregated atomics
//                           (match.any + redux, one atomic set per warp and component).
//   5. trace_kernel         for each EXTERNAL component (the background west of its first
//                           pixel reaches OUTSIDE) whose box can hold area_min: stages the
//                           component as a bitmap in shared memory and follows its outer
//                           border exactly as OpenCV's tracer does, accumulating the polygon
//                           area and first moments, then emits the contour if
//                           contourArea >= area_min.
//
// A label value is 1 + (y * LP + x) of the component's raster-first pixel (LP = padded row
// pitch of the label image); pipeline.cu converts to the canonical dense numbering.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

namespace gofpid {

constexpr int CCL_TW = 64;  // one 64-bit word of class bits per tile row
constexpr int CCL_TH = 32;
constexpr int CCL_THREADS = 16 * CCL_TH; // each thread owns 4 consecutive pixels of a row

// (a 256-thread / 48 KB variant that fits next to resident MOG2 tiles was measured and is no
// faster: the CTAs are few, and each phase then takes more rounds)
#ifndef GOFPID_TRACE_THREADS
#define GOFPID_TRACE_THREADS 1024
#endif
#ifndef GOFPID_TRACE_LOG2_CP
#define GOFPID_TRACE_LOG2_CP 12
#endif
// (Measured and dropped in round 2: blob kernels sized to fit beside resident detector CTAs -- a 128-thread
// labeller with a 16-bit forest in 26 KB, a 128-thread / 512-checkpoint instance of this tracer, and a
// one-warp-per-component sequential tracer.  All bit-exact, all slower for the whole step than the SM-sized
// kernels below: 25.5 / 25.5 / 91 us per 1080p frame against 22.9, see DESIGN.md.)
constexpr int TRACE_THREADS = GOFPID_TRACE_THREADS, TRACE_LOG2_CP = GOFPID_TRACE_LOG2_CP;
// per checkpoint: a00 (8 B) + state, length, two (next, distance) pairs (24 B); hash: 2 slots of 8 B per checkpoint
constexpr size_t trace_smem_bytes(int log2_cp) { return ((size_t)1 << log2_cp) * (8 + 24) + ((size_t)2 << log2_cp) * 8; }

// component record layout (8 int32)
enum { CR_FIRST = 0, CR_X0, CR_Y0, CR_X1, CR_Y1, CR_AREA, CR_RSTART, CR_NRUNS, CR_SIZE }; // (run slice: runs path only)

// ------------------------------------------------------------------------------------
// union-find primitives (value v: 0 = OUTSIDE, else element index v-1)
// ------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int uf_find(const T *lab, int v)
{
    while (v != 0) {
        const int n = (int)lab[v - 1];
        if (n == v) break;
        v = n;
    }
    return v;
}
__device__ __forceinline__ int uf_find_g(const int32_t *lab, int v)
{
    while (v != 0) {
        const int n = *reinterpret_cast<const volatile int32_t *>(lab + (v - 1));
        if (n == v) break;
        v = n;
    }
    return v;
}
__device__ __forceinline__ void uf_union_s(int *lab, int a, int b)
{
    while (true) {
        a = uf_find(lab, a); b = uf_find(lab, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&lab[a - 1], b);
        if (old == a) return;
        a = old;
    }
}
__device__ __forceinline__ void uf_union_g(int32_t *lab, int a, int b)
{
    while (true) {
        a = uf_find_g(lab, a); b = uf_find_g(lab, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&lab[a - 1], b);
        if (old == a) return;
        a = old;
    }
}


// ------------------------------------------------------------------------------------
// 0. run-based labelling on the foreground bit image (the usual path)
// ------------------------------------------------------------------------------------
// After morphology a foreground mask is a few compact blobs: a 1080p frame has some thousand
// horizontal runs.  Two kernels:
//   runs_extract_kernel  many CTAs, one band of RX_ROWS image rows each: run starts and run ends
//       from bit tricks on the 32-pixel words (the k-th start and the k-th end of a band belong
//       to the same run), a block scan, ONE atomicAdd per band to reserve a slice of the
//       stream's run list, and the runs (y, xs, xe) written in raster order inside the band,
//       with every row's slice [rowstart, rowend).
//   runs_ccl_kernel      one CTA per camera stream, everything in shared memory with block
//       barriers between the phases:
//       C  union-find over runs: a run is joined with the runs of the row above that overlap
//          [xs-1, xe+1] (8-connectivity), found by binary search in that row's slice;
//       D  flatten; roots get component slots (block scan);
//       E  area / bounding box / raster-first pixel per component with shared-memory atomics;
//       F  component records, and the label image filled along the runs only (the tracer asks
//          which component a border pixel belongs to).
// A frame with more runs or components than fit raises the per-stream fallback flag
// (info.reserved[1]) and is labelled by the tiled union-find kernels below instead.
constexpr int RX_ROWS = 8;      // image rows per extraction CTA
constexpr int RX_THREADS = 256;
constexpr int RX_MAXW = 8;      // words per thread at most (bands up to 2048 words)
constexpr int RC_THREADS = 1024;
constexpr int RC_MAX_COMP = 1024;                 // component slots in shared memory
constexpr size_t RC_SMEM_BUDGET = 224 * 1024;     // bytes of dynamic shared memory

// run capacity per stream for frames of H rows (0: the shape does not fit, always fall back)
int run_ccl_capacity(int H, int bit_pitch)
{
    if (RX_ROWS * bit_pitch > RX_THREADS * RX_MAXW) return 0;
    const size_t fixed = (size_t)2 * ((H + 3) & ~3) * 4 + (size_t)RC_MAX_COMP * 8 * 4;
    if (fixed + 4096 * 10 > RC_SMEM_BUDGET) return 0;
    int n = (int)((RC_SMEM_BUDGET - fixed) / 10) & ~7;
    return n > 65528 ? 65528 : n;
}
static size_t run_ccl_smem(int H, int max_runs)
{
    return (size_t)2 * ((H + 3) & ~3) * 4 + (size_t)RC_MAX_COMP * 8 * 4 + (size_t)max_runs * 10;
}

__global__ void __launch_bounds__(RX_THREADS)
runs_extract_kernel(const uint32_t *bitimg, int bit_pitch, size_t bit_stream, RunList R, int H)
{
    __shared__ int sh_scan[RX_THREADS / 32];
    __shared__ int sh_base;
    const int z = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int y0 = blockIdx.x * RX_ROWS;
    const int rows = min(RX_ROWS, H - y0);
    const int nw = rows * bit_pitch;                  // physical words of the band (guards included)
    const int k = (nw + RX_THREADS - 1) / RX_THREADS; // consecutive words per thread (<= RX_MAXW)
    const uint32_t *base = bitimg + (size_t)z * bit_stream + (size_t)(y0 + 1) * bit_pitch;
    const int j0 = tid * k;
    // guard words are zero, so a run never continues across a row end: no boundary tests
    uint32_t wv[RX_MAXW + 2];
#pragma unroll
    for (int i = 0; i < RX_MAXW + 2; i++) {
        const int j = j0 - 1 + i;
        wv[i] = (i < k + 2 && j >= 0 && j < nw) ? __ldg(base + j) : 0u;
    }
    uint32_t st[RX_MAXW], en[RX_MAXW];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < RX_MAXW; i++) {
        const uint32_t cur = i < k ? wv[i + 1] : 0u; // (slot k + 1 holds the next thread's first word)
        st[i] = cur & ~((cur << 1) | (wv[i] >> 31));
        en[i] = cur & ~((cur >> 1) | (wv[i + 2] << 31));
        cnt += __popc(st[i]) | (__popc(en[i]) << 16);
    }
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh_scan[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int v = lane < RX_THREADS / 32 ? sh_scan[lane] : 0;
#pragma unroll
        for (int o = 1; o < RX_THREADS / 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane < RX_THREADS / 32) sh_scan[lane] = v;
        if (lane == RX_THREADS / 32 - 1) sh_base = atomicAdd(R.count + z, v & 0xffff);
    }
    __syncthreads();
    const int total = sh_scan[RX_THREADS / 32 - 1] & 0xffff;
    const int gbase = sh_base;
    const bool fits = gbase + total <= R.max_runs;
    const int ex = (warp ? sh_scan[warp - 1] : 0) + inc - cnt;
    int ks = gbase + (ex & 0xffff), ke = gbase + (ex >> 16);
    uint16_t *xs = R.xs + (size_t)z * R.max_runs, *xe = R.xe + (size_t)z * R.max_runs,
             *ry = R.y + (size_t)z * R.max_runs;
    int *rowstart = R.rowstart + (size_t)z * R.row_pitch, *rowend = R.rowend + (size_t)z * R.row_pitch;
    int y = j0 / bit_pitch, c = j0 - y * bit_pitch; // c: physical word column (logical + 1)
#pragma unroll
    for (int i = 0; i < RX_MAXW; i++) {
        if (i < k && j0 + i < nw) {
            if (c == 0) { // first word of a row (a guard word): the row's slice starts here
                rowstart[y0 + y] = ks;
                if (y > 0) rowend[y0 + y - 1] = ks;
            }
            if (fits) {
                const int x0 = 32 * (c - 1);
                for (uint32_t m = st[i]; m; m &= m - 1) {
                    xs[ks] = (uint16_t)(x0 + __ffs(m) - 1);
                    ry[ks] = (uint16_t)(y0 + y);
                    ks++;
                }
                for (uint32_t m = en[i]; m; m &= m - 1) xe[ke++] = (uint16_t)(x0 + __ffs(m) - 1);
            } else {
                ks += __popc(st[i]);
            }
            if (++c == bit_pitch) { c = 0; y++; }
        }
    }
    if (tid == 0) rowend[y0 + rows - 1] = gbase + total;
}

// Union-find over run indices in shared memory.  Edges always lead to a node that stopped being
// a root later than the node they leave (unions link a root to a root; path halving replaces a
// parent by that parent's parent), so the forest stays acyclic under any interleaving; the loops
// are bounded all the same, so that a bug can only produce wrong labels, never a hung device.
constexpr int RC_GUARD = 1 << 17;
__device__ __forceinline__ int rc_find(const int *parent, int k)
{
    for (int it = 0; it < RC_GUARD; it++) {
        const int n = parent[k];
        if (n == k) return k;
        k = n;
    }
    return k;
}
// find with path halving (plain stores: losing a race only loses a shortcut)
__device__ __forceinline__ int rc_find_halve(int *parent, int k)
{
    for (int it = 0; it < RC_GUARD; it++) {
        const int p = parent[k];
        if (p == k) return k;
        const int g = parent[p];
        if (g == p) return p;
        parent[k] = g;
        k = g;
    }
    return k;
}
__device__ __forceinline__ void rc_union(int *parent, int a, int b)
{
    for (int it = 0; it < RC_GUARD; it++) {
        a = rc_find_halve(parent, a); b = rc_find_halve(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

#ifdef GOFPID_TRACE_DEBUG
__device__ long long g_rc_dbg[16];
#define RC_T(k) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 0) g_rc_dbg[k] = clock64(); } while (0)
#else
#define RC_T(k)
#endif

__global__ void __launch_bounds__(RC_THREADS, 1)
runs_ccl_kernel(RunList R, int32_t *labels, size_t LP, size_t labels_stream, int32_t *comp,
                gofpid_frame_info *info, int max_blobs, int max_runs, const FrameTick *tick, int H, int W,
                cudaGraphConditionalHandle cond, int use_cond)
{
    extern __shared__ __align__(16) unsigned char rsm[];
    const int z = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HP = (H + 3) & ~3;
    int *rowstart = reinterpret_cast<int *>(rsm);                               // [HP]
    int *rowend = rowstart + HP;                                                // [HP]
    int *cstat = rowend + HP;                                                   // [8][RC_MAX_COMP]
    int *parent = cstat + 8 * RC_MAX_COMP;                                      // [max_runs]
    uint16_t *run_xs = reinterpret_cast<uint16_t *>(parent + max_runs);         // [max_runs]
    uint16_t *run_xe = run_xs + max_runs;
    uint16_t *run_y = run_xe + max_runs;
    __shared__ int sh_scan[32];

    gofpid_frame_info *fi = info + z;
    RC_T(0);
    const int N = R.count[z];
    const bool fallback = max_runs == 0 || N > max_runs;
    __syncthreads(); // every thread has read the count
    if (tid == 0) {
        R.count[z] = 0; // for the slot's next frame
        R.n_labelled[z] = fallback ? 0 : N;
        fi->n_components = 0; fi->n_candidates = 0; fi->n_kept = 0; fi->n_points = 0;
        fi->overflow = 0; fi->frame_index = tick ? tick->nframes : 0;
        fi->reserved[0] = 0; fi->reserved[1] = fallback ? 1 : 0;
        // inside a captured step the tiled fallback kernels sit in a conditional graph node
        if (fallback && use_cond) cudaGraphSetConditional(cond, 1);
    }
    if (fallback) return;

    // ---- runs and row slices into shared memory ---------------------------------------------------
    {
        const uint16_t *gxs = R.xs + (size_t)z * R.max_runs, *gxe = R.xe + (size_t)z * R.max_runs,
                       *gy = R.y + (size_t)z * R.max_runs;
        // (max_runs is a multiple of 8: the three lists are 16-byte aligned, copied 8 runs at a time)
        for (int i = tid; i < (N + 7) / 8; i += RC_THREADS) {
            reinterpret_cast<uint4 *>(run_xs)[i] = __ldg(reinterpret_cast<const uint4 *>(gxs) + i);
            reinterpret_cast<uint4 *>(run_xe)[i] = __ldg(reinterpret_cast<const uint4 *>(gxe) + i);
            reinterpret_cast<uint4 *>(run_y)[i] = __ldg(reinterpret_cast<const uint4 *>(gy) + i);
        }
        const int *grs = R.rowstart + (size_t)z * R.row_pitch, *gre = R.rowend + (size_t)z * R.row_pitch;
        for (int y = tid; y < H; y += RC_THREADS) { rowstart[y] = __ldg(grs + y); rowend[y] = __ldg(gre + y); }
    }
    __syncthreads();
    RC_T(3);

    // ---- C: unions with the row above -----------------------------------------------------------
    // (1) every run points at the FIRST run of the row above that touches it: one plain store per
    // run, a forest whose edges all go up one image row; (2) pointer jumping flattens the vertical
    // chains in log(height) rounds; (3) the remaining overlaps -- where branches of a blob meet --
    // are merged root to root with atomicMin.
    auto overlap_range = [&](int k, int &lo, int &end, int &b) {
        const int y = run_y[k];
        lo = end = 0; b = 0;
        if (y == 0) return;
        const int a = (int)run_xs[k] - 1;
        b = (int)run_xe[k] + 1;
        lo = rowstart[y - 1];
        int hi = rowend[y - 1];
        end = hi;
        while (lo < hi) { // first run of the row above with xe >= a
            const int mid = (lo + hi) >> 1;
            if ((int)run_xe[mid] < a) lo = mid + 1; else hi = mid;
        }
        if (lo < end && (int)run_xs[lo] > b) end = lo; // nothing touches
    };
    for (int k = tid; k < N; k += RC_THREADS) {
        int lo, end, b;
        overlap_range(k, lo, end, b);
        parent[k] = lo < end ? lo : k;
    }
    __syncthreads();
    for (int round = 0; round < 16; round++) {
        bool changed = false;
        for (int k = tid; k < N; k += RC_THREADS) {
            const int p = parent[k], g = parent[p];
            if (g != p) { parent[k] = g; changed = true; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    for (int k = tid; k < N; k += RC_THREADS) {
        int lo, end, b;
        overlap_range(k, lo, end, b);
        for (int j = lo + 1; j < end && (int)run_xs[j] <= b; j++) rc_union(parent, k, j);
    }
    __syncthreads();
    RC_T(4);

    // ---- D: flatten; roots get slots ------------------------------------------------------------
    for (int k = tid; k < N; k += RC_THREADS) parent[k] = rc_find(parent, k);
    __syncthreads();
    const int per = (N + RC_THREADS - 1) / RC_THREADS;
    const int k0 = min(tid * per, N), k1 = min(k0 + per, N);
    int mine = 0;
    for (int k = k0; k < k1; k++) mine += parent[k] == k;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sh_scan[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = sh_scan[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        sh_scan[lane] = v;
    }
    __syncthreads();
    const int ncomp = sh_scan[31];
    if (ncomp > RC_MAX_COMP && ncomp <= max_blobs) { // more components than slots here: tiled path
        if (tid == 0) {
            fi->reserved[1] = 1;
            if (use_cond) cudaGraphSetConditional(cond, 1);
        }
        return;
    }
    const int nslot = min(ncomp, min(max_blobs, RC_MAX_COMP));
    int *c_first = cstat, *c_x0 = cstat + RC_MAX_COMP, *c_y0 = cstat + 2 * RC_MAX_COMP,
        *c_x1 = cstat + 3 * RC_MAX_COMP, *c_y1 = cstat + 4 * RC_MAX_COMP, *c_area = cstat + 5 * RC_MAX_COMP,
        *c_nruns = cstat + 6 * RC_MAX_COMP, *c_rstart = cstat + 7 * RC_MAX_COMP;
    {
        int slot = (warp ? sh_scan[warp - 1] : 0) + incl - mine;
        // roots are marked -(slot + 1); non-root runs keep pointing at their (flattened) root
        for (int k = k0; k < k1; k++)
            if (parent[k] == k) {
                if (slot < nslot) {
                    c_first[slot] = 0x7fffffff;
                    c_x0[slot] = 0x7fffffff; c_y0[slot] = 0x7fffffff; c_x1[slot] = -1; c_y1[slot] = -1;
                    c_area[slot] = 0; c_nruns[slot] = 0;
                }
                parent[k] = -(slot + 1);
                slot++;
            }
    }
    __syncthreads();
    RC_T(5);

    // ---- E: component statistics; every run learns its slot -------------------------------------
    // (bands land in the run list in any order, so the smallest index is not the raster-first run:
    // the first pixel is the minimum of y * LP + xs over the component)
    for (int kb = 0; kb < N; kb += RC_THREADS) {
        const int k = kb + tid;
        int slot = -1, xs = 0, xe = 0, y = 0;
        if (k < N) {
            int p = parent[k];
            if (p >= 0) p = parent[p];
            slot = -p - 1;
            if (slot >= nslot) slot = -1; // beyond the handle's capacity (overflow is flagged below)
            xs = run_xs[k]; xe = run_xe[k]; y = run_y[k];
        }
        // neighbouring runs mostly belong to a handful of components: one atomic set per warp and component
        const unsigned grp = __match_any_sync(0xffffffffu, slot);
        if (slot >= 0) {
            const int first = __reduce_min_sync(grp, (int)((size_t)y * LP + xs));
            const int area = __reduce_add_sync(grp, xe - xs + 1);
            const int mnx = __reduce_min_sync(grp, xs), mxx = __reduce_max_sync(grp, xe);
            const int mny = __reduce_min_sync(grp, y), mxy = __reduce_max_sync(grp, y);
            if (lane == __ffs(grp) - 1) {
                atomicMin(&c_first[slot], first);
                atomicAdd(&c_area[slot], area);
                atomicAdd(&c_nruns[slot], __popc(grp));
                atomicMin(&c_x0[slot], mnx); atomicMax(&c_x1[slot], mxx);
                atomicMin(&c_y0[slot], mny); atomicMax(&c_y1[slot], mxy);
            }
        }
    }
    __syncthreads();
    RC_T(6);

    // ---- F: the runs grouped by component (what the tracer walks), records, per-run labels ----------
    {
        // exclusive scan of the run counts over the slots (RC_MAX_COMP == RC_THREADS: one slot per thread)
        static_assert(RC_MAX_COMP == RC_THREADS, "one slot per thread");
        const int mineF = tid < nslot ? c_nruns[tid] : 0;
        int inc = mineF;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        __syncthreads(); // sh_scan is free again
        if (lane == 31) sh_scan[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int v = sh_scan[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            sh_scan[lane] = v;
        }
        __syncthreads();
        const int rstart = (warp ? sh_scan[warp - 1] : 0) + inc - mineF;
        if (tid < nslot) {
            c_rstart[tid] = rstart;
            int32_t *cr = comp + ((size_t)z * max_blobs + tid) * CR_SIZE;
            cr[CR_FIRST] = c_first[tid]; cr[CR_X0] = c_x0[tid]; cr[CR_Y0] = c_y0[tid]; cr[CR_X1] = c_x1[tid];
            cr[CR_Y1] = c_y1[tid]; cr[CR_AREA] = c_area[tid]; cr[CR_RSTART] = rstart; cr[CR_NRUNS] = mineF;
            c_nruns[tid] = 0; // becomes the scatter cursor
        }
        if (tid == 0) {
            fi->n_components = ncomp;
            if (ncomp > max_blobs) fi->overflow = 1;
        }
    }
    __syncthreads();
    {
        uint2 *byc = R.byc + (size_t)z * R.max_runs;
        int32_t *rlabel = R.label + (size_t)z * R.max_runs;
        for (int k = tid; k < N; k += RC_THREADS) {
            int p = parent[k];
            if (p >= 0) p = parent[p];
            const int slot = -p - 1;
            int v = 0;
            if (slot < nslot) {
                v = c_first[slot] + 1;
                const int pos = c_rstart[slot] + atomicAdd(&c_nruns[slot], 1);
                byc[pos] = make_uint2(((uint32_t)run_y[k] << 16) | run_xs[k], run_xe[k]);
            }
            rlabel[k] = v;
        }
    }
    RC_T(7);
#ifdef GOFPID_TRACE_DEBUG
    if (threadIdx.x == 0 && blockIdx.x == 0) { g_rc_dbg[8] = N; g_rc_dbg[9] = ncomp; }
#endif
}

// foreground bit image of a u8 mask (frames whose morphology step did not already produce it)
__global__ void __launch_bounds__(256)
pack_bits_kernel(ImageU8 mask, uint32_t *bitimg, int bit_pitch, size_t bit_stream, int H, int W)
{
    const int x = (blockIdx.x * 32 + threadIdx.x) * 4, y = blockIdx.y * 8 + threadIdx.y;
    if (y >= H) return; // warp-uniform: a warp is one row of 128 pixels
    const uint8_t *mrow = mask.base + (size_t)blockIdx.z * mask.stream_stride + (size_t)y * mask.pitch;
    uint32_t mw = 0;
    if (x < W) mw = *reinterpret_cast<const uint32_t *>(mrow + x);
    unsigned nib = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (x + j < W && ((mw >> (8 * j)) & 255) != 0) nib |= 1u << j;
    const unsigned lane = threadIdx.x;
    const unsigned word = __reduce_or_sync(0xffu << (lane & 24), nib << (4 * (lane & 7)));
    if ((lane & 7) == 0 && x < W)
        bitimg[(size_t)blockIdx.z * bit_stream + (size_t)(y + 1) * bit_pitch + 1 + (x >> 5)] = word;
}

// label image from the run list (not on the hot path: gofpid_pipeline_labels / gofpid_ccl_u8 only).
// One warp per run; streams labelled by the tiled fallback already have their label image.
__global__ void __launch_bounds__(256)
fill_labels_kernel(RunList R, int z, const gofpid_frame_info *info, int32_t *labels, size_t LP)
{
    if (info->reserved[1]) return;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= R.n_labelled[z]) return;
    const int v = R.label[(size_t)z * R.max_runs + k];
    if (v <= 0) return; // component beyond the handle's capacity
    const int y = R.y[(size_t)z * R.max_runs + k], xs = R.xs[(size_t)z * R.max_runs + k],
              xe = R.xe[(size_t)z * R.max_runs + k];
    for (int x = xs + lane; x <= xe; x += 32) labels[(size_t)y * LP + x] = v;
}

// ------------------------------------------------------------------------------------
// 1. tile-local labelling (run based)
// ------------------------------------------------------------------------------------
// Every tile row is one 64-bit word of foreground bits.  A pixel's provisional label is the
// start of its horizontal run (found with clz on the row word), so no horizontal unions are
// needed; only the thread that owns a run start joins the run with the runs it touches in the
// row above (8-connected for foreground, 4-connected for background).
__device__ __forceinline__ unsigned long long bits_below(int x) // bits [0, x)
{
    return x >= 64 ? ~0ULL : ((1ULL << x) - 1ULL);
}
// start / end of the maximal run of set bits of `S` that contains bit x
__device__ __forceinline__ int run_start(unsigned long long S, int x)
{
    const unsigned long long brk = ~S & bits_below(x);
    return brk ? 64 - __clzll((long long)brk) : 0;
}
__device__ __forceinline__ int run_end(unsigned long long S, int x)
{
    const unsigned long long brk = ~S & ~bits_below(x + 1);
    return brk ? __ffsll((long long)brk) - 2 : 63;
}

__global__ void __launch_bounds__(CCL_THREADS)
label_tiles_kernel(ImageU8 mask, int32_t *labels, size_t LP, size_t labels_stream, int H, int W,
                   gofpid_frame_info *info, int force)
{
    if (!force && !info[blockIdx.z].reserved[1]) return; // labelled by runs_ccl_kernel
    __shared__ int lab[CCL_TW * CCL_TH];
    __shared__ unsigned int Fw[CCL_TH][2]; // foreground bits of each tile row (two 32-bit halves)
    const int tx = threadIdx.x, ty = threadIdx.y; // blockDim = (16, CCL_TH)
    const int lane = (ty * 16 + tx) & 31;
    const int x0 = blockIdx.x * CCL_TW, y0 = blockIdx.y * CCL_TH;
    const int lx = tx * 4, x = x0 + lx, y = y0 + ty;
    const uint8_t *mrow = mask.base + (size_t)blockIdx.z * mask.stream_stride + (size_t)y * mask.pitch;

    uint32_t mw = 0;
    if (y < H && x < W) mw = *reinterpret_cast<const uint32_t *>(mrow + x); // padded pitch
    unsigned nib = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (y < H && x + j < W && ((mw >> (8 * j)) & 255) != 0) nib |= 1u << j;
    // the 8 lanes that share a 32-pixel half row OR their nibbles together
    const unsigned grp = 0xffu << (lane & 24);
    const unsigned half = __reduce_or_sync(grp, nib << (4 * (tx & 7)));
    if ((tx & 7) == 0) Fw[ty][tx >> 3] = half;
    __syncthreads();

    // most tiles of a foreground mask are empty: nothing to label, nothing to write
    if (!__syncthreads_or(nib != 0)) return;

    const bool row_ok = y < H;
    const unsigned long long F = ((unsigned long long)Fw[ty][1] << 32) | Fw[ty][0];
    const unsigned long long Fp = ty > 0 ? (((unsigned long long)Fw[ty - 1][1] << 32) | Fw[ty - 1][0]) : 0ULL;

    // provisional labels: every foreground pixel points at the start of its horizontal run.  The
    // four pixels of a thread mostly share a run, so the 64-bit run searches are done once per
    // run, not once per pixel.
    int ra[4]; // run start column of each foreground pixel (-1: background / outside)
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int c = lx + j;
        const int li = ty * CCL_TW + c;
        if (!((nib >> j) & 1u)) { ra[j] = -1; lab[li] = li + 1; continue; }
        const int a = (j > 0 && ra[j - 1] >= 0) ? ra[j - 1] : run_start(F, c);
        ra[j] = a;
        lab[li] = ty * CCL_TW + a + 1;
    }
    __syncthreads();

    // vertical unions (8-connectivity), one thread per run
    if (row_ok && ty > 0 && Fp) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = lx + j;
            if (ra[j] != c) continue; // not a run start
            const int b = run_end(F, c);
            unsigned long long m = Fp & bits_below(min(b + 1, 63) + 1) & ~bits_below(max(c - 1, 0));
            const int me = ty * CCL_TW + c + 1;
            while (m) {
                const int pos = __ffsll((long long)m) - 1;
                const int ap = run_start(Fp, pos), bp = run_end(Fp, pos);
                uf_union_s(lab, me, (ty - 1) * CCL_TW + ap + 1);
                m &= ~bits_below(bp + 1);
            }
        }
    }
    __syncthreads();

    // compress: every run start learns its root (read phase, barrier, write phase), so the
    // per-pixel lookups below are a single hop
    int rr[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        rr[j] = -1;
        if (ra[j] == lx + j) rr[j] = uf_find(lab, ty * CCL_TW + lx + j + 1);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (rr[j] >= 0) lab[ty * CCL_TW + lx + j] = rr[j];
    __syncthreads();

    if (!nib) return; // background entries of the label image are never read
    int out[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (ra[j] < 0) { out[j] = 0; continue; }
        if (j > 0 && ra[j] == ra[j - 1]) { out[j] = out[j - 1]; continue; }
        const int r = lab[ty * CCL_TW + ra[j]]; // root of the run start
        out[j] = (int)((size_t)(y0 + (r - 1) / CCL_TW) * LP + x0 + (r - 1) % CCL_TW) + 1;
    }
    int32_t *lrow = labels + (size_t)blockIdx.z * labels_stream + (size_t)y * LP;
    *reinterpret_cast<int4 *>(lrow + x) = make_int4(out[0], out[1], out[2], out[3]);
}

// ------------------------------------------------------------------------------------
// 2. unions across tile borders
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CCL_TW + 2 * CCL_TH)
merge_tiles_kernel(ImageU8 mask, int32_t *labels, size_t LP, size_t labels_stream, int H, int W,
                   const gofpid_frame_info *info, int force)
{
    if (!force && !info[blockIdx.z].reserved[1]) return;
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * CCL_TW, y0 = blockIdx.y * CCL_TH;
    int x, y;
    if (t < CCL_TW) { x = x0 + t; y = y0; }                         // top row
    else if (t < CCL_TW + CCL_TH) { x = x0; y = y0 + (t - CCL_TW); } // left column
    else { x = x0 + CCL_TW - 1; y = y0 + (t - CCL_TW - CCL_TH); }    // right column
    if (t >= CCL_TW && y == y0) return;                              // corners belong to the top row
    if (x >= W || y >= H) return;
    const uint8_t *m = mask.base + (size_t)blockIdx.z * mask.stream_stride;
    int32_t *lab = labels + (size_t)blockIdx.z * labels_stream;
    auto cl = [&](int yy, int xx) -> int {
        if (yy < 0 || xx < 0 || xx >= W) return 2;
        return m[(size_t)yy * mask.pitch + xx] != 0 ? 1 : 0;
    };
    if (cl(y, x) != 1) return; // only foreground is labelled
    const int v = (int)((size_t)y * LP + x) + 1;
    const bool top = y == y0, left = x == x0, right = x == x0 + CCL_TW - 1;
    const int cw = cl(y, x - 1), cn = cl(y - 1, x);
    if (left && cw == 1) uf_union_g(lab, v, v - 1);
    if (top && cn == 1) uf_union_g(lab, v, v - (int)LP);
    if (cn != 1) {
        if ((top || left) && cw != 1 && cl(y - 1, x - 1) == 1) uf_union_g(lab, v, v - (int)LP - 1);
        if ((top || right) && cl(y - 1, x + 1) == 1) uf_union_g(lab, v, v - (int)LP + 1);
    }
}

// ------------------------------------------------------------------------------------
// 3. flatten + component slots
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
flatten_kernel(ImageU8 mask, int32_t *labels, size_t LP, size_t labels_stream, int32_t *slot,
               int32_t *comp, gofpid_frame_info *info, int max_blobs, uint32_t *bitimg,
               int bit_pitch, size_t bit_stream, int H, int W)
{
    if (!info[blockIdx.z].reserved[1]) return; // labelled by runs_ccl_kernel
    const int x = (blockIdx.x * 32 + threadIdx.x) * 4, y = blockIdx.y * 8 + threadIdx.y;
    if (y >= H) return; // warp-uniform: a warp is one row of 128 pixels
    int32_t *lab = labels + (size_t)blockIdx.z * labels_stream;
    const uint8_t *mrow = mask.base + (size_t)blockIdx.z * mask.stream_stride + (size_t)y * mask.pitch;
    uint32_t mw = 0;
    int v[4] = {0, 0, 0, 0};
    if (x < W) mw = *reinterpret_cast<const uint32_t *>(mrow + x);
    if (mw) { // label entries exist for foreground pixels only
        const int4 v4 = *reinterpret_cast<const int4 *>(lab + (size_t)y * LP + x);
        v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
    }
    unsigned nib = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (x + j >= W || ((mw >> (8 * j)) & 255) == 0) { v[j] = 0; continue; }
        nib |= 1u << j;
        const int self = (int)((size_t)y * LP + x + j) + 1;
        const int r = uf_find_g(lab, v[j]);
        v[j] = r;
        if (r == self) {
            gofpid_frame_info *fi = info + blockIdx.z;
            const int s = atomicAdd(&fi->n_components, 1);
            int32_t *sl = slot + (size_t)blockIdx.z * LP * H;
            if (s < max_blobs) {
                sl[self - 1] = s;
                int32_t *cr = comp + ((size_t)blockIdx.z * max_blobs + s) * CR_SIZE;
                cr[CR_FIRST] = self - 1; cr[CR_X0] = 0x7fffffff; cr[CR_Y0] = 0x7fffffff;
                cr[CR_X1] = -1; cr[CR_Y1] = -1; cr[CR_AREA] = 0;
            } else {
                sl[self - 1] = -1;
                atomicOr(&fi->overflow, 1);
            }
        }
    }
    // foreground bit image for the border tracer: pixel (x, y) is bit x + 32 of row y + 1
    // (one guard word left / right and one guard row above / below stay zero)
    const unsigned lane = threadIdx.x;
    const unsigned word = __reduce_or_sync(0xffu << (lane & 24), nib << (4 * (lane & 7)));
    if ((lane & 7) == 0 && x < W)
        bitimg[(size_t)blockIdx.z * bit_stream + (size_t)(y + 1) * bit_pitch + 1 + (x >> 5)] = word;
    // roots are final here (merging finished with the previous kernel), so writing the root
    // over a non-root entry only shortens paths other threads may still be walking
    if (nib) *reinterpret_cast<int4 *>(lab + (size_t)y * LP + x) = make_int4(v[0], v[1], v[2], v[3]);
}

// ------------------------------------------------------------------------------------
// 4. component statistics
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stats_kernel(ImageU8 mask, const int32_t *labels, size_t LP, size_t labels_stream,
             const int32_t *slot, int32_t *comp, const gofpid_frame_info *info, int max_blobs, int H, int W)
{
    if (!info[blockIdx.z].reserved[1]) return; // labelled by runs_ccl_kernel
    const int x = (blockIdx.x * 32 + threadIdx.x) * 4, y = blockIdx.y * 8 + threadIdx.y;
    const bool inside = y < H && x < W;
    const int32_t *lab = labels + (size_t)blockIdx.z * labels_stream;
    const int32_t *sl = slot + (size_t)blockIdx.z * LP * H;
    int32_t *cbase = comp + (size_t)blockIdx.z * max_blobs * CR_SIZE;
    uint32_t mw = 0;
    int v[4] = {0, 0, 0, 0};
    if (inside) {
        mw = *reinterpret_cast<const uint32_t *>(mask.base + (size_t)blockIdx.z * mask.stream_stride +
                                                 (size_t)y * mask.pitch + x);
        if (mw) {
            const int4 v4 = *reinterpret_cast<const int4 *>(lab + (size_t)y * LP + x);
            v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
        }
    }
    // runs of equal component inside the thread's four pixels, flushed warp-aggregated
    int cur = -1, cnt = 0, rx0 = 0, rx1 = 0;
#pragma unroll
    for (int j = 0; j <= 4; j++) {
        int s = -1;
        if (j < 4 && inside && x + j < W && ((mw >> (8 * j)) & 255) != 0) {
            // the slot lookup only changes when the label does
            s = (j > 0 && v[j] == v[j - 1] && cur >= 0) ? cur : sl[v[j] - 1];
        }
        const bool flush = cur >= 0 && s != cur;
        const unsigned fm = __ballot_sync(0xffffffffu, flush);
        if (flush) {
            const unsigned grp = __match_any_sync(fm, cur);
            const int tot = __reduce_add_sync(grp, cnt);
            const int mn = __reduce_min_sync(grp, rx0);
            const int mx = __reduce_max_sync(grp, rx1);
            if ((int)(__ffs(grp) - 1) == (int)threadIdx.x) {
                int32_t *cr = cbase + (size_t)cur * CR_SIZE;
                atomicAdd(&cr[CR_AREA], tot);
                atomicMin(&cr[CR_X0], mn); atomicMax(&cr[CR_X1], mx);
                atomicMin(&cr[CR_Y0], y); atomicMax(&cr[CR_Y1], y);
            }
            cur = -1;
        }
        if (s >= 0) {
            if (cur < 0) { cur = s; cnt = 0; rx0 = x + j; }
            cnt++; rx1 = x + j;
        }
    }
}

// ------------------------------------------------------------------------------------
// 5. outer-border tracing
// ------------------------------------------------------------------------------------
// direction codes: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE (y grows downwards)
// dx + 1 and dy + 1 of the 8 directions, two bits each
#define DIR_DX_LUT 0x901au /* dx+1 of E NE N NW W SW S SE = 2 2 1 0 0 0 1 2 */
#define DIR_DY_LUT 0xa901u /* dy+1 of E NE N NW W SW S SE = 1 0 0 0 1 2 2 2 */
__device__ __forceinline__ int dir_dx(int d) { return (int)((DIR_DX_LUT >> (2 * d)) & 3u) - 1; }
__device__ __forceinline__ int dir_dy(int d) { return (int)((DIR_DY_LUT >> (2 * d)) & 3u) - 1; }

// 3x3 neighbourhood of (x, y) from the foreground bit image as 8 ring bits (bit d = direction d)
struct BitView {
    const uint32_t *img; // stream's bit image
    int pitch;           // words per row
    __device__ __forceinline__ unsigned nb(int x, int y) const
    {
        const int p = x + 31;                       // bit position of x - 1
        const uint32_t *r = img + (size_t)y * pitch + (p >> 5); // row of y - 1
        const int sh = p & 31;
        const unsigned t = __funnelshift_r(__ldg(r), __ldg(r + 1), sh) & 7u;
        const unsigned m = __funnelshift_r(__ldg(r + pitch), __ldg(r + pitch + 1), sh) & 7u;
        const unsigned b = __funnelshift_r(__ldg(r + 2 * pitch), __ldg(r + 2 * pitch + 1), sh) & 7u;
        return ((m >> 2) & 1u) | (((t >> 2) & 1u) << 1) | (((t >> 1) & 1u) << 2) | ((t & 1u) << 3) |
               ((m & 1u) << 4) | ((b & 1u) << 5) | (((b >> 1) & 1u) << 6) | (((b >> 2) & 1u) << 7);
    }
};

#ifdef GOFPID_TRACE_DEBUG
__device__ long long g_trace_dbg[64 * 16];
__device__ int g_trace_dbg_n;
#define DBG_T(k) if (threadIdx.x == 0) dbg_t[k] = clock64()
#else
#define DBG_T(k)
#endif

// ---- parallel tracing with checkpoints ------------------------------------------------
// A visit of the border follower is the state (pixel c, back direction b): the next visit is
// (c + D[s'], s' + 4) with s' the first foreground direction counter-clockwise after b.  On a
// border every visit has b foreground and b+1 background, so visit states can be enumerated
// from the 3x3 neighbourhood alone, and the outer border is the cycle through the start state
// (first pixel, first foreground direction clockwise from west).  Pixels selected by a hash of
// their coordinates become CHECKPOINTS; one thread per checkpoint follows the border to the
// next checkpoint (a short segment), the segments are ranked along the cycle by pointer
// jumping, and then written out in parallel at their offsets.  Tracing reads the frame's
// foreground bit image (L1-resident around a blob), so any blob size is handled alike.
__device__ __forceinline__ bool cp_hash(int x, int y, unsigned rmask)
{
    unsigned h = (unsigned)x * 0x9E3779B1u ^ (unsigned)y * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
    return (h & rmask) == 0;
}
__device__ __forceinline__ int cp_key(int x, int y, int b) { return ((y << 16) | (x << 3) | b) + 1; }

template <int LOG2_HT>
__device__ __forceinline__ int ht_find(const int *keys, const int *vals, int key)
{
    constexpr int HT = 1 << LOG2_HT;
    unsigned h = ((unsigned)key * 0x9E3779B1u) >> (32 - LOG2_HT);
    while (true) {
        const int k = keys[h & (HT - 1)];
        if (k == key) return vals[h & (HT - 1)];
        if (k == 0) return -1;
        h++;
    }
}
template <int LOG2_HT>
__device__ __forceinline__ void ht_insert(int *keys, int *vals, int key, int val)
{
    constexpr int HT = 1 << LOG2_HT;
    unsigned h = ((unsigned)key * 0x9E3779B1u) >> (32 - LOG2_HT);
    while (true) {
        const int old = atomicCAS(&keys[h & (HT - 1)], 0, key);
        if (old == 0 || old == key) { vals[h & (HT - 1)] = val; return; }
        h++;
    }
}

struct SegSums { long long a00, a10, a01; };

// Follows the border from visit (x, y, b) until the next checkpoint visit.  STORE writes the
// visited points, otherwise the Green sums of cv.contourArea / cv.moments are accumulated
// (a step from p to q contributes dxy = p.x*q.y - q.x*p.y, dxy*(p.x+q.x), dxy*(p.y+q.y)).
// Returns the id of the checkpoint reached (-1: none within max_steps).
template <bool STORE, int LOG2_HT>
__device__ int trace_segment(const BitView &V, int x, int y, int b, int xs, int ys, unsigned rmask,
                             const int *keys, const int *vals, int &len, SegSums &S, int32_t *pts,
                             int max_steps)
{
    int n = 0, id = -1;
    long long a00 = 0, a10 = 0, a01 = 0;
    while (n < max_steps) {
        const unsigned nb = V.nb(x, y);
        const unsigned rot = ((nb | (nb << 8)) >> ((b + 1) & 7)) & 0xffu;
        const int s = (b + 1 + (__ffs(rot) - 1)) & 7;
        const int dx = dir_dx(s), dy = dir_dy(s);
        const int dxy = x * dy - y * dx; // = x*(y+dy) - (x+dx)*y
        if (STORE) {
            *reinterpret_cast<int2 *>(pts + 2 * n) = make_int2(x, y);
            a10 += (long long)dxy * (2 * x + dx); a01 += (long long)dxy * (2 * y + dy);
        } else {
            a00 += dxy;
        }
        n++;
        x += dx; y += dy; b = (s + 4) & 7;
        if (cp_hash(x, y, rmask) || (x == xs && y == ys)) {
            id = ht_find<LOG2_HT>(keys, vals, cp_key(x, y, b));
            if (id >= 0) break;
        }
    }
    len = n;
    S.a00 = a00; S.a10 = a10; S.a01 = a01;
    return id;
}

template <int TRACE_THREADS, int LOG2_CP>
__global__ void __launch_bounds__(TRACE_THREADS)
trace_kernel(const int32_t *labels, size_t LP, size_t labels_stream, RunList R, const int32_t *comp,
             gofpid_frame_info *info, gofpid_blob *kept, int32_t *points, int max_blobs,
             int max_kept, int max_points, double area_min, const uint32_t *bitimg, int bit_pitch,
             size_t bit_stream, int H, int W)
{
    constexpr int TRACE_MAX_CP = 1 << LOG2_CP;  // checkpoints per contour
    constexpr int TRACE_HT = 2 * TRACE_MAX_CP;  // hash slots
    constexpr int LOG2_HT = LOG2_CP + 1;
    extern __shared__ __align__(16) unsigned char tsm[];
    long long *seg_a00 = reinterpret_cast<long long *>(tsm);             // per checkpoint
    int *ht_keys = reinterpret_cast<int *>(seg_a00 + TRACE_MAX_CP);
    int *ht_vals = ht_keys + TRACE_HT;
    int *cp_state = ht_vals + TRACE_HT;   // packed (y << 16 | x << 3 | b)
    int *seg_len = cp_state + TRACE_MAX_CP;
    int *nxtA = seg_len + TRACE_MAX_CP;
    int *nxtB = nxtA + TRACE_MAX_CP;
    unsigned *dA = reinterpret_cast<unsigned *>(nxtB + TRACE_MAX_CP);
    unsigned *dB = dA + TRACE_MAX_CP;
    __shared__ int sh_ncp, sh_ncand, sh_keep, sh_off;
    __shared__ long long sh_part[3][TRACE_THREADS / 32];
    const int z = blockIdx.y;
    const int32_t *lab = labels + (size_t)z * labels_stream;
    gofpid_frame_info *fi = info + z;
    const int ncomp = min(fi->n_components, max_blobs);
    // labelled by runs: a component's border pixels come from its own runs; tiled fallback: from
    // the bit image inside its box, filtered with the label image
    const bool by_runs = fi->reserved[1] == 0;
    const int tid = threadIdx.x;
    BitView bv; bv.img = bitimg + (size_t)z * bit_stream; bv.pitch = bit_pitch;

    for (int ci = blockIdx.x; ci < ncomp; ci += gridDim.x) {
        const int32_t *cr = comp + ((size_t)z * max_blobs + ci) * CR_SIZE;
        const int first = cr[CR_FIRST], bx0 = cr[CR_X0], by0 = cr[CR_Y0], bx1 = cr[CR_X1],
                  by1 = cr[CR_Y1];
        const int ys = first / (int)LP, xs = first - ys * (int)LP;
        const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
        // cv.contourArea <= (w-1)(h-1): skip boxes that cannot reach area_min
        if ((double)(bw - 1) * (double)(bh - 1) < area_min) continue;
        if (tid == 0) atomicAdd(&fi->n_candidates, 1);
#ifdef GOFPID_TRACE_DEBUG
        long long dbg_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
        DBG_T(0);
        __syncthreads(); // the previous candidate's tables are no longer in use

        // ---- start state ------------------------------------------------------------------------
        const unsigned nb0 = bv.nb(xs, ys);
        int b0 = -1;
#pragma unroll
        for (int k = 3; k >= -3; k--) { // first foreground neighbour clockwise from west
            const int d = k & 7;
            if (b0 < 0 && ((nb0 >> d) & 1u)) b0 = d;
        }
        if (b0 < 0) { // isolated pixel: a one-point contour of area 0
            if (tid == 0 && 0.0 >= area_min) {
                int keep = atomicAdd(&fi->n_kept, 1);
                const int off = atomicAdd(&fi->n_points, 1);
                if (keep >= max_kept) { atomicOr(&fi->overflow, 2); keep = -1; }
                else if (off + 1 > max_points) { atomicOr(&fi->overflow, 4); keep = -2; }
                if (keep >= 0) {
                    gofpid_blob *kb = kept + (size_t)z * max_kept + keep;
                    kb->first = ys * W + xs; kb->x0 = bx0; kb->y0 = by0; kb->x1 = bx1; kb->y1 = by1;
                    kb->area_px = cr[CR_AREA]; kb->n_points = 1; kb->point_offset = off; kb->nested = 0; kb->reserved = 0;
                    kb->a00 = 0; kb->a10 = 0; kb->a01 = 0;
                    int32_t *dst = points + ((size_t)z * max_points + off) * 2;
                    dst[0] = xs; dst[1] = ys;
                }
            }
            continue;
        }

        // ---- enumerate checkpoints (retry with a sparser hash when there are too many) -----------
        unsigned rmask = 0u; // every visit is a checkpoint unless the border is very long
        {
            const int est = 2 * (bw + bh);
            while ((int)(rmask + 1) * (TRACE_MAX_CP / 2) < est) rmask = (rmask << 1) | 1u;
        }
        const int w0 = (bx0 + 32) >> 5, w1 = (bx1 + 32) >> 5, wpr = w1 - w0 + 1; // bit-image words of the box
        // Balanced enumeration in three phases, each followed by a block-wide scan instead of a
        // contended counter: (A) per bit-image word, the hash-selected border pixels go to a
        // compact candidate list; (B) one thread per candidate counts its visit states;
        // (C) the states are written to their slots and entered in the hash table.
        int *cand_list = nxtA; // the ranking arrays are free until pass 1: 4 * TRACE_MAX_CP ints
        constexpr int CAND_CAP = 4 * TRACE_MAX_CP;
        // hash-selected border pixels of bit-image word wc of image row y, columns [xlo, xhi]
        auto cands_at = [&](int y, int wc, int xlo, int xhi) -> uint32_t {
            const int xw = (wc << 5) - 32;
            const uint32_t *r = bv.img + (size_t)y * bv.pitch + wc; // row y - 1
            uint32_t all = 0xffffffffu;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const uint32_t c = __ldg(r + q * bv.pitch), l = __ldg(r + q * bv.pitch - 1),
                               g = __ldg(r + q * bv.pitch + 1);
                all &= c & ((c << 1) | (l >> 31)) & ((c >> 1) | (g << 31));
            }
            // foreground with a background 8-neighbour, inside the columns asked for
            uint32_t cand = __ldg(r + bv.pitch) & ~all;
            if (xlo > xw) cand &= ~0u << min(xlo - xw, 31);
            if (xhi < xw + 31) cand &= ~0u >> (xw + 31 - xhi);
            if (rmask) {
                uint32_t keep = 0;
                for (uint32_t m = cand; m; m &= m - 1) {
                    const int bit = __ffs(m) - 1;
                    if (cp_hash(xw + bit, y, rmask)) keep |= 1u << bit;
                }
                cand = keep;
            }
            return cand;
        };
        // the thread's share of the words that can hold border pixels of the component
        const uint2 *cruns = R.byc + (size_t)z * R.max_runs + cr[CR_RSTART];
        const int nruns = by_runs ? cr[CR_NRUNS] : 0;
        auto for_each_word = [&](auto &&f) {
            if (by_runs) {
                // eight lanes share a run (256 pixels per round), 128 runs in flight per CTA
                for (int i = tid >> 3; i < nruns; i += TRACE_THREADS / 8) {
                    const uint2 rr = __ldg(cruns + i);
                    const int y = (int)(rr.x >> 16), xs_ = (int)(rr.x & 0xffffu), xe_ = (int)rr.y;
                    for (int wc = ((xs_ + 32) >> 5) + (tid & 7); wc <= (xe_ + 32) >> 5; wc += 8) f(y, wc, xs_, xe_);
                }
            } else {
                for (int i = tid; i < wpr * bh; i += TRACE_THREADS) {
                    const int ry = i / wpr;
                    f(by0 + ry, w0 + (i - ry * wpr), bx0, bx1);
                }
            }
        };
        // visit states of a candidate pixel: direction b foreground, b+1 background.  In the box
        // scan, pixels of other components inside the box are dropped: their borders may leave the
        // box, where nothing is a checkpoint, and following them would take arbitrarily long.
        const int root = first + 1;
        auto visit_states = [&](int x, int y) -> unsigned {
            if (!by_runs && lab[(size_t)y * LP + x] != root) return 0u;
            const unsigned nb = bv.nb(x, y);
            unsigned starts = nb & ~((nb >> 1) | (nb << 7)) & 0xffu;
            if (x == xs && y == ys) starts &= ~(1u << b0); // that one is checkpoint 0
            return starts;
        };
        // Two phases, slots handed out by shared-memory counters (the order of candidates and
        // checkpoints is irrelevant: the hash table maps a visit state to its id and the ranking
        // below orders the segments): (A) the hash-selected border pixels go to a candidate list;
        // (B) one thread per candidate enters its visit states in the checkpoint table.
        while (true) {
            for (int i = tid; i < TRACE_HT; i += TRACE_THREADS) ht_keys[i] = 0;
            if (tid == 0) { sh_ncand = 0; sh_ncp = 1; } // checkpoint 0 is the start state
            __syncthreads();
            for_each_word([&](int y, int wc, int xlo, int xhi) {
                const uint32_t c = cands_at(y, wc, xlo, xhi);
                if (c) {
                    int pos = atomicAdd(&sh_ncand, __popc(c));
                    for (uint32_t m = c; m; m &= m - 1, pos++)
                        if (pos < CAND_CAP) cand_list[pos] = (y << 16) | ((wc << 5) - 32 + __ffs(m) - 1);
                }
            });
            __syncthreads();
            const int ncand = sh_ncand;
            bool retry = ncand > CAND_CAP;
            if (!retry) {
                if (tid == 0) {
                    cp_state[0] = (ys << 16) | (xs << 3) | b0;
                    ht_insert<LOG2_HT>(ht_keys, ht_vals, cp_key(xs, ys, b0), 0);
                }
                for (int k = tid; k < ncand; k += TRACE_THREADS) {
                    const int x = cand_list[k] & 0xffff, y = cand_list[k] >> 16;
                    const unsigned st0 = visit_states(x, y);
                    if (!st0) continue;
                    int slot = atomicAdd(&sh_ncp, __popc(st0));
                    for (unsigned st = st0; st; st &= st - 1, slot++) {
                        if (slot >= TRACE_MAX_CP) break;
                        const int bdir = __ffs(st) - 1;
                        cp_state[slot] = (y << 16) | (x << 3) | bdir;
                        ht_insert<LOG2_HT>(ht_keys, ht_vals, cp_key(x, y, bdir), slot);
                    }
                }
                __syncthreads();
                retry = sh_ncp > TRACE_MAX_CP;
            }
            if (!retry) break;
            rmask = (rmask << 1) | 1u; // too many: halve the density and start over
            __syncthreads();
        }
        const int ncp = sh_ncp;
        DBG_T(1);

        // ---- pass 1: every checkpoint follows the border to the next checkpoint -------------------
        const int max_steps = 4 * (bw + 2) * (bh + 2) + 16;
        for (int i = tid; i < ncp; i += TRACE_THREADS) {
            const int st = cp_state[i];
            int len;
            SegSums S;
            const int nx = trace_segment<false, LOG2_HT>(bv, (st >> 3) & 8191, st >> 16, st & 7, xs, ys, rmask,
                                                ht_keys, ht_vals, len, S, nullptr, max_steps);
            seg_a00[i] = S.a00;
            seg_len[i] = len;
            dA[i] = (unsigned)len;
            nxtA[i] = nx == 0 ? -1 : (nx < 0 ? i : nx); // -1 = end of the list (back at the start)
        }
        __syncthreads();
        DBG_T(2);

        // ---- rank the segments along the cycle by pointer jumping -----------------------------------
        {
            int *na = nxtA, *nbuf = nxtB;
            unsigned *da = dA, *db = dB;
            for (int span = 1; span < ncp; span <<= 1) {
                for (int i = tid; i < ncp; i += TRACE_THREADS) {
                    const int n = na[i];
                    if (n >= 0) { db[i] = da[i] + da[n]; nbuf[i] = na[n]; }
                    else { db[i] = da[i]; nbuf[i] = n; }
                }
                __syncthreads();
                int *t = na; na = nbuf; nbuf = t;
                unsigned *u = da; da = db; db = u;
#ifdef GOFPID_TRACE_DEBUG
                if (span == 1) { DBG_T(7); }
#endif
            }
            DBG_T(6);
            // na[i] == -1: segment i is on the outer border; da[i] = points from i to the end
            const unsigned totalN = da[0];
            const bool closed = na[0] == -1;
            long long a00 = 0;
            for (int i = tid; i < ncp; i += TRACE_THREADS) {
                if (closed && na[i] == -1) {
                    a00 += seg_a00[i];
                    nbuf[i] = (int)(totalN - da[i]); // output offset of the segment
                } else nbuf[i] = -1;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a00 += __shfl_xor_sync(0xffffffffu, a00, o);
            if ((tid & 31) == 0) sh_part[0][tid >> 5] = a00;
            __syncthreads();
            DBG_T(3);
            if (tid == 0) {
                long long t00 = 0;
                for (int q = 0; q < TRACE_THREADS / 32; q++) t00 += sh_part[0][q];
                int keep = -1, base = 0;
                if (!closed) atomicOr(&fi->overflow, 8); // border did not close: internal error
                if (closed && fabs((double)t00) * 0.5 >= area_min) {
                    keep = atomicAdd(&fi->n_kept, 1);
                    base = atomicAdd(&fi->n_points, (int)totalN);
                    if (keep >= max_kept) { atomicOr(&fi->overflow, 2); keep = -1; }
                    else {
                        int npts = (int)totalN;
                        if (base + npts > max_points) { atomicOr(&fi->overflow, 4); npts = 0; }
                        gofpid_blob *kb = kept + (size_t)z * max_kept + keep;
                        kb->first = ys * W + xs;
                        kb->x0 = bx0; kb->y0 = by0; kb->x1 = bx1; kb->y1 = by1;
                        kb->area_px = cr[CR_AREA];
                        kb->n_points = npts; kb->point_offset = base; kb->nested = 0; kb->reserved = 0;
                        kb->a00 = t00; kb->a10 = 0; kb->a01 = 0;
                        if (npts == 0) keep = -1;
                    }
                }
                sh_keep = keep; sh_off = base;
            }
            __syncthreads();
            DBG_T(4);

            // ---- pass 2: write the contour, one thread per ranked segment; the first moments are
            // accumulated on the way (every thread takes part in the reduction)
            if (sh_keep >= 0) {
                int32_t *dst = points + ((size_t)z * max_points + sh_off) * 2;
                long long a10 = 0, a01 = 0;
                for (int i = tid; i < ncp; i += TRACE_THREADS) {
                    if (nbuf[i] < 0) continue;
                    const int st = cp_state[i];
                    int len;
                    SegSums S;
                    trace_segment<true, LOG2_HT>(bv, (st >> 3) & 8191, st >> 16, st & 7, xs, ys, rmask, ht_keys,
                                        ht_vals, len, S, dst + 2 * (size_t)nbuf[i], max_steps);
                    a10 += S.a10; a01 += S.a01;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a10 += __shfl_xor_sync(0xffffffffu, a10, o);
                    a01 += __shfl_xor_sync(0xffffffffu, a01, o);
                }
                if ((tid & 31) == 0) { sh_part[1][tid >> 5] = a10; sh_part[2][tid >> 5] = a01; }
                __syncthreads();
                if (tid == 0) {
                    long long t10 = 0, t01 = 0;
                    for (int q = 0; q < TRACE_THREADS / 32; q++) { t10 += sh_part[1][q]; t01 += sh_part[2][q]; }
                    gofpid_blob *kb = kept + (size_t)z * max_kept + sh_keep;
                    kb->a10 = t10; kb->a01 = t01;
                }
            }
        }
#ifdef GOFPID_TRACE_DEBUG
        __syncthreads();
        DBG_T(5);
        if (tid == 0) {
            const int k = atomicAdd(&g_trace_dbg_n, 1);
            if (k < 64) {
                long long *d = g_trace_dbg + k * 16;
                d[0] = ci; d[1] = bw; d[2] = bh; d[3] = ncp; d[4] = rmask; d[5] = sh_keep;
                for (int q = 0; q < 5; q++) d[6 + q] = dbg_t[q + 1] - dbg_t[q];
                d[11] = blockIdx.x; d[12] = dbg_t[6] - dbg_t[2]; d[13] = dbg_t[7] - dbg_t[2];
            }
        }
#endif
    }
}

#ifdef GOFPID_TRACE_DEBUG
extern "C" __attribute__((visibility("default"))) void gofpid_debug_rc_dump(long long *out)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_rc_dbg, sizeof(long long) * 16);
}
extern "C" __attribute__((visibility("default"))) int gofpid_debug_trace_dump(long long *out)
{
    int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_trace_dbg_n, sizeof(int));
    cudaMemcpyFromSymbol(out, g_trace_dbg, sizeof(long long) * 64 * 16);
    int z = 0;
    cudaMemcpyToSymbol(g_trace_dbg_n, &z, sizeof(int));
    return n;
}
#endif

// ------------------------------------------------------------------------------------
// 6. RETR_EXTERNAL: drop blobs that lie inside a hole of another blob
// ------------------------------------------------------------------------------------
// cv.findContours(RETR_EXTERNAL) returns only components that are not enclosed by another
// component's outer border.  A blob k is enclosed iff some blob j has a strictly larger box and
// k's first pixel lies inside j's outer contour polygon (even-odd crossing count of the ray to
// +x; contour edges are unit steps, so an edge crosses the ray exactly at its endpoint with
// y == fy).  Enclosing blobs are at least as large as what they enclose, so testing the kept
// blobs against each other is enough.  One warp per blob.
__global__ void __launch_bounds__(256)
nest_kernel(gofpid_frame_info *info, gofpid_blob *kept, const int32_t *points, int max_kept, int max_points,
            int W)
{
    constexpr int CAP = 1024; // boxes staged in shared memory (further blobs are read from global)
    __shared__ short4 box[CAP];
    const int z = blockIdx.y;
    gofpid_frame_info *fi = info + z;
    const int nk = min(fi->n_kept, max_kept);
    if (nk < 2) return;
    gofpid_blob *kb = kept + (size_t)z * max_kept;
    const int32_t *pts = points + (size_t)z * max_points * 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < min(nk, CAP); i += blockDim.x)
        box[i] = make_short4((short)kb[i].x0, (short)kb[i].y0, (short)kb[i].x1, (short)kb[i].y1);
    __syncthreads();
    auto get_box = [&](int i) -> short4 {
        if (i < CAP) return box[i];
        return make_short4((short)kb[i].x0, (short)kb[i].y0, (short)kb[i].x1, (short)kb[i].y1);
    };
    for (int k = blockIdx.x * nwarps + warp; k < nk; k += gridDim.x * nwarps) { // (several CTAs: frames with very many blobs)
        const short4 bk = get_box(k);
        bool nested = false;
        int fy = 0, fx = 0;
        bool have_first = false;
        for (int j0 = 0; j0 < nk && !nested; j0 += 32) {
            const int j = j0 + lane;
            bool hit = false;
            if (j < nk && j != k) {
                const short4 bj = get_box(j);
                hit = bj.x < bk.x && bj.y < bk.y && bj.z > bk.z && bj.w > bk.w;
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m && !nested) {
                const int jj = j0 + __ffs(m) - 1;
                m &= m - 1;
                if (!have_first) { const int f = kb[k].first; fy = f / W; fx = f - fy * W; have_first = true; }
                const int np = kb[jj].n_points;
                if (np <= 0) continue;
                const int32_t *cj = pts + (size_t)kb[jj].point_offset * 2;
                unsigned crossings = 0;
                for (int i = lane; i < np; i += 32) {
                    const int i2 = i + 1 == np ? 0 : i + 1;
                    const int2 pa = *reinterpret_cast<const int2 *>(cj + 2 * i);
                    const int2 pb = *reinterpret_cast<const int2 *>(cj + 2 * i2);
                    if ((pa.y > fy) != (pb.y > fy)) {
                        const int xe = pa.y == fy ? pa.x : pb.x; // the endpoint on the ray's row
                        if (fx < xe) crossings++;
                    }
                }
                crossings = __reduce_add_sync(0xffffffffu, crossings);
                nested = crossings & 1u;
            }
        }
        if (lane == 0 && nested) {
            kb[k].nested = 1;
            atomicAdd(&fi->reserved[0], 1);
        }
    }
}

// ------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------
int run_list_alloc(RunList *R, int H, int bit_pitch, int n_streams)
{
    memset(R, 0, sizeof(*R));
    R->max_runs = run_ccl_capacity(H, bit_pitch);
    R->row_pitch = (H + 3) & ~3;
    const size_t n = (size_t)(R->max_runs ? R->max_runs : 8) * n_streams;
    GOFPID_CUDA_TRY(cudaMalloc(&R->xs, n * sizeof(uint16_t)));
    GOFPID_CUDA_TRY(cudaMalloc(&R->xe, n * sizeof(uint16_t)));
    GOFPID_CUDA_TRY(cudaMalloc(&R->y, n * sizeof(uint16_t)));
    GOFPID_CUDA_TRY(cudaMalloc(&R->rowstart, (size_t)R->row_pitch * n_streams * sizeof(int)));
    GOFPID_CUDA_TRY(cudaMalloc(&R->rowend, (size_t)R->row_pitch * n_streams * sizeof(int)));
    GOFPID_CUDA_TRY(cudaMalloc(&R->label, n * sizeof(int32_t)));
    GOFPID_CUDA_TRY(cudaMalloc(&R->byc, n * sizeof(uint2)));
    GOFPID_CUDA_TRY(cudaMalloc(&R->count, (size_t)n_streams * sizeof(int)));
    GOFPID_CUDA_TRY(cudaMemset(R->count, 0, (size_t)n_streams * sizeof(int)));
    GOFPID_CUDA_TRY(cudaMalloc(&R->n_labelled, (size_t)n_streams * sizeof(int)));
    GOFPID_CUDA_TRY(cudaMemset(R->n_labelled, 0, (size_t)n_streams * sizeof(int)));
    return GOFPID_OK;
}
void run_list_free(RunList *R)
{
    cudaFree(R->xs); cudaFree(R->xe); cudaFree(R->y); cudaFree(R->rowstart); cudaFree(R->rowend); cudaFree(R->count); cudaFree(R->label); cudaFree(R->byc); cudaFree(R->n_labelled);
    memset(R, 0, sizeof(*R));
}

// run-based labelling in shared memory; streams that do not fit raise their fallback flag
static int launch_labelling(ImageU8 mask, const uint32_t *bitimg, int bit_pitch, size_t bit_stream,
                            const RunList &R, int32_t *labels, size_t LP, size_t labels_stream,
                            int32_t *comp, int max_blobs, int H, int W, int n_streams,
                            gofpid_frame_info *info, const FrameTick *tick, int force_tiled, cudaStream_t s,
                            const cudaGraphConditionalHandle *cond, int runs_cap)
{
    int max_runs = R.max_runs;
    if (runs_cap >= 0 && runs_cap < max_runs) max_runs = runs_cap & ~7; // tests: force the tiled fallback with a tiny capacity
    if (force_tiled) max_runs = 0;
    if (R.max_runs > 0) {
        runs_extract_kernel<<<dim3((H + RX_ROWS - 1) / RX_ROWS, n_streams), RX_THREADS, 0, s>>>(
            bitimg, bit_pitch, bit_stream, R, H);
        GOFPID_CUDA_TRY(cudaGetLastError());
    }
    GOFPID_CUDA_TRY(cudaFuncSetAttribute(runs_ccl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)RC_SMEM_BUDGET));
    runs_ccl_kernel<<<n_streams, RC_THREADS, max_runs ? run_ccl_smem(H, R.max_runs) : 0, s>>>(
        R, labels, LP, labels_stream, comp, info, max_blobs, max_runs, tick, H, W, cond ? *cond : 0, cond ? 1 : 0);
    GOFPID_CUDA_TRY(cudaGetLastError());
    return GOFPID_OK;
}

// tiled labelling of the streams that raised the fallback flag (all of them with force = 1)
static int launch_tiled_labelling(ImageU8 mask, int32_t *labels, size_t LP, size_t labels_stream, int H, int W,
                                  int n_streams, gofpid_frame_info *info, int force, cudaStream_t s)
{
    const dim3 tgrid((W + CCL_TW - 1) / CCL_TW, (H + CCL_TH - 1) / CCL_TH, n_streams);
    label_tiles_kernel<<<tgrid, dim3(16, CCL_TH), 0, s>>>(mask, labels, LP, labels_stream, H, W, info, force);
    GOFPID_CUDA_TRY(cudaGetLastError());
    merge_tiles_kernel<<<tgrid, CCL_TW + 2 * CCL_TH, 0, s>>>(mask, labels, LP, labels_stream, H, W, info, force);
    GOFPID_CUDA_TRY(cudaGetLastError());
    return GOFPID_OK;
}

// the tiled fallback: labelling, flatten + slots, statistics
static int launch_fallback(const BlobBuffers &B, ImageU8 mask, int H, int W, int n_streams, cudaStream_t s)
{
    GOFPID_TRY(launch_tiled_labelling(mask, B.labels, B.labels_pitch, B.labels_stream, H, W, n_streams, B.info, 0, s));
    const dim3 pgrid((W + 127) / 128, (H + 7) / 8, n_streams);
    flatten_kernel<<<pgrid, dim3(32, 8), 0, s>>>(mask, B.labels, B.labels_pitch, B.labels_stream,
                                                 B.slot, B.comp, B.info, B.max_blobs, B.bitimg,
                                                 B.bit_pitch, B.bit_stream, H, W);
    GOFPID_CUDA_TRY(cudaGetLastError());
    stats_kernel<<<pgrid, dim3(32, 8), 0, s>>>(mask, B.labels, B.labels_pitch, B.labels_stream,
                                               B.slot, B.comp, B.info, B.max_blobs, H, W);
    GOFPID_CUDA_TRY(cudaGetLastError());
    return GOFPID_OK;
}

// Results of a frame -> pinned host memory, written by the SMs (the host buffers are mapped into the
// device's address space): exactly the records and points the frame produced, and no copy-engine
// work that would queue up with the host->device frame copies.
__global__ void __launch_bounds__(256)
results_to_host_kernel(const gofpid_frame_info *info, const gofpid_blob *kept, const int32_t *points,
                       gofpid_frame_info *h_info, gofpid_blob *h_kept, int32_t *h_points, int max_kept, int max_points)
{
    const int z = blockIdx.y;
    const int nk = min(max(info[z].n_kept, 0), max_kept), np = min(max(info[z].n_points, 0), max_points);
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    static_assert(sizeof(gofpid_frame_info) == 32 && sizeof(gofpid_blob) == 64, "records are copied as 16-byte words");
    if (t < 2) reinterpret_cast<uint4 *>(h_info + z)[t] = reinterpret_cast<const uint4 *>(info + z)[t];
    const uint4 *ks = reinterpret_cast<const uint4 *>(kept + (size_t)z * max_kept);
    uint4 *kd = reinterpret_cast<uint4 *>(h_kept + (size_t)z * max_kept);
    for (int i = t; i < nk * 4; i += nt) kd[i] = ks[i];
    const uint2 *ps = reinterpret_cast<const uint2 *>(points + (size_t)z * max_points * 2);
    uint2 *pd = reinterpret_cast<uint2 *>(h_points + (size_t)z * max_points * 2);
    for (int i = t; i < np; i += nt) pd[i] = ps[i];
}
int launch_results_to_host(const BlobBuffers &B, gofpid_frame_info *h_info, gofpid_blob *h_kept, int32_t *h_points,
                           int n_streams, cudaStream_t s)
{
    const dim3 grid(n_streams == 1 ? 8 : (n_streams <= 8 ? 4 : 2), n_streams);
    results_to_host_kernel<<<grid, 256, 0, s>>>(B.info, B.kept, B.points, h_info, h_kept, h_points, B.max_kept, B.max_points);
    GOFPID_CUDA_TRY(cudaGetLastError());
    return GOFPID_OK;
}

int launch_blobs(const BlobBuffers &B, ImageU8 mask, int H, int W, int n_streams,
                 const FrameTick *tick, cudaStream_t s, int *n_launches, bool bits_ready,
                 const cudaGraphConditionalHandle *cond, int nest_ctas)
{
    int n = 0;
    const dim3 pgrid((W + 127) / 128, (H + 7) / 8, n_streams);
    if (!bits_ready) {
        pack_bits_kernel<<<pgrid, dim3(32, 8), 0, s>>>(mask, B.bitimg, B.bit_pitch, B.bit_stream, H, W);
        GOFPID_CUDA_TRY(cudaGetLastError());
        n++;
    }
    GOFPID_TRY(launch_labelling(mask, B.bitimg, B.bit_pitch, B.bit_stream, B.runs, B.labels, B.labels_pitch,
                                B.labels_stream, B.comp, B.max_blobs, H, W, n_streams, B.info, tick, 0, s, cond, B.runs_cap));
    if (cond) {
        // `s` is being captured: the fallback kernels go into the body of an IF node whose condition
        // runs_ccl_kernel raises, so a normal frame does not even launch them
        cudaStreamCaptureStatus st;
        cudaGraph_t g = nullptr;
        const cudaGraphNode_t *deps = nullptr;
        size_t nd = 0;
        GOFPID_CUDA_TRY(cudaStreamGetCaptureInfo(s, &st, nullptr, &g, &deps, &nd));
        if (st != cudaStreamCaptureStatusActive || !g) { set_error("internal: stage B is not being captured"); return GOFPID_ERR_CUDA; }
        cudaGraphNodeParams cp = {};
        cp.type = cudaGraphNodeTypeConditional;
        cp.conditional.handle = *cond;
        cp.conditional.type = cudaGraphCondTypeIf;
        cp.conditional.size = 1;
        cudaGraphNode_t cn;
        GOFPID_CUDA_TRY(cudaGraphAddNode(&cn, g, deps, nd, &cp));
        cudaGraph_t body = cp.conditional.phGraph_out[0];
        GOFPID_CUDA_TRY(cudaStreamUpdateCaptureDependencies(s, &cn, 1, cudaStreamSetCaptureDependencies));
        cudaStream_t aux;
        GOFPID_CUDA_TRY(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
        cudaError_t e = cudaStreamBeginCaptureToGraph(aux, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
        int rc = GOFPID_OK;
        if (e == cudaSuccess) {
            rc = launch_fallback(B, mask, H, W, n_streams, aux);
            cudaGraph_t tmp = nullptr;
            e = cudaStreamEndCapture(aux, &tmp);
        }
        cudaStreamDestroy(aux);
        if (rc != GOFPID_OK) return rc;
        if (e != cudaSuccess) { set_error("capturing the fallback body failed: %s", cudaGetErrorString(e)); return GOFPID_ERR_CUDA; }
    } else {
        GOFPID_TRY(launch_fallback(B, mask, H, W, n_streams, s));
        n += 4;
    }
    constexpr size_t sm = trace_smem_bytes(TRACE_LOG2_CP);
    GOFPID_CUDA_TRY(cudaFuncSetAttribute(trace_kernel<TRACE_THREADS, TRACE_LOG2_CP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const dim3 cgrid(n_streams >= 64 ? 8 : (n_streams >= 8 ? 32 : 148), n_streams);
    trace_kernel<TRACE_THREADS, TRACE_LOG2_CP><<<cgrid, TRACE_THREADS, sm, s>>>(
        B.labels, B.labels_pitch, B.labels_stream, B.runs, B.comp, B.info, B.kept, B.points, B.max_blobs,
        B.max_kept, B.max_points, B.area_min, B.bitimg, B.bit_pitch, B.bit_stream, H, W);
    GOFPID_CUDA_TRY(cudaGetLastError());
    nest_kernel<<<dim3(nest_ctas > 0 ? nest_ctas : 1, n_streams), 256, 0, s>>>(B.info, B.kept, B.points, B.max_kept, B.max_points, W);
    GOFPID_CUDA_TRY(cudaGetLastError());
    if (n_launches) *n_launches = n + (B.runs.max_runs > 0 ? 4 : 3);
    return GOFPID_OK;
}

int launch_fill_labels(const RunList &R, int z, const gofpid_frame_info *info, int32_t *labels, size_t LP,
                       cudaStream_t s)
{
    if (R.max_runs <= 0) return GOFPID_OK;
    fill_labels_kernel<<<(R.max_runs + 7) / 8, 256, 0, s>>>(R, z, info, labels, LP);
    GOFPID_CUDA_TRY(cudaGetLastError());
    return GOFPID_OK;
}

// canonical dense labels for the stage-level entry point
__global__ void canonical_labels_kernel(const uint8_t *mask, size_t mask_pitch,
                                        const int32_t *labels, size_t LP, int32_t *out,
                                        size_t out_pitch, int H, int W)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    int v = 0;
    if (mask[(size_t)y * mask_pitch + x] != 0) {
        const int r = uf_find_g(labels, labels[(size_t)y * LP + x]) - 1;
        const int ry = r / (int)LP, rx = r - ry * (int)LP;
        v = ry * W + rx + 1;
    }
    out[(size_t)y * out_pitch + x] = v;
}

int launch_canonical_labels(const uint8_t *mask, size_t mask_pitch, const int32_t *labels, size_t LP,
                            int32_t *out, size_t out_pitch_elems, int H, int W, cudaStream_t s)
{
    canonical_labels_kernel<<<dim3((W + 255) / 256, H), 256, 0, s>>>(mask, mask_pitch, labels, LP, out,
                                                                    out_pitch_elems, H, W);
    GOFPID_CUDA_TRY(cudaGetLastError());
    return GOFPID_OK;
}

int launch_ccl_only(const uint8_t *d_mask, size_t mask_pitch, int32_t *d_labels,
                    size_t labels_pitch_elems, int H, int W, cudaStream_t s)
{
    // stage a padded copy of the mask (the kernels read whole words), its bit image, provisional
    // labels and component records; same labelling sequence as the pipeline.
    // GOFPID_CCL_FORCE_TILED=1 sends everything through the tiled fallback kernels.
    const size_t MP = round_up((size_t)W, 128), LP = round_up((size_t)W, 4);
    const int bit_pitch = (int)round_up((size_t)(W + 31) / 32 + 2, 4);
    const size_t bit_words = (size_t)bit_pitch * (H + 2);
    const int max_blobs = 1 << 16;
    // (stage-level test entry point: its two knobs are read per call, it is not on the pipeline's path)
    const char *e = getenv("GOFPID_CCL_FORCE_TILED");
    const int force = e && e[0] == '1';
    const char *rc_e = getenv("GOFPID_RUNS_CAP");
    const int runs_cap = rc_e && atoi(rc_e) >= 0 ? atoi(rc_e) : -1;
    uint8_t *pm = nullptr;
    int32_t *pl = nullptr, *comp = nullptr, *slot = nullptr;
    uint32_t *bits = nullptr;
    gofpid_frame_info *fi = nullptr;
    RunList R;
    if (run_list_alloc(&R, H, bit_pitch, 1) != GOFPID_OK) return GOFPID_ERR_CUDA;
    GOFPID_CUDA_TRY(cudaMalloc(&pm, MP * H));
    GOFPID_CUDA_TRY(cudaMalloc(&pl, LP * H * sizeof(int32_t)));
    GOFPID_CUDA_TRY(cudaMalloc(&slot, LP * H * sizeof(int32_t)));
    GOFPID_CUDA_TRY(cudaMalloc(&bits, bit_words * sizeof(uint32_t)));
    GOFPID_CUDA_TRY(cudaMalloc(&comp, (size_t)max_blobs * CR_SIZE * sizeof(int32_t)));
    GOFPID_CUDA_TRY(cudaMalloc(&fi, sizeof(gofpid_frame_info)));
    int rc = GOFPID_OK;
    do {
        if (cudaMemsetAsync(bits, 0, bit_words * sizeof(uint32_t), s) != cudaSuccess ||
            cudaMemsetAsync(pm, 0, MP * H, s) != cudaSuccess ||
            cudaMemcpy2DAsync(pm, MP, d_mask, mask_pitch, W, H, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
            set_error("mask staging copy failed"); rc = GOFPID_ERR_CUDA; break;
        }
        ImageU8 m{pm, MP, 0};
        pack_bits_kernel<<<dim3((W + 127) / 128, (H + 7) / 8, 1), dim3(32, 8), 0, s>>>(m, bits, bit_pitch, 0, H, W);
        rc = launch_labelling(m, bits, bit_pitch, 0, R, pl, LP, 0, comp, max_blobs, H, W, 1, fi, nullptr, force, s, nullptr, runs_cap);
        if (rc) break;
        rc = launch_tiled_labelling(m, pl, LP, 0, H, W, 1, fi, 0, s);
        if (rc) break;
        rc = launch_fill_labels(R, 0, fi, pl, LP, s);
        if (rc) break;
        rc = launch_canonical_labels(pm, MP, pl, LP, d_labels, labels_pitch_elems, H, W, s);
        if (rc) break;
        if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("labelling failed: %s", cudaGetErrorString(cudaGetLastError())); rc = GOFPID_ERR_CUDA; }
    } while (0);
    cudaFree(pm); cudaFree(pl); cudaFree(slot); cudaFree(bits); cudaFree(comp); cudaFree(fi);
    run_list_free(&R);
    return rc;
}

} // namespace gofpid
