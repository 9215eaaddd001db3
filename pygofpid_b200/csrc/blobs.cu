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

namespace gofpid {

constexpr int CCL_TW = 64;  // one 64-bit word of class bits per tile row
constexpr int CCL_TH = 32;
constexpr int CCL_THREADS = 16 * CCL_TH; // each thread owns 4 consecutive pixels of a row

constexpr int TRACE_THREADS = 256;
constexpr int TRACE_SMEM_WORDS = 32 * 1024; // 128 KB bitmap: boxes up to ~1 Mpx
constexpr int TRACE_MAX_CP = 1024;          // checkpoints per contour
constexpr int TRACE_HT = 2048;              // hash slots (power of two, >= 2 * TRACE_MAX_CP)
constexpr size_t TRACE_SMEM_BYTES = (size_t)TRACE_SMEM_WORDS * 4 + (size_t)TRACE_MAX_CP * (32 + 8) +
                                    (size_t)TRACE_HT * 8;

// component record layout (8 int32)
enum { CR_FIRST = 0, CR_X0, CR_Y0, CR_X1, CR_Y1, CR_AREA, CR_PAD0, CR_PAD1, CR_SIZE };

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
                   gofpid_frame_info *info, const FrameTick *tick)
{
    __shared__ int lab[CCL_TW * CCL_TH];
    __shared__ unsigned int Fw[CCL_TH][2]; // foreground bits of each tile row (two 32-bit halves)
    const int tx = threadIdx.x, ty = threadIdx.y; // blockDim = (16, CCL_TH)
    const int lane = (ty * 16 + tx) & 31;
    const int x0 = blockIdx.x * CCL_TW, y0 = blockIdx.y * CCL_TH;
    const int lx = tx * 4, x = x0 + lx, y = y0 + ty;
    const uint8_t *mrow = mask.base + (size_t)blockIdx.z * mask.stream_stride + (size_t)y * mask.pitch;

    if (blockIdx.x == 0 && blockIdx.y == 0 && tx == 0 && ty == 0) {
        gofpid_frame_info *fi = info + blockIdx.z;
        fi->n_components = 0; fi->n_candidates = 0; fi->n_kept = 0; fi->n_points = 0;
        fi->overflow = 0; fi->frame_index = tick ? tick->nframes : 0;
    }

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

    const int wvalid = min(CCL_TW, W - x0);                 // valid columns of this tile
    const unsigned long long V = bits_below(wvalid);
    const bool row_ok = y < H;
    const unsigned long long F = ((unsigned long long)Fw[ty][1] << 32) | Fw[ty][0];
    unsigned long long Fp = 0, Vp = 0;
    if (ty > 0) { Fp = ((unsigned long long)Fw[ty - 1][1] << 32) | Fw[ty - 1][0]; Vp = V; }

    // provisional labels: every pixel points at its run start; a background run that touches
    // the frame border is the virtual OUTSIDE node (0)
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int c = lx + j;
        const int li = ty * CCL_TW + c;
        if (!row_ok || c >= wvalid) { lab[li] = li + 1; continue; }
        const bool fg = (F >> c) & 1ULL;
        const unsigned long long S = fg ? F : (~F & V);
        const int a = run_start(S, c);
        int v = ty * CCL_TW + a + 1;
        if (a == c && !fg) {
            const int b = run_end(S, c);
            if (y == 0 || y == H - 1 || x0 + a == 0 || x0 + b == W - 1) v = 0;
        }
        lab[li] = v;
    }
    __syncthreads();

    // vertical unions, one thread per run
    if (row_ok && ty > 0) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = lx + j;
            if (c >= wvalid) continue;
            const bool fg = (F >> c) & 1ULL;
            const unsigned long long S = fg ? F : (~F & V);
            if (run_start(S, c) != c) continue;
            const int b = run_end(S, c);
            const unsigned long long Sp = fg ? Fp : (~Fp & Vp);
            const int lo = fg ? max(c - 1, 0) : c, hi = fg ? min(b + 1, 63) : b;
            unsigned long long m = Sp & bits_below(hi + 1) & ~bits_below(lo);
            const int me = ty * CCL_TW + c + 1;
            while (m) {
                const int pos = __ffsll((long long)m) - 1;
                const int ap = run_start(Sp, pos), bp = run_end(Sp, pos);
                uf_union_s(lab, me, (ty - 1) * CCL_TW + ap + 1);
                m &= ~bits_below(bp + 1);
            }
        }
    }
    __syncthreads();

    if (y >= H || x >= W) return;
    int out[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int r = uf_find(lab, ty * CCL_TW + lx + j + 1);
        out[j] = r == 0 ? 0
                        : (int)((size_t)(y0 + (r - 1) / CCL_TW) * LP + x0 + (r - 1) % CCL_TW) + 1;
    }
    int32_t *lrow = labels + (size_t)blockIdx.z * labels_stream + (size_t)y * LP;
    *reinterpret_cast<int4 *>(lrow + x) = make_int4(out[0], out[1], out[2], out[3]);
}

// ------------------------------------------------------------------------------------
// 2. unions across tile borders
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CCL_TW + 2 * CCL_TH)
merge_tiles_kernel(ImageU8 mask, int32_t *labels, size_t LP, size_t labels_stream, int H, int W)
{
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
    const int c = cl(y, x);
    const int v = (int)((size_t)y * LP + x) + 1;
    const bool top = y == y0, left = x == x0, right = x == x0 + CCL_TW - 1;
    const int cw = cl(y, x - 1), cn = cl(y - 1, x);
    if (left && cw == c) uf_union_g(lab, v, v - 1);
    if (top && cn == c) uf_union_g(lab, v, v - (int)LP);
    if (c == 1 && cn != 1) {
        if ((top || left) && cw != 1 && cl(y - 1, x - 1) == 1) uf_union_g(lab, v, v - (int)LP - 1);
        if ((top || right) && cl(y - 1, x + 1) == 1) uf_union_g(lab, v, v - (int)LP + 1);
    }
}

// ------------------------------------------------------------------------------------
// 3. flatten + component slots
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
flatten_kernel(ImageU8 mask, int32_t *labels, size_t LP, size_t labels_stream, int32_t *slot,
               int32_t *comp, gofpid_frame_info *info, int max_blobs, int H, int W)
{
    const int x = (blockIdx.x * 32 + threadIdx.x) * 4, y = blockIdx.y * 8 + threadIdx.y;
    if (y >= H || x >= W) return;
    int32_t *lab = labels + (size_t)blockIdx.z * labels_stream;
    const uint8_t *mrow = mask.base + (size_t)blockIdx.z * mask.stream_stride + (size_t)y * mask.pitch;
    const uint32_t mw = *reinterpret_cast<const uint32_t *>(mrow + x);
    int4 v4 = *reinterpret_cast<const int4 *>(lab + (size_t)y * LP + x);
    int v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (x + j >= W) continue;
        const int self = (int)((size_t)y * LP + x + j) + 1;
        const int r = uf_find_g(lab, v[j]);
        v[j] = r;
        if (r == self && ((mw >> (8 * j)) & 255) != 0) {
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
    // roots are final here (merging finished with the previous kernel), so writing the root
    // over a non-root entry only shortens paths other threads may still be walking
    *reinterpret_cast<int4 *>(lab + (size_t)y * LP + x) = make_int4(v[0], v[1], v[2], v[3]);
}

// ------------------------------------------------------------------------------------
// 4. component statistics
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stats_kernel(ImageU8 mask, const int32_t *labels, size_t LP, size_t labels_stream,
             const int32_t *slot, int32_t *comp, int max_blobs, int H, int W)
{
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
        const int4 v4 = *reinterpret_cast<const int4 *>(lab + (size_t)y * LP + x);
        v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
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
__device__ __constant__ int8_t DIR_DX[8] = {1, 1, 0, -1, -1, -1, 0, 1};
__device__ __constant__ int8_t DIR_DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};

struct BitmapView { // component staged in smem with a 1-pixel empty frame
    const uint32_t *bits;
    int row_words;
    int bx0, by0; // image coords of bitmap (1, 1)
    int sx, sy;   // the component's first pixel (start of the border)
    __device__ __forceinline__ unsigned nb(int x, int y) const
    {
        const int bx = x - bx0, by = y - by0; // bitmap column of x-1 is bx, row of y-1 is by
        const int w = bx >> 5, sh = bx & 31;
        const uint32_t *r = bits + (size_t)by * row_words + w;
        const unsigned t = __funnelshift_r(r[0], r[1], sh) & 7u;
        const unsigned m = __funnelshift_r(r[row_words], r[row_words + 1], sh) & 7u;
        const unsigned b = __funnelshift_r(r[2 * row_words], r[2 * row_words + 1], sh) & 7u;
        return ((m >> 2) & 1u) | (((t >> 2) & 1u) << 1) | (((t >> 1) & 1u) << 2) | ((t & 1u) << 3) |
               ((m & 1u) << 4) | ((b & 1u) << 5) | (((b >> 1) & 1u) << 6) | (((b >> 2) & 1u) << 7);
    }
};
struct GlobalView { // fallback for components whose box does not fit in shared memory
    const int32_t *lab;
    size_t LP;
    int H, W, root;
    __device__ __forceinline__ unsigned nb(int x, int y) const
    {
        unsigned r = 0;
#pragma unroll
        for (int d = 0; d < 8; d++) {
            const int xx = x + DIR_DX[d], yy = y + DIR_DY[d];
            if (xx >= 0 && xx < W && yy >= 0 && yy < H && lab[(size_t)yy * LP + xx] == root)
                r |= 1u << d;
        }
        return r;
    }
};

struct TraceResult {
    int n;
    long long a00, a10, a01;
};

// Follows the outer border from the component's raster-first pixel (xs, ys) the way
// cv.findContours does for an outer border.  STORE: write the points, else only count and
// accumulate the Green sums of cv.contourArea / cv.moments.
template <bool STORE, typename View>
__device__ TraceResult trace_border(const View &V, int xs, int ys, int32_t *pts)
{
    TraceResult R; R.n = 0; R.a00 = 0; R.a10 = 0; R.a01 = 0;
    // first neighbour clockwise from west
    const unsigned nb0 = V.nb(xs, ys);
    int s = -1;
#pragma unroll
    for (int k = 3; k >= -3; k--) { // 3,2,1,0,7,6,5
        const int d = k & 7;
        if (s < 0 && ((nb0 >> d) & 1u)) s = d;
    }
    if (s < 0) { // isolated pixel
        if (STORE) { pts[0] = xs; pts[1] = ys; }
        R.n = 1;
        return R;
    }
    const int x1 = xs + DIR_DX[s], y1 = ys + DIR_DY[s];
    int cx = xs, cy = ys;
    int px = 0, py = 0; // previous emitted point
    while (true) {
        const unsigned nb = V.nb(cx, cy);
        // counter-clockwise search starting after the direction we came from
        const unsigned rot = ((nb | (nb << 8)) >> ((s + 1) & 7)) & 0xffu;
        s = (s + 1 + (__ffs(rot) - 1)) & 7;
        const int nx = cx + DIR_DX[s], ny = cy + DIR_DY[s];
        if (STORE) { pts[2 * R.n] = cx; pts[2 * R.n + 1] = cy; }
        else if (R.n > 0) {
            const long long dxy = (long long)px * cy - (long long)cx * py;
            R.a00 += dxy; R.a10 += dxy * (px + cx); R.a01 += dxy * (py + cy);
        }
        px = cx; py = cy;
        R.n++;
        if (nx == xs && ny == ys && cx == x1 && cy == y1) break;
        cx = nx; cy = ny;
        s = (s + 4) & 7;
    }
    if (!STORE) { // closing edge: last point -> first point
        const long long dxy = (long long)px * ys - (long long)xs * py;
        R.a00 += dxy; R.a10 += dxy * (px + xs); R.a01 += dxy * (py + ys);
    }
    return R;
}

// ---- parallel tracing with checkpoints ------------------------------------------------
// A visit of the border follower is the state (pixel c, back direction b): the next visit is
// (c + D[s'], s' + 4) with s' the first foreground direction counter-clockwise after b.  On a
// border every visit has b foreground and b+1 background, so visit states can be enumerated
// from the 3x3 neighbourhood alone, and the outer border is the cycle through the start state
// (first pixel, first foreground direction clockwise from west).  Pixels selected by a hash of
// their coordinates become CHECKPOINTS; one thread per checkpoint follows the border to the
// next checkpoint (a short segment), one thread links the segments starting from the start
// state, and the segments are then written out in parallel at their offsets.
struct Segment {
    int len, next;
    long long a00, a10, a01;
};

__device__ __forceinline__ bool cp_hash(int x, int y, unsigned rmask)
{
    unsigned h = (unsigned)x * 0x9E3779B1u ^ (unsigned)y * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
    return (h & rmask) == 0;
}
__device__ __forceinline__ int cp_key(int lx, int ly, int b) { return ((ly << 15) | (lx << 3) | b) + 1; }

__device__ __forceinline__ int ht_find(const int *keys, const int *vals, int key)
{
    unsigned h = ((unsigned)key * 0x9E3779B1u) >> 21; // 11 bits
    while (true) {
        const int k = keys[h & (TRACE_HT - 1)];
        if (k == key) return vals[h & (TRACE_HT - 1)];
        if (k == 0) return -1;
        h++;
    }
}
__device__ __forceinline__ void ht_insert(int *keys, int *vals, int key, int val)
{
    unsigned h = ((unsigned)key * 0x9E3779B1u) >> 21;
    while (true) {
        const int old = atomicCAS(&keys[h & (TRACE_HT - 1)], 0, key);
        if (old == 0 || old == key) { vals[h & (TRACE_HT - 1)] = val; return; }
        h++;
    }
}

// Follows the border from visit (x, y, b) until the next checkpoint visit; STORE writes the
// visited points.  Returns the id of the checkpoint reached.
template <bool STORE>
__device__ int trace_segment(const BitmapView &V, int x, int y, int b, unsigned rmask,
                             const int *keys, const int *vals, Segment &S, int32_t *pts, int max_steps)
{
    int n = 0;
    long long a00 = 0, a10 = 0, a01 = 0;
    int id = -1;
    while (n < max_steps) {
        const unsigned nb = V.nb(x, y);
        const unsigned rot = ((nb | (nb << 8)) >> ((b + 1) & 7)) & 0xffu;
        const int s = (b + 1 + (__ffs(rot) - 1)) & 7;
        const int nx = x + DIR_DX[s], ny = y + DIR_DY[s];
        if (STORE) { pts[2 * n] = x; pts[2 * n + 1] = y; }
        else {
            const long long dxy = (long long)x * ny - (long long)nx * y;
            a00 += dxy; a10 += dxy * (x + nx); a01 += dxy * (y + ny);
        }
        n++;
        x = nx; y = ny; b = (s + 4) & 7;
        if (cp_hash(x, y, rmask) || (x == V.sx && y == V.sy)) {
            id = ht_find(keys, vals, cp_key(x - V.bx0, y - V.by0, b));
            if (id >= 0) break;
        }
    }
    if (!STORE) { S.len = n; S.next = id; S.a00 = a00; S.a10 = a10; S.a01 = a01; }
    return id;
}

__global__ void __launch_bounds__(TRACE_THREADS)
trace_kernel(const int32_t *labels, size_t LP, size_t labels_stream, const int32_t *comp,
             gofpid_frame_info *info, gofpid_blob *kept, int32_t *points, int max_blobs,
             int max_kept, int max_points, double area_min, int H, int W)
{
    extern __shared__ __align__(16) uint32_t bits[]; // bitmap, then the tables below
    Segment *seg = reinterpret_cast<Segment *>(bits + TRACE_SMEM_WORDS);
    int *ht_keys = reinterpret_cast<int *>(seg + TRACE_MAX_CP);
    int *ht_vals = ht_keys + TRACE_HT;
    int *cp_state = ht_vals + TRACE_HT;     // packed (ly << 15 | lx << 3 | b)
    int *cp_off = cp_state + TRACE_MAX_CP;  // output offset of the segment, -1 = not on the outer border
    __shared__ int sh_ncp, sh_keep, sh_off;
    const int z = blockIdx.y;
    const int32_t *lab = labels + (size_t)z * labels_stream;
    gofpid_frame_info *fi = info + z;
    const int ncomp = min(fi->n_components, max_blobs);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int ci = blockIdx.x; ci < ncomp; ci += gridDim.x) {
        const int32_t *cr = comp + ((size_t)z * max_blobs + ci) * CR_SIZE;
        const int first = cr[CR_FIRST], bx0 = cr[CR_X0], by0 = cr[CR_Y0], bx1 = cr[CR_X1],
                  by1 = cr[CR_Y1];
        const int ys = first / (int)LP, xs = first - ys * (int)LP;
        const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
        // cv.contourArea <= (w-1)(h-1): skip boxes that cannot reach area_min
        if ((double)(bw - 1) * (double)(bh - 1) < area_min) continue;
        // RETR_EXTERNAL: the background left of the first pixel must reach the frame border
        if (xs > 0 && lab[first - 1] != 0) continue;
        if (tid == 0) atomicAdd(&fi->n_candidates, 1);
        const int root = first + 1;

        const int row_words = (bw + 2 + 31) / 32 + 1;
        const bool fits = (size_t)row_words * (bh + 2) <= (size_t)TRACE_SMEM_WORDS && bw < 4096 && bh < 4096;
        __syncthreads(); // the previous candidate's smem state is no longer in use

        if (!fits) {
            // serial fallback straight from the label image (boxes beyond the smem bitmap)
            if (tid == 0) {
                GlobalView gv; gv.lab = lab; gv.LP = LP; gv.H = H; gv.W = W; gv.root = root;
                TraceResult R = trace_border<false>(gv, xs, ys, nullptr);
                if (fabs((double)R.a00) * 0.5 >= area_min) {
                    int keep = atomicAdd(&fi->n_kept, 1);
                    const int off = atomicAdd(&fi->n_points, R.n);
                    if (keep >= max_kept) { atomicOr(&fi->overflow, 2); keep = -1; }
                    else if (off + R.n > max_points) { atomicOr(&fi->overflow, 4); R.n = 0; }
                    if (keep >= 0) {
                        gofpid_blob *kb = kept + (size_t)z * max_kept + keep;
                        kb->first = ys * W + xs;
                        kb->x0 = bx0; kb->y0 = by0; kb->x1 = bx1; kb->y1 = by1;
                        kb->area_px = cr[CR_AREA];
                        kb->n_points = R.n; kb->point_offset = off;
                        kb->a00 = R.a00; kb->a10 = R.a10; kb->a01 = R.a01;
                        if (R.n > 0) trace_border<true>(gv, xs, ys, points + ((size_t)z * max_points + off) * 2);
                    }
                }
            }
            continue;
        }

        // ---- stage the component as a bitmap with a 1-pixel empty frame ---------------------
        const int total = row_words * (bh + 2);
        for (int i = warp; i < total; i += TRACE_THREADS / 32) {
            const int r = i / row_words, wc = i - r * row_words;
            const int yy = by0 - 1 + r, xx = bx0 - 1 + wc * 32 + lane;
            const bool on = yy >= by0 && yy <= by1 && xx >= bx0 && xx <= bx1 &&
                            lab[(size_t)yy * LP + xx] == root;
            const unsigned w = __ballot_sync(0xffffffffu, on);
            if (lane == 0) bits[i] = w;
        }
        BitmapView bv; bv.bits = bits; bv.row_words = row_words; bv.bx0 = bx0; bv.by0 = by0;
        bv.sx = xs; bv.sy = ys;
        __syncthreads();

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
                    kb->area_px = cr[CR_AREA]; kb->n_points = 1; kb->point_offset = off;
                    kb->a00 = 0; kb->a10 = 0; kb->a01 = 0;
                    int32_t *dst = points + ((size_t)z * max_points + off) * 2;
                    dst[0] = xs; dst[1] = ys;
                }
            }
            continue;
        }

        // ---- enumerate checkpoints (retry with a sparser hash when there are too many) -----------
        unsigned rmask = 3u;
        {
            const int est = 2 * (bw + bh);
            while ((int)(rmask + 1) * (TRACE_MAX_CP / 2) < est) rmask = (rmask << 1) | 1u;
        }
        while (true) {
            for (int i = tid; i < TRACE_HT; i += TRACE_THREADS) ht_keys[i] = 0;
            if (tid == 0) {
                sh_ncp = 1;
                cp_state[0] = ((ys - by0) << 15) | ((xs - bx0) << 3) | b0;
            }
            __syncthreads();
            if (tid == 0) ht_insert(ht_keys, ht_vals, cp_key(xs - bx0, ys - by0, b0), 0);
            for (int i = tid; i < bw * bh; i += TRACE_THREADS) {
                const int ly = i / bw, lx = i - ly * bw;
                const int x = bx0 + lx, y = by0 + ly;
                if (!cp_hash(x, y, rmask)) continue;
                // is the pixel part of the component?
                const int bx = lx + 1;
                if (!((bits[(size_t)(ly + 1) * row_words + (bx >> 5)] >> (bx & 31)) & 1u)) continue;
                const unsigned nb = bv.nb(x, y);
                // visit states: direction b foreground, b+1 background
                unsigned starts = nb & ~((nb >> 1) | (nb << 7)) & 0xffu;
                while (starts) {
                    const int b = __ffs(starts) - 1;
                    starts &= starts - 1;
                    if (x == xs && y == ys && b == b0) continue; // already checkpoint 0
                    const int id = atomicAdd(&sh_ncp, 1);
                    if (id < TRACE_MAX_CP) {
                        cp_state[id] = (ly << 15) | (lx << 3) | b;
                        ht_insert(ht_keys, ht_vals, cp_key(lx, ly, b), id);
                    }
                }
            }
            __syncthreads();
            if (sh_ncp <= TRACE_MAX_CP) break;
            rmask = (rmask << 1) | 1u; // too many: halve the density and start over
            __syncthreads();
        }
        const int ncp = sh_ncp;

        // ---- pass 1: every checkpoint follows the border to the next checkpoint -------------------
        const int max_steps = 4 * bw * bh + 16;
        for (int i = tid; i < ncp; i += TRACE_THREADS) {
            const int st = cp_state[i];
            Segment S;
            trace_segment<false>(bv, bx0 + ((st >> 3) & 4095), by0 + (st >> 15), st & 7, rmask, ht_keys,
                                 ht_vals, S, nullptr, max_steps);
            seg[i] = S;
            cp_off[i] = -1;
        }
        __syncthreads();

        // ---- link the segments of the outer border, starting from the start state -----------------
        if (tid == 0) {
            int id = 0, off = 0, guard = 0;
            long long a00 = 0, a10 = 0, a01 = 0;
            do {
                cp_off[id] = off;
                off += seg[id].len;
                a00 += seg[id].a00; a10 += seg[id].a10; a01 += seg[id].a01;
                id = seg[id].next;
            } while (id > 0 && ++guard < ncp);
            int keep = -1, base = 0;
            if (id != 0) atomicOr(&fi->overflow, 8); // border did not close: internal error
            if (id == 0 && fabs((double)a00) * 0.5 >= area_min) {
                keep = atomicAdd(&fi->n_kept, 1);
                base = atomicAdd(&fi->n_points, off);
                if (keep >= max_kept) { atomicOr(&fi->overflow, 2); keep = -1; }
                else {
                    int npts = off;
                    if (base + off > max_points) { atomicOr(&fi->overflow, 4); npts = 0; }
                    gofpid_blob *kb = kept + (size_t)z * max_kept + keep;
                    kb->first = ys * W + xs;
                    kb->x0 = bx0; kb->y0 = by0; kb->x1 = bx1; kb->y1 = by1;
                    kb->area_px = cr[CR_AREA];
                    kb->n_points = npts; kb->point_offset = base;
                    kb->a00 = a00; kb->a10 = a10; kb->a01 = a01;
                    if (npts == 0) keep = -1;
                }
            }
            sh_keep = keep; sh_off = base;
        }
        __syncthreads();

        // ---- pass 2: write the contour, one thread per linked segment -----------------------------
        if (sh_keep >= 0) {
            int32_t *dst = points + ((size_t)z * max_points + sh_off) * 2;
            for (int i = tid; i < ncp; i += TRACE_THREADS) {
                if (cp_off[i] < 0) continue;
                const int st = cp_state[i];
                Segment S;
                trace_segment<true>(bv, bx0 + ((st >> 3) & 4095), by0 + (st >> 15), st & 7, rmask, ht_keys,
                                    ht_vals, S, dst + 2 * (size_t)cp_off[i], max_steps);
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------
static int launch_labelling(ImageU8 mask, int32_t *labels, size_t LP, size_t labels_stream, int H,
                            int W, int n_streams, gofpid_frame_info *info, const FrameTick *tick,
                            cudaStream_t s)
{
    const dim3 tgrid((W + CCL_TW - 1) / CCL_TW, (H + CCL_TH - 1) / CCL_TH, n_streams);
    label_tiles_kernel<<<tgrid, dim3(16, CCL_TH), 0, s>>>(mask, labels, LP, labels_stream, H, W,
                                                          info, tick);
    GOFPID_CUDA_TRY(cudaGetLastError());
    merge_tiles_kernel<<<tgrid, CCL_TW + 2 * CCL_TH, 0, s>>>(mask, labels, LP, labels_stream, H, W);
    GOFPID_CUDA_TRY(cudaGetLastError());
    return GOFPID_OK;
}

int launch_blobs(const BlobBuffers &B, ImageU8 mask, int H, int W, int n_streams,
                 const FrameTick *tick, cudaStream_t s, int *n_launches)
{
    GOFPID_TRY(launch_labelling(mask, B.labels, B.labels_pitch, B.labels_stream, H, W, n_streams,
                                B.info, tick, s));
    const dim3 pgrid((W + 127) / 128, (H + 7) / 8, n_streams);
    flatten_kernel<<<pgrid, dim3(32, 8), 0, s>>>(mask, B.labels, B.labels_pitch, B.labels_stream,
                                                 B.slot, B.comp, B.info, B.max_blobs, H, W);
    GOFPID_CUDA_TRY(cudaGetLastError());
    stats_kernel<<<pgrid, dim3(32, 8), 0, s>>>(mask, B.labels, B.labels_pitch, B.labels_stream,
                                               B.slot, B.comp, B.max_blobs, H, W);
    GOFPID_CUDA_TRY(cudaGetLastError());
    GOFPID_CUDA_TRY(cudaFuncSetAttribute(trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)TRACE_SMEM_BYTES));
    const dim3 cgrid(n_streams >= 32 ? 8 : (n_streams >= 4 ? 32 : 128), n_streams);
    trace_kernel<<<cgrid, TRACE_THREADS, TRACE_SMEM_BYTES, s>>>(
        B.labels, B.labels_pitch, B.labels_stream, B.comp, B.info, B.kept, B.points, B.max_blobs,
        B.max_kept, B.max_points, B.area_min, H, W);
    GOFPID_CUDA_TRY(cudaGetLastError());
    if (n_launches) *n_launches = 5;
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
    // stage a padded copy of the mask (the tile kernel reads whole words) and provisional labels
    const size_t MP = round_up((size_t)W, 128), LP = round_up((size_t)W, 4);
    uint8_t *pm = nullptr;
    int32_t *pl = nullptr;
    gofpid_frame_info *fi = nullptr;
    GOFPID_CUDA_TRY(cudaMalloc(&pm, MP * H));
    GOFPID_CUDA_TRY(cudaMalloc(&pl, LP * H * sizeof(int32_t)));
    GOFPID_CUDA_TRY(cudaMalloc(&fi, sizeof(gofpid_frame_info)));
    int rc = GOFPID_OK;
    do {
        if (cudaMemcpy2DAsync(pm, MP, d_mask, mask_pitch, W, H, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
            set_error("mask staging copy failed"); rc = GOFPID_ERR_CUDA; break;
        }
        ImageU8 m{pm, MP, 0};
        rc = launch_labelling(m, pl, LP, 0, H, W, 1, fi, nullptr, s);
        if (rc) break;
        rc = launch_canonical_labels(pm, MP, pl, LP, d_labels, labels_pitch_elems, H, W, s);
        if (rc) break;
        if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("labelling failed: %s", cudaGetErrorString(cudaGetLastError())); rc = GOFPID_ERR_CUDA; }
    } while (0);
    cudaFree(pm); cudaFree(pl); cudaFree(fi);
    return rc;
}

} // namespace gofpid
