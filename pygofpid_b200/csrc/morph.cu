// morph.cu -- fused chain of grey-level erode / dilate steps (sm_100a).
//
// Replaces the morphology loop of GOFPID.detect (pipeline.py:473-475):
//     for d in self.mat_morph: mask = d['fun'](mask, d['kernel'])       fun in {cv.erode, cv.dilate}
// OpenCV semantics: out(y,x) = min / max of src over {(y+i-ay, x+j-ax) : K[i][j] != 0} with
// anchor (kh/2, kw/2); positions outside the image are ignored (they behave as the identity
// of the operation).  Masks are grey-level (0 / 127 / 255 with shadow detection on).
//
// The whole chain runs in ONE kernel: a CTA stages its output tile plus the accumulated halo
// of all steps in shared memory and ping-pongs between smem buffers, so the mask is read
// once and written once (2 B/px) however long the chain is.  Pixels are processed four at a
// time: a 32-bit word of bytes is split into its even and odd bytes as two u16x2 registers and
// combined with Blackwell's packed VIMNMX.U16x2 min/max.  Full rectangles run as a row pass
// and a column pass; any other structuring element is decomposed at plan time into the
// horizontal runs of its rows and evaluated as sliding windows over those runs.
#include "common.cuh"

#include <string.h>

#include <vector>

namespace gofpid {

constexpr int MORPH_THREADS = 256;
constexpr int MORPH_SLACK = 4; // bytes of slack left/right of the staged region

// ------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------
int morph_plan_create(const gofpid_morph_op *ops, int n_ops, MorphPlan *plan)
{
    memset(plan, 0, sizeof(*plan));
    if (n_ops < 0 || n_ops > GOFPID_MAX_MORPH_OPS) {
        set_error("mat_morph has %d steps; at most %d are supported", n_ops, GOFPID_MAX_MORPH_OPS);
        return GOFPID_ERR_UNSUPPORTED;
    }
    std::vector<MorphRun> runs;
    int anchored = 1; // every structuring element contains its anchor
    for (int i = 0; i < n_ops; i++) {
        const gofpid_morph_op &o = ops[i];
        if (o.op != GOFPID_MORPH_ERODE && o.op != GOFPID_MORPH_DILATE) {
            set_error("mat_morph[%d]: unknown operator %d", i, o.op);
            return GOFPID_ERR_INVALID;
        }
        if (!o.kernel || o.kh <= 0 || o.kw <= 0) {
            set_error("mat_morph[%d]: structuring element missing or empty", i);
            return GOFPID_ERR_INVALID;
        }
        if (o.kh > GOFPID_MAX_SE || o.kw > GOFPID_MAX_SE) {
            set_error("mat_morph[%d]: structuring element %dx%d exceeds the supported %d", i,
                      o.kh, o.kw, GOFPID_MAX_SE);
            return GOFPID_ERR_UNSUPPORTED;
        }
        MorphOpDev &d = plan->ops[i];
        d.op = o.op; d.kh = o.kh; d.kw = o.kw; d.ay = o.kh / 2; d.ax = o.kw / 2;
        int nz = 0;
        for (int k = 0; k < o.kh * o.kw; k++) nz += o.kernel[k] != 0;
        if (nz == 0) {
            set_error("mat_morph[%d]: structuring element has no non-zero element", i);
            return GOFPID_ERR_UNSUPPORTED;
        }
        d.is_rect = nz == o.kh * o.kw;
        if (!o.kernel[d.ay * o.kw + d.ax]) anchored = 0;
        d.run_begin = (int)runs.size();
        if (!d.is_rect) {
            for (int r = 0; r < o.kh; r++) {
                int c = 0;
                while (c < o.kw) {
                    if (!o.kernel[r * o.kw + c]) { c++; continue; }
                    int e = c;
                    while (e < o.kw && o.kernel[r * o.kw + e]) e++;
                    MorphRun run;
                    run.dy = (int8_t)(r - d.ay); run.dx = (int8_t)(c - d.ax);
                    run.len = (uint8_t)(e - c); run.pad = 0;
                    runs.push_back(run);
                    c = e;
                }
            }
        }
        d.run_end = (int)runs.size();
        plan->halo_t += d.ay; plan->halo_b += o.kh - 1 - d.ay;
        plan->halo_l += d.ax; plan->halo_r += o.kw - 1 - d.ax;
    }
    plan->n_ops = n_ops;
    plan->anchored = anchored;
    plan->n_runs = (int)runs.size();
    if (!runs.empty()) {
        GOFPID_CUDA_TRY(cudaMalloc(&plan->d_runs, runs.size() * sizeof(MorphRun)));
        GOFPID_CUDA_TRY(cudaMemcpy(plan->d_runs, runs.data(), runs.size() * sizeof(MorphRun),
                                   cudaMemcpyHostToDevice));
    }
    const int hmax = plan->halo_t > plan->halo_b ? plan->halo_t : plan->halo_b;
    plan->tile_w = 128;
    plan->tile_h = hmax > 8 ? 64 : 32;
    const int hl4 = (plan->halo_l + 3) & ~3, hr4 = (plan->halo_r + 3) & ~3;
    const int stride = MORPH_SLACK + hl4 + plan->tile_w + hr4 + MORPH_SLACK;
    const int rows = plan->tile_h + plan->halo_t + plan->halo_b;
    plan->smem_bytes = (size_t)6 * rows * stride; // three images, each split in two u16x2 planes
    return GOFPID_OK;
}

void morph_plan_destroy(MorphPlan *plan)
{
    if (plan->d_runs) cudaFree(plan->d_runs);
    plan->d_runs = nullptr;
}

// ------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------
// Shared-memory images are kept SPLIT: a word of four pixels p0..p3 is stored as
// E = (p0, p2) and O = (p1, p3), two u16 lanes each, in two planes.  Packed VIMNMX (u16x2,
// three inputs) then works on loaded words directly, and a horizontal shift by one pixel is a
// swap of planes plus a 16-bit funnel shift.
__device__ __forceinline__ uint32_t split_e(uint32_t w) { return __byte_perm(w, 0, 0x4240); }
__device__ __forceinline__ uint32_t split_o(uint32_t w) { return __byte_perm(w, 0, 0x4341); }
__device__ __forceinline__ uint32_t join_eo(uint32_t e, uint32_t o) { return e | (o << 8); }

template <bool DIL> __device__ __forceinline__ uint32_t mm2(uint32_t a, uint32_t b)
{
    return DIL ? __vmaxu2(a, b) : __vminu2(a, b);
}
template <bool DIL> __device__ __forceinline__ uint32_t mm3(uint32_t a, uint32_t b, uint32_t c)
{
    return DIL ? __vimax3_u16x2(a, b, c) : __vimin3_u16x2(a, b, c);
}
// 32 bits of a u16 stream starting at u16 index p (words lo = plane[p>>1], hi = plane[(p>>1)+1])
__device__ __forceinline__ uint32_t win16(uint32_t lo, uint32_t hi, int p)
{
    return __funnelshift_r(lo, hi, (p & 1) << 4);
}

struct Planes {
    uint32_t *e, *o; // [rows][nwords]
};

// Row pass of a K-wide full rectangle (anchor K/2) for output word wc, taps known at compile time.
//   E_out(t) = even t: Ewin(t/2)      odd t: Owin((t-1)/2)
//   O_out(t) = even t: Owin(t/2)      odd t: Ewin((t+1)/2)      (win index relative to 2*wc)
template <bool DIL, int K>
__device__ __forceinline__ void row_rect(const uint32_t *er, const uint32_t *orow, int wc, int nwords,
                                         uint32_t &oute, uint32_t &outo)
{
    constexpr int R = K / 2;
    static_assert(K % 2 == 1 && R <= 3, "specialised for 3, 5, 7");
    const int wm = max(wc - 1, 0), wp = min(wc + 1, nwords - 1);
    const uint32_t e[3] = {er[wm], er[wc], er[wp]}, o[3] = {orow[wm], orow[wc], orow[wp]};
    // windows at u16 offsets -2..+2 relative to the word: index i -> array slot i + 2
    uint32_t ew[5], ow[5];
    ew[0] = e[0]; ew[2] = e[1]; ew[4] = e[2];
    ow[0] = o[0]; ow[2] = o[1]; ow[4] = o[2];
    ew[1] = __funnelshift_r(e[0], e[1], 16); ew[3] = __funnelshift_r(e[1], e[2], 16);
    ow[1] = __funnelshift_r(o[0], o[1], 16); ow[3] = __funnelshift_r(o[1], o[2], 16);
    uint32_t ae = DIL ? 0u : 0x00ff00ffu, ao = ae;
#pragma unroll
    for (int t = -R; t <= R; t++) {
        if ((t & 1) == 0) {
            ae = mm2<DIL>(ae, ew[t / 2 + 2]);
            ao = mm2<DIL>(ao, ow[t / 2 + 2]);
        } else {
            ae = mm2<DIL>(ae, ow[(t - 1) / 2 + 2]);
            ao = mm2<DIL>(ao, ew[(t + 1) / 2 + 2]);
        }
    }
    oute = ae; outo = ao;
}

// Generic horizontal window: min/max over `len` taps starting at pixel offset t0 (relative to
// the first pixel of word wc), any length / anchor.
template <bool DIL>
__device__ __forceinline__ void window_split(uint32_t &ae, uint32_t &ao, const uint32_t *er,
                                             const uint32_t *orow, int wc, int t0, int len, int nwords)
{
    int p = 2 * wc + (t0 >> 1); // u16 stream index (floor division also for negative t0)
    int parity = t0 & 1;
    int wi = p >> 1;
    auto ld = [&](const uint32_t *pl, int i) { return pl[min(max(i, 0), nwords - 1)]; };
    uint32_t e_lo = ld(er, wi), e_hi = ld(er, wi + 1), o_lo = ld(orow, wi), o_hi = ld(orow, wi + 1);
    for (int j = 0; j < len; j++) {
        if (parity == 0) {
            ae = mm2<DIL>(ae, win16(e_lo, e_hi, p));
            ao = mm2<DIL>(ao, win16(o_lo, o_hi, p));
            parity = 1;
        } else {
            ae = mm2<DIL>(ae, win16(o_lo, o_hi, p));
            p++;
            if ((p >> 1) != wi) {
                wi = p >> 1;
                e_lo = e_hi; o_lo = o_hi;
                e_hi = ld(er, wi + 1); o_hi = ld(orow, wi + 1);
            }
            ao = mm2<DIL>(ao, win16(e_lo, e_hi, p));
            parity = 0;
        }
    }
}

struct OpGeom {
    int r0, r1, w0, nw;     // valid output rows [r0, r1) and word columns [w0, w0 + nw)
    int gy0, gx0, H, W;     // image coordinates of buffer (0, 0)
    int nwords, rows;
};

template <bool DIL>
__device__ __forceinline__ void run_op(const MorphOpDev &op, const MorphRun *__restrict__ runs,
                                       Planes A, Planes T, Planes B, const OpGeom &G,
                                       uint32_t ident_next)
{
    const int tid = threadIdx.x;
    const int nwords = G.nwords;
    const uint32_t init = DIL ? 0u : 0x00ff00ffu;
    // one division per step and thread: column c and first row r of this thread
    const int c = tid % G.nw, rofs = tid / G.nw, rstep = MORPH_THREADS / G.nw;
    const int wc = G.w0 + c;
    const bool active = rofs < rstep;
    // in-image lanes of this word column (E = pixels 0, 2; O = pixels 1, 3)
    const int gx = G.gx0 + wc * 4;
    const uint32_t inE = ((gx >= 0 && gx < G.W) ? 0x0000ffffu : 0u) | ((gx + 2 >= 0 && gx + 2 < G.W) ? 0xffff0000u : 0u);
    const uint32_t inO = ((gx + 1 >= 0 && gx + 1 < G.W) ? 0x0000ffffu : 0u) | ((gx + 3 >= 0 && gx + 3 < G.W) ? 0xffff0000u : 0u);
    const uint32_t idn = ident_next & 0x00ff00ffu;
    const bool fast_rect = op.is_rect && op.kw == op.kh && (op.kw == 3 || op.kw == 5 || op.kw == 7);

    if (op.is_rect) {
        // row pass over the rows the column pass will read
        const int rr0 = G.r0 - op.ay, rr1 = G.r1 + (op.kh - 1 - op.ay);
        if (active) {
            for (int r = rr0 + rofs; r < rr1; r += rstep) {
                const uint32_t *er = A.e + (size_t)r * nwords, *orow = A.o + (size_t)r * nwords;
                uint32_t ae = init, ao = init;
                if (fast_rect && op.kw == 5) row_rect<DIL, 5>(er, orow, wc, nwords, ae, ao);
                else if (fast_rect && op.kw == 3) row_rect<DIL, 3>(er, orow, wc, nwords, ae, ao);
                else if (fast_rect) row_rect<DIL, 7>(er, orow, wc, nwords, ae, ao);
                else window_split<DIL>(ae, ao, er, orow, wc, -op.ax, op.kw, nwords);
                T.e[(size_t)r * nwords + wc] = ae;
                T.o[(size_t)r * nwords + wc] = ao;
            }
        }
        __syncthreads();
    }
    if (active) {
        for (int r = G.r0 + rofs; r < G.r1; r += rstep) {
            uint32_t ae = init, ao = init;
            if (op.is_rect) {
                const uint32_t *te = T.e + (size_t)(r - op.ay) * nwords + wc;
                const uint32_t *to = T.o + (size_t)(r - op.ay) * nwords + wc;
                if (fast_rect && op.kh == 5) {
                    ae = mm3<DIL>(mm3<DIL>(te[0], te[nwords], te[2 * nwords]), te[3 * nwords], te[4 * nwords]);
                    ao = mm3<DIL>(mm3<DIL>(to[0], to[nwords], to[2 * nwords]), to[3 * nwords], to[4 * nwords]);
                } else if (fast_rect && op.kh == 3) {
                    ae = mm3<DIL>(te[0], te[nwords], te[2 * nwords]);
                    ao = mm3<DIL>(to[0], to[nwords], to[2 * nwords]);
                } else {
                    for (int k = 0; k < op.kh; k++) {
                        ae = mm2<DIL>(ae, te[(size_t)k * nwords]);
                        ao = mm2<DIL>(ao, to[(size_t)k * nwords]);
                    }
                }
            } else {
                for (int q = op.run_begin; q < op.run_end; q++) {
                    const MorphRun run = runs[q];
                    window_split<DIL>(ae, ao, A.e + (size_t)(r + run.dy) * nwords,
                                      A.o + (size_t)(r + run.dy) * nwords, wc, run.dx, run.len, nwords);
                }
            }
            // cells outside the image take the identity of the NEXT step (OpenCV ignores them)
            const int gy = G.gy0 + r;
            const bool rin = gy >= 0 && gy < G.H;
            const uint32_t me = rin ? inE : 0u, mo = rin ? inO : 0u;
            B.e[(size_t)r * nwords + wc] = (ae & me) | (idn & ~me);
            B.o[(size_t)r * nwords + wc] = (ao & mo) | (idn & ~mo);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(MORPH_THREADS)
morph_chain_kernel(MorphPlan plan, ImageU8 src, ImageU8 dst, int H, int W)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int hl4 = (plan.halo_l + 3) & ~3, hr4 = (plan.halo_r + 3) & ~3;
    const int stride = MORPH_SLACK + hl4 + plan.tile_w + hr4 + MORPH_SLACK; // pixels per staged row
    const int rows = plan.tile_h + plan.halo_t + plan.halo_b;
    const int nwords = stride >> 2;
    const size_t plane = (size_t)rows * nwords;
    uint32_t *base = reinterpret_cast<uint32_t *>(smem);
    Planes A{base, base + plane}, T{base + 2 * plane, base + 3 * plane}, B{base + 4 * plane, base + 5 * plane};

    const int x0 = blockIdx.x * plan.tile_w, y0 = blockIdx.y * plan.tile_h;
    const int gx0 = x0 - hl4 - MORPH_SLACK; // image x of buffer column 0 (multiple of 4)
    const int gy0 = y0 - plan.halo_t;
    const uint8_t *sbase = src.base + (size_t)blockIdx.z * src.stream_stride;
    uint8_t *dbase = dst.base + (size_t)blockIdx.z * dst.stream_stride;
    const bool aligned = ((reinterpret_cast<uintptr_t>(sbase) | src.pitch) & 3) == 0;
    const bool daligned = ((reinterpret_cast<uintptr_t>(dbase) | dst.pitch) & 3) == 0;
    const uint32_t ident0 = plan.n_ops && plan.ops[0].op == GOFPID_MORPH_ERODE ? 0xffffffffu : 0u;

    // ---- stage the tile + halo; out-of-image cells get the first step's identity --------
    bool any_set = false, any_clear = false; // over in-image bytes: some != 0, some != 255
    for (int i = tid; i < rows * nwords; i += MORPH_THREADS) {
        const int r = i / nwords, wc = i - r * nwords;
        const int gy = gy0 + r, gx = gx0 + wc * 4;
        uint32_t v = ident0;
        if (gy >= 0 && gy < H && gx + 3 >= 0 && gx < W) {
            const uint8_t *row = sbase + (size_t)gy * src.pitch;
            if (aligned && gx >= 0 && gx + 3 < W) {
                v = *reinterpret_cast<const uint32_t *>(row + gx);
                any_set |= v != 0u; any_clear |= v != 0xffffffffu;
            } else {
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (gx + b >= 0 && gx + b < W) {
                        const uint32_t px = row[gx + b];
                        v = (v & ~(0xffu << (8 * b))) | (px << (8 * b));
                        any_set |= px != 0u; any_clear |= px != 255u;
                    }
            }
        }
        A.e[i] = split_e(v);
        A.o[i] = split_o(v);
    }
    // (__syncthreads_or returns a truth value, not the OR of the arguments: one call per flag)
    const int uniform = (__syncthreads_or(any_set) ? 1 : 0) | (__syncthreads_or(any_clear) ? 2 : 0);

    const int tw4 = plan.tile_w >> 2;
    const int cbase = (MORPH_SLACK + hl4) >> 2; // buffer word column of x0
    if (plan.n_ops > 0 && uniform != 3 && plan.anchored) {
        // every in-image pixel the tile can see is 0 (or every one is 255): min / max over a
        // structuring element that contains its anchor reproduces that value, so the whole chain
        // is the identity here
        const uint32_t v = (uniform & 1) ? 0xffffffffu : 0u;
        for (int i = tid; i < plan.tile_h * tw4; i += MORPH_THREADS) {
            const int r = i / tw4, wc = i - r * tw4;
            const int gy = y0 + r, gx = x0 + wc * 4;
            if (gy >= H || gx >= W) continue;
            uint8_t *row = dbase + (size_t)gy * dst.pitch;
            if (daligned && gx + 3 < W) *reinterpret_cast<uint32_t *>(row + gx) = v;
            else {
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (gx + b < W) row[gx + b] = (uint8_t)v;
            }
        }
        return;
    }

    // ---- run the chain, shrinking the valid region by each step's extent -----------------
    OpGeom G;
    G.gy0 = gy0; G.gx0 = gx0; G.H = H; G.W = W; G.nwords = nwords; G.rows = rows;
    int r0 = 0, r1 = rows, c0 = MORPH_SLACK + hl4 - plan.halo_l,
        c1 = MORPH_SLACK + hl4 + plan.tile_w + plan.halo_r;
    for (int k = 0; k < plan.n_ops; k++) {
        const MorphOpDev &op = plan.ops[k];
        r0 += op.ay; r1 -= op.kh - 1 - op.ay;
        c0 += op.ax; c1 -= op.kw - 1 - op.ax;
        G.r0 = r0; G.r1 = r1; G.w0 = c0 >> 2; G.nw = ((c1 + 3) >> 2) - G.w0;
        const uint32_t ident_next =
            (k + 1 < plan.n_ops && plan.ops[k + 1].op == GOFPID_MORPH_ERODE) ? 0xffffffffu : 0u;
        if (op.op == GOFPID_MORPH_DILATE) run_op<true>(op, plan.d_runs, A, T, B, G, ident_next);
        else run_op<false>(op, plan.d_runs, A, T, B, G, ident_next);
        const Planes t = A; A = B; B = t;
    }

    // ---- write the tile -----------------------------------------------------------------
    for (int i = tid; i < plan.tile_h * tw4; i += MORPH_THREADS) {
        const int r = i / tw4, wc = i - r * tw4;
        const int gy = y0 + r, gx = x0 + wc * 4;
        if (gy >= H || gx >= W) continue;
        const size_t si = (size_t)(r + plan.halo_t) * nwords + cbase + wc;
        const uint32_t v = join_eo(A.e[si], A.o[si]);
        uint8_t *row = dbase + (size_t)gy * dst.pitch;
        if (daligned && gx + 3 < W) {
            *reinterpret_cast<uint32_t *>(row + gx) = v;
        } else {
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (gx + b < W) row[gx + b] = (uint8_t)(v >> (8 * b));
        }
    }
}


// =====================================================================================
// bit-plane path
// =====================================================================================
// Inside the pipeline the mask that reaches the chain only holds the detector's alphabet: 0 and
// at most three non-zero levels (MOG2 / KNN: shadow value and 255; frame differencing: 255, or
// 84 / 85 for colour frames).  Flat erosion / dilation commute with thresholding, so the chain
// is evaluated on one-bit threshold planes  P_i = (mask >= level_i)  with 32 pixels per word:
// a horizontal run of the structuring element is a handful of funnel shifts and ANDs (window
// doubling on a 64-bit register), the vertical pass of a rectangle is one AND per row, and the
// grey-level result is rebuilt from the planes on the way out.  A CTA owns 256 x tile_h pixels,
// stages tile + accumulated halo (whole halo WORDS left and right) as bits, and also emits the
// foreground bit image the blob stage works on.
#ifndef GOFPID_MB_THREADS
#define GOFPID_MB_THREADS 256
#endif
#ifndef GOFPID_MB_TWW
#define GOFPID_MB_TWW 8
#endif
#ifndef GOFPID_MB_R
#define GOFPID_MB_R 4
#endif
constexpr int MB_THREADS = GOFPID_MB_THREADS;
constexpr int MB_TWW = GOFPID_MB_TWW; // tile width in 32-pixel words
constexpr int MB_R_PAD = 4; // spare rows for the last, partial block of output rows (>= MB_R - 1)

void morph_plan_set_levels(MorphPlan *plan, const int *levels, int n_levels)
{
    plan->n_levels = 0;
    if (plan->n_ops <= 0 || n_levels < 1 || n_levels > 3) return;
    for (int i = 0; i < n_levels; i++) {
        if (levels[i] < 1 || levels[i] > 255 || (i > 0 && levels[i] <= levels[i - 1])) return;
        plan->levels[i] = levels[i];
    }
    // all taps outside the image (possible only without the anchor): OpenCV yields 255 / 0, which
    // the planes reproduce only when the top level is 255
    if (!plan->anchored && levels[n_levels - 1] != 255) return;
    plan->bit_hw_l = (plan->halo_l + 31) / 32;
    plan->bit_hw_r = (plan->halo_r + 31) / 32;
    const int hmax = plan->halo_t > plan->halo_b ? plan->halo_t : plan->halo_b;
    plan->bit_tile_h = hmax > 8 ? 64 : 32;
    const size_t nw = plan->bit_hw_l + MB_TWW + plan->bit_hw_r;
    const size_t rows = plan->bit_tile_h + plan->halo_t + plan->halo_b;
    int pad = 0;
    for (int i = 0; i < plan->n_ops; i++) {
        const MorphOpDev &d = plan->ops[i];
        pad = d.ay > pad ? d.ay : pad;
        pad = d.kh - 1 - d.ay > pad ? d.kh - 1 - d.ay : pad;
    }
    plan->bit_pad = pad + MB_R_PAD;
    plan->bit_smem_bytes = 2 * (size_t)n_levels * (rows + 2 * plan->bit_pad) * (nw + 2) * sizeof(uint32_t);
    if (plan->bit_smem_bytes > 160 * 1024) return; // enormous halo: stay on the byte path
    // A threshold plane is a single bit of the byte when some bit is set in every level at or above
    // the threshold and clear in every level below it (0 / 127 / 255: bit 0 and bit 7); otherwise the
    // kernel compares bytes.
    for (int i = 0; i < n_levels; i++) {
        int mask = 0xff;
        for (int j = 0; j < n_levels; j++) mask &= j >= i ? levels[j] : ~levels[j];
        plan->bit_sel[i] = -1;
        for (int b = 0; b < 8; b++)
            if (mask & (1 << b)) { plan->bit_sel[i] = b; break; }
    }
    plan->n_levels = n_levels;
}

// 4 bits -> 4 bytes of 0x00 / 0xff (bit i -> byte i)
__device__ __forceinline__ uint32_t spread4(uint32_t n)
{
    return (((n & 15u) * 0x00204081u) & 0x01010101u) * 0xffu;
}
// 4 bytes -> 4 bits: byte i >= t  ->  bit i   (t4 = t replicated in the four bytes)
__device__ __forceinline__ uint32_t ge4(uint32_t w, uint32_t t4)
{
    return ((__vcmpgeu4(w, t4) & 0x80808080u) * 0x00204081u) >> 28;
}

// 4 bytes -> 4 bits: bit k of byte i -> bit i
__device__ __forceinline__ uint32_t bit4(uint32_t w, int k)
{
    return (((w >> k) & 0x01010101u) * 0x10204080u) >> 28;
}

// AND (erode) / OR (dilate) of the taps dx0 .. dx0+len-1 of a bit row, for the 32 pixels of word b
// (a, c: the words left and right of it).  -31 <= dx0, dx0 + len <= 32, len >= 1.
template <bool DIL>
__device__ __forceinline__ uint32_t run_window(uint32_t a, uint32_t b, uint32_t c, int dx0, int len)
{
    if (DIL) { a = ~a; b = ~b; c = ~c; } // OR = complement of AND of complements
    uint32_t lo, hi;
    if (dx0 < 0) { lo = __funnelshift_r(a, b, 32 + dx0); hi = __funnelshift_r(b, c, 32 + dx0); }
    else { lo = __funnelshift_r(b, c, dx0); hi = c >> dx0; }
    unsigned long long u = ((unsigned long long)hi << 32) | lo; // bit j = pixel 32w + dx0 + j
    int cov = 1;
    while (2 * cov <= len) { u &= u >> cov; cov *= 2; }
    if (cov < len) u &= u >> (len - cov);
    const uint32_t r = (uint32_t)u;
    return DIL ? ~r : r;
}

struct BitGeom {
    int NW, rows, plane; // words per row, rows, words per plane
    int gx0, gy0, H, W;  // image coordinates of buffer bit (0, 0)
};

// bits of word column w that lie inside the image
__device__ __forceinline__ uint32_t in_image_bits(const BitGeom &G, int w)
{
    const int gx = G.gx0 + 32 * w;
    if (gx + 31 < 0 || gx >= G.W) return 0u;
    uint32_t m = 0xffffffffu;
    if (gx < 0) m &= 0xffffffffu << (-gx);
    if (gx + 32 > G.W) m &= 0xffffffffu >> (gx + 32 - G.W);
    return m;
}

// Horizontal window of a K-wide rectangle (anchor K/2), taps known at compile time.
template <bool DIL, int K>
__device__ __forceinline__ uint32_t row_window_k(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t acc = b;
#pragma unroll
    for (int d = 1; d <= K / 2; d++) {
        const uint32_t l = __funnelshift_r(a, b, 32 - d); // pixel x - d
        const uint32_t r = __funnelshift_r(b, c, d);      // pixel x + d
        acc = DIL ? (acc | l | r) : (acc & l & r);
    }
    return acc;
}

constexpr int MB_R = GOFPID_MB_R;    // output rows per work item (a thread owns MB_R rows of one word column)
constexpr int MB_MAXI = 2; // work items per thread decoded up front

// Shared-memory images are padded: one spare word left and right of every row and `pad` spare
// rows above and below, so the windows are read without clamping.  Whatever the spare cells
// hold only reaches bits whose support left the staged region, and those lie in the part of
// the halo that later steps no longer read.
struct BitItem {
    int base;      // offset of (row r0, word w) in a padded plane, plus the plane offset
    int r0, w;
    uint32_t inm;  // in-image bits of word column w
};

// One step for one work item: output rows r0 .. r0+MB_R-1 of word column w.
// Returns the OR of the in-image bits it wrote (0: the item's output is empty).
template <bool DIL, int K>
__device__ __forceinline__ uint32_t op_item(const MorphOpDev &op, const MorphRun *__restrict__ runs,
                                            const uint32_t *A, uint32_t *B, const BitItem &it, const BitGeom &G,
                                            int NWP, uint32_t ident_next, bool anchored)
{
    const uint32_t ident = DIL ? 0u : 0xffffffffu;
    const uint32_t *Ap = A + it.base;
    uint32_t acc[MB_R];
#pragma unroll
    for (int q = 0; q < MB_R; q++) acc[q] = ident;
    bool skip = false;
    if (!DIL && anchored) {
        // an erosion whose structuring element holds its anchor only removes pixels: where the item's own
        // words are empty (most of a detector's mask, and its salt noise dies here) so is its output
        uint32_t own = 0u;
#pragma unroll
        for (int q = 0; q < MB_R; q++) own |= Ap[q * NWP];
        skip = own == 0u;
    }
    if (skip) {
#pragma unroll
        for (int q = 0; q < MB_R; q++) acc[q] = 0u;
    } else if (K > 0) {
        // K x K rectangle: K + MB_R - 1 row windows, each output row is the AND / OR of K of them
        constexpr int KK = K > 0 ? K : 1;
        uint32_t h[MB_R + KK - 1];
#pragma unroll
        for (int i = 0; i < MB_R + KK - 1; i++) {
            const uint32_t *row = Ap + (i - KK / 2) * NWP;
            h[i] = row_window_k<DIL, KK>(row[-1], row[0], row[1]);
        }
#pragma unroll
        for (int q = 0; q < MB_R; q++)
#pragma unroll
            for (int k = 0; k < KK; k++) acc[q] = DIL ? (acc[q] | h[q + k]) : (acc[q] & h[q + k]);
    } else if (op.is_rect) {
        const int top = op.ay, bot = op.kh - 1 - op.ay;
        for (int rr = -top; rr < MB_R + bot; rr++) {
            const uint32_t *row = Ap + rr * NWP;
            const uint32_t hv = run_window<DIL>(row[-1], row[0], row[1], -op.ax, op.kw);
#pragma unroll
            for (int q = 0; q < MB_R; q++)
                if (rr >= q - top && rr <= q + bot) acc[q] = DIL ? (acc[q] | hv) : (acc[q] & hv);
        }
    } else {
        for (int qi = op.run_begin; qi < op.run_end; qi++) {
            const MorphRun run = runs[qi];
#pragma unroll
            for (int q = 0; q < MB_R; q++) {
                const uint32_t *row = Ap + (q + run.dy) * NWP;
                const uint32_t v = run_window<DIL>(row[-1], row[0], row[1], run.dx, run.len);
                acc[q] = DIL ? (acc[q] | v) : (acc[q] & v);
            }
        }
    }
    // cells outside the image take the identity of the NEXT step (OpenCV ignores them)
    uint32_t *Bp = B + it.base;
    uint32_t nz = 0u;
#pragma unroll
    for (int q = 0; q < MB_R; q++) {
        const int r = it.r0 + q;
        if (r >= G.rows) break;
        const int gy = G.gy0 + r;
        const uint32_t m = (gy >= 0 && gy < G.H) ? it.inm : 0u;
        nz |= acc[q] & m;
        Bp[q * NWP] = (acc[q] & m) | (ident_next & ~m);
    }
    return nz;
}

template <bool DIL>
__device__ __forceinline__ uint32_t op_item_dispatch(const MorphOpDev &op, const MorphRun *__restrict__ runs,
                                                     const uint32_t *A, uint32_t *B, const BitItem &it,
                                                     const BitGeom &G, int NWP, uint32_t ident_next, bool anchored)
{
    const bool sq = op.is_rect && op.kw == op.kh && (op.kw & 1);
    if (sq && op.kw == 5) return op_item<DIL, 5>(op, runs, A, B, it, G, NWP, ident_next, anchored);
    if (sq && op.kw == 3) return op_item<DIL, 3>(op, runs, A, B, it, G, NWP, ident_next, anchored);
    return op_item<DIL, 0>(op, runs, A, B, it, G, NWP, ident_next, anchored);
}

__global__ void __launch_bounds__(MB_THREADS)
morph_bits_kernel(MorphPlan plan, ImageU8 src, ImageU8 dst, uint32_t *bitimg, int bit_pitch,
                  size_t bit_stream, int H, int W)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int nl = plan.n_levels;
    BitGeom G;
    G.NW = plan.bit_hw_l + MB_TWW + plan.bit_hw_r;
    G.rows = plan.bit_tile_h + plan.halo_t + plan.halo_b;
    const int NWP = G.NW + 2, pad = plan.bit_pad;
    G.plane = (G.rows + 2 * pad) * NWP;                   // words per padded plane
    const int org = pad * NWP + 1;                        // offset of cell (0, 0) inside a plane
    const int x0 = blockIdx.x * (MB_TWW * 32), y0 = blockIdx.y * plan.bit_tile_h;
    G.gx0 = x0 - 32 * plan.bit_hw_l; G.gy0 = y0 - plan.halo_t; G.H = H; G.W = W;
    uint32_t *A = reinterpret_cast<uint32_t *>(smem);
    uint32_t *B = A + (size_t)nl * G.plane;
    const uint8_t *sbase = src.base + (size_t)blockIdx.z * src.stream_stride;
    uint8_t *dbase = dst.base + (size_t)blockIdx.z * dst.stream_stride;
    const bool aligned = ((reinterpret_cast<uintptr_t>(sbase) | src.pitch) & 15) == 0;
    const bool daligned = ((reinterpret_cast<uintptr_t>(dbase) | dst.pitch) & 15) == 0;
    const uint32_t ident0 = plan.ops[0].op == GOFPID_MORPH_ERODE ? 0xffffffffu : 0u;
    uint32_t lv4[3];
#pragma unroll
    for (int l = 0; l < 3; l++) lv4[l] = (uint32_t)(plan.levels[l] & 255) * 0x01010101u;

    // ---- stage tile + halo as threshold bits; out-of-image cells get the first step's identity
    bool any = false;
    const int ncell = G.rows * G.NW;
    for (int i = tid; i < ncell; i += MB_THREADS) {
        const int r = i / G.NW, w = i - r * G.NW;
        const int gy = G.gy0 + r, gx = G.gx0 + 32 * w;
        uint32_t bits[3] = {0u, 0u, 0u};
        const uint32_t m = (gy >= 0 && gy < H) ? in_image_bits(G, w) : 0u;
        if (m) {
            const uint8_t *row = sbase + (size_t)gy * src.pitch;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t v[4] = {0u, 0u, 0u, 0u};
                const int hx = gx + 16 * h;
                if (aligned && hx >= 0 && (size_t)hx + 16 <= src.pitch) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(row + hx);
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                } else {
                    for (int b = 0; b < 16; b++)
                        if (hx + b >= 0 && hx + b < W) v[b >> 2] |= (uint32_t)row[hx + b] << (8 * (b & 3));
                }
                // (most of a foreground mask is background: sixteen zero bytes are zero bits in every plane)
                if ((v[0] | v[1] | v[2] | v[3]) == 0u) continue;
#pragma unroll
                for (int l = 0; l < 3; l++)
                    if (l < nl) {
                        const int k = plan.bit_sel[l];
                        uint32_t b16;
                        if (k >= 0) b16 = bit4(v[0], k) | (bit4(v[1], k) << 4) | (bit4(v[2], k) << 8) | (bit4(v[3], k) << 12);
                        else b16 = ge4(v[0], lv4[l]) | (ge4(v[1], lv4[l]) << 4) | (ge4(v[2], lv4[l]) << 8) | (ge4(v[3], lv4[l]) << 12);
                        bits[l] |= b16 << (16 * h);
                    }
            }
        }
        const int cell = org + r * NWP + w;
#pragma unroll
        for (int l = 0; l < 3; l++)
            if (l < nl) A[(size_t)l * G.plane + cell] = (bits[l] & m) | (ident0 & ~m);
        any |= (bits[0] & m) != 0u;
    }
    bool nonempty = __syncthreads_or(any) || !plan.anchored;

    const int tile_h = plan.bit_tile_h;
    if (nonempty) {
        // the thread's work items, decoded once: (plane, word column, block of MB_R rows)
        const int nrb = (G.rows + MB_R - 1) / MB_R;
        const int nitems = nl * G.NW * nrb;
        BitItem items[MB_MAXI];
        int n_mine = 0;
#pragma unroll
        for (int q = 0; q < MB_MAXI; q++) {
            const int item = tid + q * MB_THREADS;
            items[q].base = 0; items[q].r0 = 0; items[q].w = 0; items[q].inm = 0u;
            if (item < nitems) {
                const int w = item % G.NW, t = item / G.NW;
                const int rb = t % nrb, l = t / nrb;
                items[q].w = w; items[q].r0 = rb * MB_R;
                items[q].base = l * G.plane + org + rb * MB_R * NWP + w;
                items[q].inm = in_image_bits(G, w);
                n_mine = q + 1;
            }
        }
        const bool small = nitems <= MB_MAXI * MB_THREADS;
        for (int k = 0; k < plan.n_ops; k++) {
            const MorphOpDev &op = plan.ops[k];
            const uint32_t ident_next =
                (k + 1 < plan.n_ops && plan.ops[k + 1].op == GOFPID_MORPH_ERODE) ? 0xffffffffu : 0u;
            const bool dil = op.op == GOFPID_MORPH_DILATE;
            const bool anch = plan.anchored != 0;
            uint32_t nz = 0u;
            if (small) {
#pragma unroll
                for (int q = 0; q < MB_MAXI; q++)
                    if (q < n_mine) {
                        if (dil) nz |= op_item_dispatch<true>(op, plan.d_runs, A, B, items[q], G, NWP, ident_next, anch);
                        else nz |= op_item_dispatch<false>(op, plan.d_runs, A, B, items[q], G, NWP, ident_next, anch);
                    }
            } else {
                for (int item = tid; item < nitems; item += MB_THREADS) {
                    BitItem it;
                    const int w = item % G.NW, t = item / G.NW;
                    const int rb = t % nrb, l = t / nrb;
                    it.w = w; it.r0 = rb * MB_R; it.base = l * G.plane + org + rb * MB_R * NWP + w;
                    it.inm = in_image_bits(G, w);
                    if (dil) nz |= op_item_dispatch<true>(op, plan.d_runs, A, B, it, G, NWP, ident_next, anch);
                    else nz |= op_item_dispatch<false>(op, plan.d_runs, A, B, it, G, NWP, ident_next, anch);
                }
            }
            // (the barrier between two steps also tells whether anything is left: the salt noise of a
            // detector's mask is gone after the first erosion, and an empty tile stays empty)
            const bool left = __syncthreads_or(nz != 0u) != 0;
            uint32_t *t = A; A = B; B = t;
            if (!left && anch) { nonempty = false; break; }
        }
    }
    // (an empty tile -- no foreground anywhere the chain can see -- stays empty through a chain of
    // structuring elements that contain their anchors)

    // ---- grey levels back from the planes: 16 pixels (one uint4) per thread and step -----------
    for (int i = tid; i < tile_h * MB_TWW * 2; i += MB_THREADS) {
        const int r = i / (MB_TWW * 2), hw = i - r * (MB_TWW * 2);
        const int w = hw >> 1, h = hw & 1;
        const int gy = y0 + r, gx = x0 + 32 * w + 16 * h;
        if (gy >= H || gx >= W) continue;
        const size_t si = (size_t)org + (size_t)(r + plan.halo_t) * NWP + plan.bit_hw_l + w;
        uint32_t p[3] = {0u, 0u, 0u};
        if (nonempty) {
#pragma unroll
            for (int l = 0; l < 3; l++)
                if (l < nl) p[l] = (A[(size_t)l * G.plane + si] >> (16 * h)) & 0xffffu;
        }
        uint32_t o[4] = {0u, 0u, 0u, 0u};
        if (p[0] | p[1] | p[2]) { // (sixteen background pixels are sixteen zero bytes: nothing to spread)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t v = spread4(p[0] >> (4 * q)) & lv4[0];
#pragma unroll
                for (int l = 1; l < 3; l++)
                    if (l < nl) {
                        const uint32_t ml = spread4(p[l] >> (4 * q));
                        v = (v & ~ml) | (ml & lv4[l]);
                    }
                o[q] = v;
            }
        }
        uint8_t *row = dbase + (size_t)gy * dst.pitch;
        if (daligned && (size_t)gx + 16 <= dst.pitch) {
            *reinterpret_cast<uint4 *>(row + gx) = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
            for (int b = 0; b < 16; b++)
                if (gx + b < W) row[gx + b] = (uint8_t)(o[b >> 2] >> (8 * (b & 3)));
        }
        if (bitimg && h == 0) {
            const uint32_t fg = nonempty ? A[si] : 0u; // plane 0 = (mask != 0)
            bitimg[(size_t)blockIdx.z * bit_stream + (size_t)(gy + 1) * bit_pitch + 1 + (gx >> 5)] = fg;
        }
    }
}

int launch_morph(const MorphPlan &plan, ImageU8 src, ImageU8 dst, int H, int W, int n_streams,
                 cudaStream_t s, uint32_t *bitimg, int bit_pitch, size_t bit_stream)
{
    if (plan.n_levels > 0) {
        if (plan.bit_smem_bytes > 48 * 1024)
            GOFPID_CUDA_TRY(cudaFuncSetAttribute(morph_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)plan.bit_smem_bytes));
        const dim3 grid((W + MB_TWW * 32 - 1) / (MB_TWW * 32), (H + plan.bit_tile_h - 1) / plan.bit_tile_h,
                        n_streams);
        morph_bits_kernel<<<grid, MB_THREADS, plan.bit_smem_bytes, s>>>(plan, src, dst, bitimg, bit_pitch,
                                                                        bit_stream, H, W);
        GOFPID_CUDA_TRY(cudaGetLastError());
        return GOFPID_OK;
    }
    if (bitimg) {
        set_error("internal: the byte morphology path does not produce the foreground bit image");
        return GOFPID_ERR_INVALID;
    }

    if (plan.smem_bytes > 200 * 1024) {
        set_error("morphology chain needs %zu bytes of shared memory (halo too large)",
                  plan.smem_bytes);
        return GOFPID_ERR_UNSUPPORTED;
    }
    if (plan.smem_bytes > 48 * 1024)
        GOFPID_CUDA_TRY(cudaFuncSetAttribute(morph_chain_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)plan.smem_bytes));
    const dim3 grid((W + plan.tile_w - 1) / plan.tile_w, (H + plan.tile_h - 1) / plan.tile_h,
                    n_streams);
    morph_chain_kernel<<<grid, MORPH_THREADS, plan.smem_bytes, s>>>(plan, src, dst, H, W);
    GOFPID_CUDA_TRY(cudaGetLastError());
    return GOFPID_OK;
}

} // namespace gofpid
