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
    plan->smem_bytes = (size_t)3 * rows * stride;
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
struct Pair { uint32_t e, o; }; // even bytes (0,2) and odd bytes (1,3) of a word as u16x2

__device__ __forceinline__ Pair split(uint32_t w)
{
    Pair p;
    p.e = __byte_perm(w, 0, 0x4240);
    p.o = __byte_perm(w, 0, 0x4341);
    return p;
}
template <bool DIL>
__device__ __forceinline__ void acc(Pair &a, uint32_t w)
{
    const Pair p = split(w);
    if (DIL) { a.e = __vmaxu2(a.e, p.e); a.o = __vmaxu2(a.o, p.o); }
    else { a.e = __vminu2(a.e, p.e); a.o = __vminu2(a.o, p.o); }
}
__device__ __forceinline__ uint32_t join(const Pair &a) { return a.e | (a.o << 8); }

// min/max over `len` consecutive byte-shifted words starting at byte column `start`
template <bool DIL>
__device__ __forceinline__ void window(Pair &a, const uint32_t *row, int start, int len,
                                       int nwords)
{
    int wi = start >> 2, sh = start & 3;
    uint32_t lo = row[min(max(wi, 0), nwords - 1)];
    uint32_t hi = row[min(max(wi + 1, 0), nwords - 1)];
    for (int j = 0; j < len; j++) {
        acc<DIL>(a, __funnelshift_r(lo, hi, 8 * sh));
        if (++sh == 4) {
            sh = 0; wi++;
            lo = hi;
            hi = row[min(max(wi + 1, 0), nwords - 1)];
        }
    }
}

template <bool DIL>
__device__ __forceinline__ void run_op(const MorphOpDev &op, const MorphRun *__restrict__ runs,
                                       const uint8_t *A, uint8_t *T, uint8_t *B, int stride,
                                       int r0, int r1, int c0, int c1, // valid output region
                                       int gy0, int gx0,               // image coords of (0,0)
                                       int H, int W, uint32_t ident_next)
{
    const int tid = threadIdx.x;
    const int nwords = stride >> 2;
    const int w0 = c0 >> 2, w1 = (c1 + 3) >> 2; // word columns covering the valid region
    const int nw = w1 - w0;
    const Pair init = DIL ? Pair{0u, 0u} : Pair{0x00ff00ffu, 0x00ff00ffu};

    if (op.is_rect) {
        // row pass over the rows the column pass will read
        const int rr0 = r0 - op.ay, rr1 = r1 + (op.kh - 1 - op.ay);
        for (int i = tid; i < (rr1 - rr0) * nw; i += MORPH_THREADS) {
            const int r = rr0 + i / nw, wc = w0 + i % nw;
            Pair a = init;
            window<DIL>(a, reinterpret_cast<const uint32_t *>(A + (size_t)r * stride),
                        wc * 4 - op.ax, op.kw, nwords);
            reinterpret_cast<uint32_t *>(T + (size_t)r * stride)[wc] = join(a);
        }
        __syncthreads();
    }
    for (int i = tid; i < (r1 - r0) * nw; i += MORPH_THREADS) {
        const int r = r0 + i / nw, wc = w0 + i % nw;
        Pair a = init;
        if (op.is_rect) {
            for (int k = 0; k < op.kh; k++)
                acc<DIL>(a, reinterpret_cast<const uint32_t *>(T + (size_t)(r + k - op.ay) * stride)[wc]);
        } else {
            for (int q = op.run_begin; q < op.run_end; q++) {
                const MorphRun run = runs[q];
                window<DIL>(a, reinterpret_cast<const uint32_t *>(A + (size_t)(r + run.dy) * stride),
                            wc * 4 + run.dx, run.len, nwords);
            }
        }
        // cells outside the image take the identity of the NEXT step (OpenCV ignores them)
        uint32_t res = join(a);
        const int gy = gy0 + r, gx = gx0 + wc * 4;
        uint32_t inmask = 0;
        if (gy >= 0 && gy < H) {
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (gx + b >= 0 && gx + b < W) inmask |= 0xffu << (8 * b);
        }
        res = (res & inmask) | (ident_next & ~inmask);
        reinterpret_cast<uint32_t *>(B + (size_t)r * stride)[wc] = res;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(MORPH_THREADS)
morph_chain_kernel(MorphPlan plan, ImageU8 src, ImageU8 dst, int H, int W)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int hl4 = (plan.halo_l + 3) & ~3, hr4 = (plan.halo_r + 3) & ~3;
    const int stride = MORPH_SLACK + hl4 + plan.tile_w + hr4 + MORPH_SLACK;
    const int rows = plan.tile_h + plan.halo_t + plan.halo_b;
    const int nwords = stride >> 2;
    uint8_t *A = smem, *T = smem + (size_t)rows * stride, *B = smem + (size_t)2 * rows * stride;

    const int x0 = blockIdx.x * plan.tile_w, y0 = blockIdx.y * plan.tile_h;
    const int gx0 = x0 - hl4 - MORPH_SLACK; // image x of buffer column 0 (multiple of 4)
    const int gy0 = y0 - plan.halo_t;
    const uint8_t *sbase = src.base + (size_t)blockIdx.z * src.stream_stride;
    const bool aligned = ((reinterpret_cast<uintptr_t>(sbase) | src.pitch) & 3) == 0;
    const uint32_t ident0 = plan.n_ops && plan.ops[0].op == GOFPID_MORPH_ERODE ? 0xffffffffu : 0u;

    // ---- stage the tile + halo; out-of-image cells get the first step's identity --------
    for (int i = tid; i < rows * nwords; i += MORPH_THREADS) {
        const int r = i / nwords, wc = i - r * nwords;
        const int gy = gy0 + r, gx = gx0 + wc * 4;
        uint32_t v = ident0;
        if (gy >= 0 && gy < H && gx + 3 >= 0 && gx < W) {
            const uint8_t *row = sbase + (size_t)gy * src.pitch;
            if (aligned && gx >= 0 && gx + 3 < W) {
                v = *reinterpret_cast<const uint32_t *>(row + gx);
            } else {
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (gx + b >= 0 && gx + b < W)
                        v = (v & ~(0xffu << (8 * b))) | ((uint32_t)row[gx + b] << (8 * b));
            }
        }
        reinterpret_cast<uint32_t *>(A + (size_t)r * stride)[wc] = v;
    }
    __syncthreads();

    // ---- run the chain, shrinking the valid region by each step's extent -----------------
    int r0 = 0, r1 = rows, c0 = MORPH_SLACK + hl4 - plan.halo_l,
        c1 = MORPH_SLACK + hl4 + plan.tile_w + plan.halo_r;
    for (int k = 0; k < plan.n_ops; k++) {
        const MorphOpDev &op = plan.ops[k];
        r0 += op.ay; r1 -= op.kh - 1 - op.ay;
        c0 += op.ax; c1 -= op.kw - 1 - op.ax;
        const uint32_t ident_next =
            (k + 1 < plan.n_ops && plan.ops[k + 1].op == GOFPID_MORPH_ERODE) ? 0xffffffffu : 0u;
        if (op.op == GOFPID_MORPH_DILATE)
            run_op<true>(op, plan.d_runs, A, T, B, stride, r0, r1, c0, c1, gy0, gx0, H, W, ident_next);
        else
            run_op<false>(op, plan.d_runs, A, T, B, stride, r0, r1, c0, c1, gy0, gx0, H, W, ident_next);
        uint8_t *t = A; A = B; B = t;
    }

    // ---- write the tile -----------------------------------------------------------------
    uint8_t *dbase = dst.base + (size_t)blockIdx.z * dst.stream_stride;
    const bool daligned = ((reinterpret_cast<uintptr_t>(dbase) | dst.pitch) & 3) == 0;
    const int tw4 = plan.tile_w >> 2;
    const int cbase = MORPH_SLACK + hl4; // buffer column of x0
    for (int i = tid; i < plan.tile_h * tw4; i += MORPH_THREADS) {
        const int r = i / tw4, wc = i - r * tw4;
        const int gy = y0 + r, gx = x0 + wc * 4;
        if (gy >= H || gx >= W) continue;
        const uint32_t v = reinterpret_cast<const uint32_t *>(
            A + (size_t)(r + plan.halo_t) * stride + cbase)[wc];
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

int launch_morph(const MorphPlan &plan, ImageU8 src, ImageU8 dst, int H, int W, int n_streams,
                 cudaStream_t s)
{
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
