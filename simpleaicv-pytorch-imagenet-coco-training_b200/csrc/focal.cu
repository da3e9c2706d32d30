// focal.cu -- the classification sweep of RetinaLoss / FCOSLoss: clamp + focal loss (+ its
// gradient) in ONE streaming pass over the per-level cls tensors (no torch.cat, no one-hot,
// no temporaries), plus the deterministic fp64 reduction of all block partials.
//
// Roofline: HBM.  Algorithmic bytes = 4*C per row read (+4*C written when gradients are
// requested) + 4 per row for the label; every byte is touched exactly once with 128-bit
// coalesced loads, 4 loads in flight per thread.
//
// ALU budget matters here (one log per element): at 6.5 TB/s the SMs have ~22 issue slots per
// float32 element.  Background elements (>99.9 % of the tensor) take a branch-free path:
//   x  = 1 - (1 - max(p, 1e-4))                 (the reference's (1 - pt), losses.py:248-250)
//   -log(1 - x) = x * S(x),  S = degree-5 polynomial on [0, 1/4]  (|rel err| <= 1.3e-7)
//   term = x^gamma * x * S(x)
// (~12 instructions, no MUFU); elements with p > 1/4, the target class of positive rows and
// gamma != 2 fall back to logf/powf.  Loss tolerance vs the reference: 1e-5 relative.
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "focal_terms.cuh"

namespace b200det {

// ---------------------------------------------------------------------------------------
// host-side geometry + bookkeeping
// ---------------------------------------------------------------------------------------
static std::atomic<unsigned long long> g_launches{0};
thread_local bool g_skip_memset = false;
thread_local int g_assign_chunk = 0;

// ---- per-kernel event profiler -------------------------------------------------------------
struct ProfRec {
    int id;
    cudaEvent_t a, b;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

ProfScope::ProfScope(int kernel_id, void *stream) : slot(-1), st((cudaStream_t)stream) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    r.id = kernel_id;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_prof[slot].b, st);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }

int make_geo(const b200det_geometry *s, Geo *g) {
    if (!s || !g) return B200DET_EINVAL;
    if (s->n_levels < 1 || s->n_levels > kMaxLevels) return B200DET_ERANGE;
    if (s->batch < 1 || s->num_classes < 1) return B200DET_EINVAL;
    if (s->per_loc < 1 || s->per_loc > kMaxPerLoc) return B200DET_ERANGE;
    g->n_levels = s->n_levels;
    g->batch = s->batch;
    g->per_loc = s->per_loc;
    g->num_classes = s->num_classes;
    long long off = 0;
    for (int l = 0; l < kMaxLevels; ++l) {
        g->H[l] = g->W[l] = g->rows[l] = 0;
        g->stride[l] = 0.f;
        g->off[l] = (int)off;
        if (l < s->n_levels) {
            if (s->height[l] < 1 || s->width[l] < 1 || !(s->stride[l] > 0.f)) return B200DET_EINVAL;
            const long long rows = (long long)s->height[l] * s->width[l] * s->per_loc;
            if (rows * s->batch >= (1ll << 31)) return B200DET_ERANGE;
            if (rows * s->batch * s->num_classes >= (1ll << 38)) return B200DET_ERANGE;
            g->H[l] = s->height[l];
            g->W[l] = s->width[l];
            g->rows[l] = (int)rows;
            g->stride[l] = s->stride[l];
            off += rows;
            if (off * s->batch >= (1ll << 31)) return B200DET_ERANGE;
        }
    }
    for (int l = s->n_levels; l <= kMaxLevels; ++l) g->off[l] = (int)off;
    return 0;
}

constexpr int kFocalThreads = 256;
#ifndef B200DET_FOCAL_UNROLL
#define B200DET_FOCAL_UNROLL 4
#endif
#ifndef B200DET_FOCAL_BATCHES
#define B200DET_FOCAL_BATCHES 2
#endif
constexpr int kFocalUnroll = B200DET_FOCAL_UNROLL;                 // loads in flight per thread
constexpr int kFocalBatches = B200DET_FOCAL_BATCHES;               // batches per chunk
constexpr int kChunkUnits = kFocalThreads * kFocalUnroll * kFocalBatches;  // 2048 units / CTA

static inline int focal_vec(const Geo &g) { return (g.num_classes % 4 == 0) ? 4 : 1; }

LossWs loss_ws_layout(const Geo &g) {
    LossWs w;
    const long long N = g.off[g.n_levels];
    w.assign_blocks_per_image = (size_t)assign_blocks_per_image(g);
    w.assign_blocks = w.assign_blocks_per_image * (size_t)g.batch;
    w.sparse_blocks = (size_t)sparse_blocks(g);
    // vector width is a pure function of C (float4 when C % 4 == 0; pointers must then be
    // 16-byte aligned), so the partial count is exact
    size_t chunks = 0;
    for (int l = 0; l < g.n_levels; ++l) {
        const long long units = (long long)g.batch * g.rows[l] * (g.num_classes / focal_vec(g));
        chunks += (size_t)((units + kChunkUnits - 1) / kChunkUnits);
    }
    w.focal_chunks = chunks;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    w.off_assign = 0;
    w.off_sparse = up(w.assign_blocks * sizeof(int));
    w.off_focal = w.off_sparse + up(w.sparse_blocks * sizeof(SparsePartial));
    w.off_counters = w.off_focal + up(kSweepWords * sizeof(long long));
    w.off_pos_queue = w.off_counters + 256;
    w.off_ign_queue = w.off_pos_queue + up((size_t)g.batch * (size_t)N * sizeof(int2));
    w.total = w.off_ign_queue + up((size_t)g.batch * (size_t)N * sizeof(int));
    return w;
}

// ---------------------------------------------------------------------------------------
// focal kernels
// ---------------------------------------------------------------------------------------
struct FocalArgs {
    PtrTab cls;
    MutPtrTab grad;
    long long units[kMaxLevels];          // units (float4 or float) per level, < 2^31
    long long row_base[kMaxLevels];       // level-major row base of level l (= B*off_l)
    int chunk_off[kMaxLevels + 1];        // first chunk of level l
    unsigned magic;                       // row = umulhi(u, magic) >> magic_shift  (u < 2^31)
    int magic_shift;                      // -1: units_per_row == 1 (row = u)
    int units_per_row;                    // C/4 or C
    int n_levels;
    float alpha, gamma;
    float grad_scale;
    const double *sums;
};

__device__ __forceinline__ int chunk_level(const FocalArgs &a) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i)
        if (i < a.n_levels && (int)blockIdx.x >= a.chunk_off[i]) l = i;
    return l;
}

#ifndef B200DET_FOCAL_LOAD
// 0: ld.global.cs; 1: ld.global.nc.L1::no_allocate -- measured identical (tools/tune_loads.sh: 3.285 vs
// 3.314 ms per batch-256 step with the assignment running beside the sweep, 0.457 vs 0.461 at batch
// 32): what the co-running kernels cost the sweep is issue slots, not L1 capacity
#define B200DET_FOCAL_LOAD 0
#endif
__device__ __forceinline__ float4 sweep_ld4(const float4 *p) {
#if B200DET_FOCAL_LOAD == 1
    float4 t;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                 : "l"(p));
    return t;
#else
    return __ldcs(p);
#endif
}

template <int VEC>
__device__ __forceinline__ void load_unit(const float *__restrict__ src, long long u, float (&v)[VEC]) {
    if (VEC == 4) {
        const float4 t = sweep_ld4(reinterpret_cast<const float4 *>(src) + u);
        v[0] = t.x;
        v[1 % VEC] = t.y;
        v[2 % VEC] = t.z;
        v[3 % VEC] = t.w;
    } else {
        v[0] = __ldcs(src + u);
    }
}

__device__ __forceinline__ void block_store_partial(float value, long long *slots) {
    sweep_accumulate<kFocalThreads>(value, slots);
}

// Label-free sweep (forward only): every element is treated as background; the rows that are
// not (positives' target class, ignored rows) are corrected by the assignment kernel, which
// knows them.  No label traffic, no index arithmetic beyond the load address.
template <int VEC, bool GAMMA2, bool FULL>
__device__ __forceinline__ float focal_all_chunk(const float *__restrict__ src, long long chunk_start,
                                                 long long n_units, float gamma) {
    float acc = 0.f;
    float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int bt = 0; bt < kFocalBatches; ++bt) {
        const long long u0 = chunk_start + (long long)bt * kFocalThreads * kFocalUnroll + threadIdx.x;
        float v[kFocalUnroll][VEC];
#pragma unroll
        for (int k = 0; k < kFocalUnroll; ++k) {
            const long long u = u0 + (long long)k * kFocalThreads;
            if (FULL || u < n_units) {
                load_unit<VEC>(src, u, v[k]);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[k][e] = -1.f;  // marks "no element"
            }
        }
#pragma unroll
        for (int k = 0; k < kFocalUnroll; ++k) {
            float x[VEC];
            float mx = 0.f;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                x[e] = fmax_nan(v[k][e], kClampLo);
                mx = fmaxf(mx, x[e]);
            }
            if (!FULL && v[k][0] < 0.f) continue;
            if (GAMMA2 && mx <= kFastMax) {
                if (VEC == 4) {
                    float2 xr, xs;
                    const float2 x01 = make_float2(x[0], x[1 % VEC]);
                    const float2 x23 = make_float2(x[2 % VEC], x[3 % VEC]);
                    acc2 = neg_term_fast2_acc(x01, acc2, xr, xs);
                    acc2 = neg_term_fast2_acc(x23, acc2, xr, xs);
                } else {
                    float xr, xs;
                    acc += neg_term_fast(x[0], xr, xs);
                }
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc += neg_term_slow(v[k][e], gamma, GAMMA2);
            }
        }
    }
    return acc + (acc2.x + acc2.y);
}

template <int VEC, bool GAMMA2>
__global__ void __launch_bounds__(kFocalThreads)
    focal_all_kernel(FocalArgs a, long long *__restrict__ partials) {
    const int l = chunk_level(a);
    const long long chunk_start = (long long)(blockIdx.x - a.chunk_off[l]) * kChunkUnits;
    const long long n_units = a.units[l];
    const float *__restrict__ src = static_cast<const float *>(a.cls.p[l]);
    float acc;
    if (chunk_start + kChunkUnits <= n_units)
        acc = focal_all_chunk<VEC, GAMMA2, true>(src, chunk_start, n_units, a.gamma);
    else
        acc = focal_all_chunk<VEC, GAMMA2, false>(src, chunk_start, n_units, a.gamma);
    block_store_partial((1.f - a.alpha) * acc, partials);
}

// exact-form element with gradient (reference op order, accurate log/pow)
template <bool GRAD>
__device__ __forceinline__ void slow_element(float p, bool is_target, float alpha, float gamma,
                                             bool gamma2, float &acc_pos, float &acc_neg,
                                             float &g) {
    const float pc = clamp_prob(p);
    const bool in_range = (p >= kClampLo) && (p <= kClampHi);
    if (is_target) {
        const float om = 1.f - pc;  // 1 - pt
        const float w = gamma2 ? om * om : powf(om, gamma);
        const float lg = logf(pc);
        acc_pos += w * (-lg);
        if (GRAD) {
            // d/dp [ om^g * (-log p) ] = g*om^(g-1)*log p - om^g / p
            const float wm1 = gamma2 ? om : powf(om, gamma - 1.f);
            g = in_range ? alpha * (gamma * wm1 * lg - w / pc) : 0.f;
        }
    } else {
        const float q = 1.f - pc;   // pt
        const float x = 1.f - q;    // 1 - pt
        const float w = gamma2 ? x * x : powf(x, gamma);
        const float lg = logf(q);
        acc_neg += w * (-lg);
        if (GRAD) {
            // d/dp [ x^g * (-log(1-p)) ] = g*x^(g-1)*(-log q) + x^g / q
            const float wm1 = gamma2 ? x : powf(x, gamma - 1.f);
            g = in_range ? (1.f - alpha) * (gamma * wm1 * (-lg) + w / q) : 0.f;
        }
    }
}

// One 4-class unit through the exact-form path, out of line: it runs for < 1 % of the units (a
// target class inside the unit, a probability above the polynomial's range, gamma != 2), and
// inlining it at all 32 call sites made the kernel 77 KB of code (instruction-fetch stalls were the
// top stall reason in ncu).
struct SlowUnit {
    float4 g;
    float pos, neg;
};
template <bool GRAD>
__device__ __noinline__ SlowUnit slow_unit4(float4 v, int tgt, float alpha, float gamma, bool gamma2) {
    SlowUnit r;
    r.pos = r.neg = 0.f;
    r.g = make_float4(0.f, 0.f, 0.f, 0.f);
    slow_element<GRAD>(v.x, tgt == 0, alpha, gamma, gamma2, r.pos, r.neg, r.g.x);
    slow_element<GRAD>(v.y, tgt == 1, alpha, gamma, gamma2, r.pos, r.neg, r.g.y);
    slow_element<GRAD>(v.z, tgt == 2, alpha, gamma, gamma2, r.pos, r.neg, r.g.z);
    slow_element<GRAD>(v.w, tgt == 3, alpha, gamma, gamma2, r.pos, r.neg, r.g.w);
    return r;
}

// A 4-float unit that straddles two rows (class counts that are not a multiple of 4, XROW below):
// the first 4 - nx elements are classes c0.. of a row labelled lab0, the last nx elements are classes
// 0.. of the next row, labelled lab1.  One unit in C / 4; exact-form path, out of line.
template <bool GRAD>
__device__ __noinline__ SlowUnit cross_unit4(float4 v, int lab0, int tgt0, int lab1, int nx,
                                             float alpha, float gamma, bool gamma2) {
    SlowUnit r;
    r.pos = r.neg = 0.f;
    const float x[4] = {v.x, v.y, v.z, v.w};
    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool first = e < 4 - nx;
        const int lab = first ? lab0 : lab1;
        const bool is_target = first ? (lab0 > 0 && tgt0 == e) : (lab1 - 1 == e - (4 - nx));
        if (lab >= 0) slow_element<GRAD>(x[e], is_target, alpha, gamma, gamma2, r.pos, r.neg, g[e]);
    }
    r.g = make_float4(g[0], g[1], g[2], g[3]);
    return r;
}

// Label-aware sweep: used when gradients are requested (one read of cls, one write of its
// gradient already scaled by weight / positives), or when the caller did not let the
// assignment kernel apply the corrections.
// XROW (VEC == 4 only): the class count is not a multiple of 4 (Objects365: 365), but every level
// holds a multiple of 4 floats and is 16-byte aligned, so the level is still swept as a flat stream
// of 128-bit units; a unit's (row, first class) come from its element index, and the one unit in
// C / 4 that straddles two rows takes cross_unit4.  The scalar kernel (VEC == 1) this replaces ran
// 75 instructions per element, issue-bound at half the HBM peak (profiles/r01_cfg4_sweeps.txt).
#ifndef B200DET_FOCAL_GRAD_MINB
#define B200DET_FOCAL_GRAD_MINB 6
#endif
#ifndef B200DET_FOCAL_XROW_MINB
#define B200DET_FOCAL_XROW_MINB 5
#endif
template <int VEC, bool GRAD, bool GAMMA2, bool XROW = false>
__global__ void __launch_bounds__(kFocalThreads, XROW ? B200DET_FOCAL_XROW_MINB : B200DET_FOCAL_GRAD_MINB)
    focal_kernel(FocalArgs a, const int *__restrict__ labels, long long *__restrict__ partials) {
    static_assert(!XROW || VEC == 4, "XROW sweeps 128-bit units");
    const int l = chunk_level(a);
    const long long chunk_start = (long long)(blockIdx.x - a.chunk_off[l]) * kChunkUnits;
    // XROW: row and class of the chunk's first element (one 64-bit division per CTA); units_per_row
    // and magic are those of the divisor C here
    unsigned row_c = 0, rem_c = 0;
    if (XROW) {
        const unsigned long long e_chunk = (unsigned long long)chunk_start * 4ull;
        const unsigned long long q = e_chunk / (unsigned)a.units_per_row;
        row_c = (unsigned)q;
        rem_c = (unsigned)(e_chunk - q * (unsigned)a.units_per_row);
    }
    const long long n_units = a.units[l];
    const float *__restrict__ src = static_cast<const float *>(a.cls.p[l]);
    float *__restrict__ dst = GRAD ? static_cast<float *>(a.grad.p[l]) : nullptr;
    const int *__restrict__ lab_l = labels + a.row_base[l];

    float gs = 0.f;
    if (GRAD) {
        const double npos = a.sums[0];
        // a NaN count (failed cross-rank exchange) must poison the gradient, not zero it
        gs = npos > 0.0 ? (float)((double)a.grad_scale / npos) : (npos != npos ? (float)npos : 0.f);
    }
    const float one_m_alpha = 1.f - a.alpha;

    float acc_neg = 0.f, acc_pos = 0.f;
    float2 acc2 = make_float2(0.f, 0.f);
    // XROW: one copy of the batch body (the second copy made the kernel 95 KB of code and
    // instruction-fetch stalls its top stall reason, profiles/r01_cfg4_sweeps.txt)
#pragma unroll(XROW ? 1 : kFocalBatches)
    for (int bt = 0; bt < kFocalBatches; ++bt) {
        const long long u0 = chunk_start + (long long)bt * kFocalThreads * kFocalUnroll + threadIdx.x;
        float v[kFocalUnroll][VEC];
        int lab[kFocalUnroll];
        int tgt[kFocalUnroll];
#pragma unroll
        for (int k = 0; k < kFocalUnroll; ++k) {
            const long long u = u0 + (long long)k * kFocalThreads;
            lab[k] = -1;
            tgt[k] = -1;
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[k][e] = 0.f;
            if (u < n_units) {
                unsigned row;
                int c0;
                if (XROW) {
                    const unsigned el = 4u * (unsigned)(u - chunk_start) + rem_c;
                    const unsigned r_off = __umulhi(el, a.magic) >> a.magic_shift;
                    c0 = (int)(el - r_off * (unsigned)a.units_per_row);
                    row = row_c + r_off;
                } else {
                    const unsigned uu = (unsigned)u;
                    row = a.magic_shift < 0 ? uu : (__umulhi(uu, a.magic) >> a.magic_shift);
                    c0 = (int)(uu - row * (unsigned)a.units_per_row) * VEC;
                }
                load_unit<VEC>(src, u, v[k]);
                const int lb = __ldg(lab_l + row);
                lab[k] = lb;
                tgt[k] = lb - 1 - c0;  // position of the target class inside this unit
            }
        }
#pragma unroll
        for (int k = 0; k < kFocalUnroll; ++k) {
            const long long u = u0 + (long long)k * kFocalThreads;
            float g[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) g[e] = 0.f;
            // XROW: a warp's 128 elements meet a row boundary in a third of its iterations, so the
            // straddling unit (nx > 0 of its elements belong to the next row) stays on the fast path
            // whenever it can: both rows valid and no target class inside -> a plain background unit;
            // both ignored -> skipped; the rest (next to a positive / an ignored row) is exact-form.
            // The next row's label is read here, not in the load phase (registers): it is the label
            // the neighbouring lanes have just loaded, an L1 hit.
            int nx = 0, lb1 = -1;
            bool cross_slow = false;
            if (XROW) {
                nx = (lab[k] - 1 - tgt[k]) + 4 - a.units_per_row;   // c0 + 4 - C
                if (nx > 0) {
                    const unsigned el = 4u * (unsigned)(u - chunk_start) + rem_c;
                    lb1 = __ldg(lab_l + row_c + (__umulhi(el, a.magic) >> a.magic_shift) + 1);
                    const int lb = lab[k];
                    const bool t0 = lb > 0 && (unsigned)tgt[k] < 4u;
                    const bool t1 = lb1 > 0 && lb1 - 1 < nx;
                    if (lb >= 0 && lb1 >= 0 && !t0 && !t1) {
                        lab[k] = 0;
                        tgt[k] = -1;
                    } else if (lb >= 0 || lb1 >= 0) {
                        cross_slow = true;
                    }
                }
            }
            // XROW: ONE out-of-line exact-form function for straddling and ordinary units (code size)
            if (XROW && !cross_slow && lab[k] >= 0) {
                float mx = 0.f;
#pragma unroll
                for (int e = 0; e < VEC; ++e) mx = fmaxf(mx, fmax_nan(v[k][e], kClampLo));
                const bool has_target = lab[k] > 0 && (unsigned)tgt[k] < (unsigned)VEC;
                if (!(GAMMA2 && !has_target && mx <= kFastMax)) {
                    cross_slow = true;
                    nx = 0;
                    if (!has_target) tgt[k] = -1;
                }
            }
            if (XROW && cross_slow) {
                const SlowUnit r = cross_unit4<GRAD>(
                    make_float4(v[k][0], v[k][1 % VEC], v[k][2 % VEC], v[k][3 % VEC]), lab[k], tgt[k],
                    lb1, nx, a.alpha, a.gamma, GAMMA2);
                acc_pos += r.pos;
                acc_neg += r.neg;
                g[0] = r.g.x, g[1 % VEC] = r.g.y, g[2 % VEC] = r.g.z, g[3 % VEC] = r.g.w;
            } else if (lab[k] >= 0) {
                float x[VEC];
                float mx = 0.f;
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    x[e] = fmax_nan(v[k][e], kClampLo);
                    mx = fmaxf(mx, x[e]);
                }
                const bool has_target = lab[k] > 0 && (unsigned)tgt[k] < (unsigned)VEC;
                if (XROW || (GAMMA2 && !has_target && mx <= kFastMax)) {   // XROW: decided above
                    if (VEC == 4) {
                        // packed FP32 (FFMA2): two elements per instruction
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float2 xr, xs;
                            const float2 xx = make_float2(x[(2 * h) % VEC], x[(2 * h + 1) % VEC]);
                            acc2 = neg_term_fast2_acc(xx, acc2, xr, xs);
                            if (GRAD) {
                                // (1-a) * x * (2*(-log q) + x/q), zero below the clamp
                                const float2 rq = make_float2(__fdividef(1.f, 1.f - xr.x),
                                                              __fdividef(1.f, 1.f - xr.y));
                                const float2 t = __ffma2_rn(make_float2(2.f, 2.f), xs,
                                                            __fmul2_rn(xr, rq));
                                const float2 gg = __fmul2_rn(__fmul2_rn(xr, t),
                                                             make_float2(one_m_alpha, one_m_alpha));
                                g[(2 * h) % VEC] = v[k][(2 * h) % VEC] >= kClampLo ? gg.x : 0.f;
                                g[(2 * h + 1) % VEC] = v[k][(2 * h + 1) % VEC] >= kClampLo ? gg.y : 0.f;
                            }
                        }
                    } else {
                        float xr, xs;
                        acc_neg += neg_term_fast(x[0], xr, xs);
                        if (GRAD) {
                            const float t = fmaf(2.f, xs, __fdividef(xr, 1.f - xr));
                            g[0] = v[k][0] >= kClampLo ? one_m_alpha * xr * t : 0.f;
                        }
                    }
                } else if (VEC == 4) {
                    const SlowUnit r = slow_unit4<GRAD>(
                        make_float4(v[k][0], v[k][1 % VEC], v[k][2 % VEC], v[k][3 % VEC]),
                        has_target ? tgt[k] : -1, a.alpha, a.gamma, GAMMA2);
                    acc_pos += r.pos;
                    acc_neg += r.neg;
                    g[0] = r.g.x, g[1 % VEC] = r.g.y, g[2 % VEC] = r.g.z, g[3 % VEC] = r.g.w;
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e)
                        slow_element<GRAD>(v[k][e], has_target && tgt[k] == e, a.alpha, a.gamma,
                                           GAMMA2, acc_pos, acc_neg, g[e]);
                }
            }
            if (GRAD && u < n_units) {
                if (VEC == 4) {
                    __stcs(reinterpret_cast<float4 *>(dst) + u,
                           make_float4(g[0] * gs, g[1 % VEC] * gs, g[2 % VEC] * gs, g[3 % VEC] * gs));
                } else {
                    __stcs(dst + u, g[0] * gs);
                }
            }
        }
    }
    // one fixed-point atomic per WARP: no CTA barrier at the end of these short CTAs
    sweep_accumulate_warp(a.alpha * acc_pos + one_m_alpha * (acc_neg + (acc2.x + acc2.y)), partials);
}

// ---------------------------------------------------------------------------------------
// deterministic reduction of block partials (fixed order, fp64)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
    loss_reduce_kernel(const int *__restrict__ npos, long long n_assign,
                       const SparsePartial *__restrict__ sp, long long n_sparse,
                       const long long *__restrict__ fp, long long n_focal, int which,
                       double *__restrict__ sums, float w_cls, float w_box, float w_ctr,
                       float *__restrict__ losses) {
    __shared__ double red[4][32];
    // a dependent launched programmatically (the decoder's select kernel on handed-over keys, which
    // reads nothing this kernel writes) may run beside this one-CTA kernel
    pdl_launch_dependents();
    double s_pos = 0.0, s_cls = 0.0, s_box = 0.0, s_ctr = 0.0;
    if (which & 5)   // bit 2: the positive count alone (it is complete as soon as the assignment is)
        for (long long i = threadIdx.x; i < n_assign; i += blockDim.x) s_pos += (double)npos[i];
    if (which & 1) {
        for (long long i = threadIdx.x; i < n_sparse; i += blockDim.x) {
            const SparsePartial p = sp[i];
            s_box += p.box;
            s_ctr += p.ctr;
            s_cls += p.focal;  // focal corrections (zero unless the sweep ran label-free)
        }
    }
    if (which & 2) {
        for (long long i = threadIdx.x; i < n_focal; i += blockDim.x)
            s_cls += (double)fp[i] / kFxSweep;
        if (threadIdx.x == 0 && fp[n_focal] != 0) s_cls = __longlong_as_double(0x7ff8000000000000ll);
    }
    // fixed-order tree: xor-shuffle inside the warp, then warp 0 over the 32 warp sums
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_pos += __shfl_xor_sync(0xffffffffu, s_pos, o);
        s_cls += __shfl_xor_sync(0xffffffffu, s_cls, o);
        s_box += __shfl_xor_sync(0xffffffffu, s_box, o);
        s_ctr += __shfl_xor_sync(0xffffffffu, s_ctr, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = s_pos;
        red[1][warp] = s_cls;
        red[2][warp] = s_box;
        red[3][warp] = s_ctr;
    }
    __syncthreads();
    if (warp == 0) {
        double a = red[0][lane], b = red[1][lane], c = red[2][lane], d = red[3][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
            d += __shfl_xor_sync(0xffffffffu, d, o);
        }
        if (lane == 0) {
            if (which & 5) sums[0] = a;
            if (which & 1) {
                sums[2] = c;
                sums[3] = d;
            }
            if (which & 3) sums[1] = b;  // focal partials (bit 1) + the sparse kernel's corrections (bit 0)
            if (losses) {
                // loss_finish_kernel fused (which == 3): float32 sum / count, then * weight
                const float w[3] = {w_cls, w_box, w_ctr};
                const double t[3] = {b, c, d};
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    float v = 0.f;
                    if (a > 0.0) v = w[i] * ((float)t[i] / (float)a);
                    losses[i] = v;
                }
            }
        }
    }
}

__global__ void loss_finish_kernel(const double *__restrict__ sums, float w_cls, float w_box,
                                   float w_ctr, float *__restrict__ losses) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double npos = sums[0];
        const float w[3] = {w_cls, w_box, w_ctr};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float v = 0.f;
            // float32 sum / count, then * weight  (losses.py:259, :293, :318, :210-211)
            if (npos > 0.0) v = w[i] * ((float)sums[1 + i] / (float)npos);
            if (npos != npos) v = (float)npos;   // failed cross-rank exchange: NaN, never a silent 0
            losses[i] = v;
        }
    }
}

struct ScaleArgs {
    float *ptr[kMaxLevels];
    long long count[kMaxLevels];
};
// x[l][i] *= (*g) * (sums ? weight / sums[0] : 1); whole call skipped when the factor is 1
__global__ void scale_levels_kernel(ScaleArgs a, const float *__restrict__ g,
                                    const double *__restrict__ sums, float weight) {
    float k = *g;
    if (sums) {
        const double npos = sums[0];
        k = npos > 0.0 ? (float)((double)k * (double)weight / npos) : (npos != npos ? (float)npos : 0.f);
    }
    if (k == 1.f) return;
    float *x = a.ptr[blockIdx.y];
    const long long n = a.count[blockIdx.y];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        x[i] *= k;
}

__global__ void scale_kernel(float *__restrict__ x, long long n, const float *__restrict__ s) {
    const float k = *s;
    if (k == 1.f) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        x[i] *= k;
}

}  // namespace b200det

using namespace b200det;

extern "C" int b200det_abi_version(void) { return B200DET_ABI_VERSION; }

extern "C" const char *b200det_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case B200DET_EINVAL: return "invalid argument (null pointer, bad enum or bad size)";
        case B200DET_ERANGE: return "size exceeds a compiled limit";
        case B200DET_EWORKSPACE: return "workspace too small";
        case B200DET_EALIGN: return "pointer not sufficiently aligned";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

extern "C" unsigned long long b200det_launch_count(void) { return launches(); }

extern "C" int b200det_profile(int enable) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (enable == 1) {
        for (auto &r : g_prof) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        g_prof.clear();
    }
    g_prof_on = enable != 0;   // 1: clear + record, 2: resume (keep the records), 0: pause
    return 0;
}

extern "C" int b200det_profile_read(int kernel_id, double *total_ms, int *n_launches) {
    if (kernel_id < 0 || kernel_id >= kKernCount || !total_ms || !n_launches) return B200DET_EINVAL;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double t = 0.0;
    int n = 0;
    for (auto &r : g_prof) {
        if (r.id != kernel_id) continue;
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.a, r.b);
        if (e != cudaSuccess) return (int)e;
        t += ms;
        ++n;
    }
    *total_ms = t;
    *n_launches = n;
    return 0;
}

extern "C" const char *b200det_kernel_name(int kernel_id) {
    static const char *names[kKernCount] = {"focal_loss",  "assign",       "sparse_losses",
                                            "loss_reduce", "loss_finish",  "score_argmax",
                                            "select_decode_nms", "other", "head_tail", "logits_sweep"};
    return (kernel_id >= 0 && kernel_id < kKernCount) ? names[kernel_id] : "?";
}

extern "C" long long b200det_rows_per_image(const b200det_geometry *geo) {
    Geo g;
    const int rc = make_geo(geo, &g);
    if (rc) return rc;
    return g.off[g.n_levels];
}

extern "C" size_t b200det_loss_workspace_bytes(const b200det_geometry *geo) {
    Geo g;
    if (make_geo(geo, &g)) return 0;
    return loss_ws_layout(g).total;
}

static cudaError_t launch_focal_xrow(const FocalArgs &a, int chunks, bool grad, bool gamma2,
                                     const int *labels, long long *partials, cudaStream_t st) {
    if (grad) {
        if (gamma2)
            focal_kernel<4, true, true, true><<<chunks, kFocalThreads, 0, st>>>(a, labels, partials);
        else
            focal_kernel<4, true, false, true><<<chunks, kFocalThreads, 0, st>>>(a, labels, partials);
    } else {
        if (gamma2)
            focal_kernel<4, false, true, true><<<chunks, kFocalThreads, 0, st>>>(a, labels, partials);
        else
            focal_kernel<4, false, false, true><<<chunks, kFocalThreads, 0, st>>>(a, labels, partials);
    }
    return cudaGetLastError();
}

template <int VEC>
static cudaError_t launch_focal(const FocalArgs &a, int chunks, bool grad, bool gamma2,
                                const int *labels, long long *partials, cudaStream_t st) {
    if (grad) {
        if (gamma2)
            focal_kernel<VEC, true, true><<<chunks, kFocalThreads, 0, st>>>(a, labels, partials);
        else
            focal_kernel<VEC, true, false><<<chunks, kFocalThreads, 0, st>>>(a, labels, partials);
    } else {
        if (gamma2)
            focal_kernel<VEC, false, true><<<chunks, kFocalThreads, 0, st>>>(a, labels, partials);
        else
            focal_kernel<VEC, false, false><<<chunks, kFocalThreads, 0, st>>>(a, labels, partials);
    }
    return cudaGetLastError();
}

extern "C" int b200det_focal_loss(const b200det_geometry *geo, const void *const *cls,
                                  const int32_t *labels, float alpha, float gamma,
                                  void *const *cls_grad, const double *sums, float grad_scale,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!cls || !workspace) return B200DET_EINVAL;
    if (cls_grad && (!sums || !labels)) return B200DET_EINVAL;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;

    FocalArgs a;
    int vec = focal_vec(g);
    // the label-free sweep has no notion of rows: it can use 128-bit loads for ANY class count as
    // long as every level holds a multiple of 4 floats and is 16-byte aligned (e.g. C = 365)
    // ... and so can the label-aware sweep (xrow): a unit's row comes from its element index
    bool flat4 = false, xrow = false;
    if (vec == 1 && g.num_classes >= 4 && g.num_classes < 16384) {
        flat4 = true;
        for (int l = 0; l < g.n_levels; ++l) {
            if (((long long)g.batch * g.rows[l] * g.num_classes) & 3) flat4 = false;
            if (cls[l] && (reinterpret_cast<uintptr_t>(cls[l]) & 15)) flat4 = false;
            if (cls_grad && cls_grad[l] && (reinterpret_cast<uintptr_t>(cls_grad[l]) & 15))
                flat4 = false;
        }
        const bool no_xrow = getenv("B200DET_FOCAL_NO_XROW") != nullptr;   // tests compare both
        if (labels != nullptr) {
            xrow = flat4 && !no_xrow;
            flat4 = false;
        }
        if (flat4 || xrow) vec = 4;
    }
    const uintptr_t amask = vec == 4 ? 15 : 3;
    for (int l = 0; l < kMaxLevels; ++l) {
        a.cls.p[l] = nullptr;
        a.grad.p[l] = nullptr;
        a.units[l] = 0;
        a.row_base[l] = 0;
    }
    for (int l = 0; l < g.n_levels; ++l) {
        if (!cls[l]) return B200DET_EINVAL;
        if (reinterpret_cast<uintptr_t>(cls[l]) & amask) return B200DET_EALIGN;
        a.cls.p[l] = cls[l];
        if (cls_grad) {
            if (!cls_grad[l]) return B200DET_EINVAL;
            if (reinterpret_cast<uintptr_t>(cls_grad[l]) & amask) return B200DET_EALIGN;
            a.grad.p[l] = cls_grad[l];
        }
    }
    a.units_per_row = flat4 ? 1 : (xrow ? g.num_classes : g.num_classes / vec);
    // row = u / d for u < 2^31 by multiply-high: s = ceil(log2 d), m = floor(2^(31+s)/d) + 1,
    // row = umulhi(u, m) >> (s - 1)   (Granlund-Montgomery, 31-bit dividends)
    if (a.units_per_row == 1) {
        a.magic = 0;
        a.magic_shift = -1;
    } else {
        int sft = 0;
        while ((1u << sft) < (unsigned)a.units_per_row) ++sft;
        a.magic = (unsigned)(((1ull << (31 + sft)) / (unsigned long long)a.units_per_row) + 1ull);
        a.magic_shift = sft - 1;
    }
    a.n_levels = g.n_levels;
    a.alpha = alpha;
    a.gamma = gamma;
    a.grad_scale = grad_scale;
    a.sums = sums;
    int chunks = 0;
    for (int l = 0; l < g.n_levels; ++l) {
        a.units[l] = (flat4 || xrow) ? (long long)g.batch * g.rows[l] * g.num_classes / 4
                           : (long long)g.batch * g.rows[l] * a.units_per_row;
        if (a.units[l] >= (1ll << 31)) return B200DET_ERANGE;
        a.row_base[l] = (long long)g.batch * g.off[l];
        a.chunk_off[l] = chunks;
        chunks += (int)((a.units[l] + kChunkUnits - 1) / kChunkUnits);
    }
    for (int l = g.n_levels; l <= kMaxLevels; ++l) a.chunk_off[l] = chunks;

    long long *partials =
        reinterpret_cast<long long *>(static_cast<char *>(workspace) + ws.off_focal);
    if (!g_skip_memset) {
        cudaError_t me = cudaMemsetAsync(partials, 0, kSweepWords * sizeof(long long),
                                         (cudaStream_t)stream);
        if (me != cudaSuccess) return (int)me;
    }
    const bool gamma2 = gamma == 2.f;
    cudaError_t e;
    ProfScope prof(kKernFocal, stream);
    if (labels == nullptr) {
        cudaStream_t st = (cudaStream_t)stream;
        if (vec == 4) {
            if (gamma2) focal_all_kernel<4, true><<<chunks, kFocalThreads, 0, st>>>(a, partials);
            else focal_all_kernel<4, false><<<chunks, kFocalThreads, 0, st>>>(a, partials);
        } else {
            if (gamma2) focal_all_kernel<1, true><<<chunks, kFocalThreads, 0, st>>>(a, partials);
            else focal_all_kernel<1, false><<<chunks, kFocalThreads, 0, st>>>(a, partials);
        }
        e = cudaGetLastError();
    } else if (xrow) {
        e = launch_focal_xrow(a, chunks, cls_grad != nullptr, gamma2, labels, partials,
                              (cudaStream_t)stream);
    } else {
        e = vec == 4 ? launch_focal<4>(a, chunks, cls_grad != nullptr, gamma2, labels, partials,
                                       (cudaStream_t)stream)
                     : launch_focal<1>(a, chunks, cls_grad != nullptr, gamma2, labels, partials,
                                       (cudaStream_t)stream);
    }
    count_launch();
    return (int)e;
}

extern "C" int b200det_loss_reduce(const b200det_geometry *geo, int which, const void *workspace,
                                   size_t workspace_bytes, double *sums, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!workspace || !sums || (which & 7) == 0 || (which & ~7)) return B200DET_EINVAL;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    const char *base = static_cast<const char *>(workspace);
    ProfScope prof(kKernReduce, stream);
    loss_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const int *>(base + ws.off_assign), (long long)ws.assign_blocks,
        reinterpret_cast<const SparsePartial *>(base + ws.off_sparse), (long long)ws.sparse_blocks,
        reinterpret_cast<const long long *>(base + ws.off_focal), (long long)kSweepSlots, which,
        sums, 0.f, 0.f, 0.f, nullptr);
    count_launch();
    return (int)cudaGetLastError();
}

// b200det_loss_reduce(which = 3) + b200det_loss_finish in one launch (the single-process forward)
extern "C" int b200det_loss_reduce_finish(const b200det_geometry *geo, const void *workspace,
                                          size_t workspace_bytes, float w_cls, float w_box,
                                          float w_ctr, double *sums, float *losses, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!workspace || !sums || !losses) return B200DET_EINVAL;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    const char *base = static_cast<const char *>(workspace);
    ProfScope prof(kKernReduce, stream);
    loss_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const int *>(base + ws.off_assign), (long long)ws.assign_blocks,
        reinterpret_cast<const SparsePartial *>(base + ws.off_sparse), (long long)ws.sparse_blocks,
        reinterpret_cast<const long long *>(base + ws.off_focal), (long long)kSweepSlots, 3, sums,
        w_cls, w_box, w_ctr, losses);
    count_launch();
    return (int)cudaGetLastError();
}

extern "C" int b200det_loss_finish(const double *sums, float w_cls, float w_box, float w_ctr,
                                   float *losses, void *stream) {
    if (!sums || !losses) return B200DET_EINVAL;
    ProfScope prof(kKernFinish, stream);
    loss_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, w_cls, w_box, w_ctr, losses);
    count_launch();
    return (int)cudaGetLastError();
}

extern "C" int b200det_scale_levels(void *const *ptrs, const long long *counts, int n_levels,
                                    const float *g_dev, const double *sums, float weight,
                                    void *stream) {
    if (!ptrs || !counts || !g_dev || n_levels < 1 || n_levels > kMaxLevels) return B200DET_EINVAL;
    ScaleArgs a;
    long long mx = 0;
    for (int l = 0; l < kMaxLevels; ++l) {
        a.ptr[l] = nullptr;
        a.count[l] = 0;
    }
    for (int l = 0; l < n_levels; ++l) {
        if (!ptrs[l] || counts[l] < 0) return B200DET_EINVAL;
        a.ptr[l] = static_cast<float *>(ptrs[l]);
        a.count[l] = counts[l];
        if (counts[l] > mx) mx = counts[l];
    }
    if (mx == 0) return 0;
    long long blocks = (mx + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    ProfScope prof(kKernOther, stream);
    scale_levels_kernel<<<dim3((unsigned)blocks, (unsigned)n_levels), 256, 0, (cudaStream_t)stream>>>(
        a, g_dev, sums, weight);
    count_launch();
    return (int)cudaGetLastError();
}

extern "C" int b200det_scale_f32(float *x, long long n, const float *scale_dev, void *stream) {
    if (!x || !scale_dev || n < 0) return B200DET_EINVAL;
    if (n == 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, scale_dev);
    count_launch();
    return (int)cudaGetLastError();
}
