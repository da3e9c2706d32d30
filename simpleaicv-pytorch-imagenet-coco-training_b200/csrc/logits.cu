// logits.cu -- loss + decode sweeps that read the classification head's NCHW LOGITS directly
// (second half): no probability tensor is ever written to HBM.
//
// Reference chain this replaces for the classification branch:
//   x = cls_out(feat)                                   [B, A*C, H, W] logits
//   p = sigmoid(x.float())                              models/head.py:46-50
//   p = p.permute(0, 2, 3, 1).contiguous().view(...)    models/retinanet.py:73-76
//   focal loss over p  (losses.py:220-261)  /  arg-max, score, threshold over p  (decode.py:208-238)
// = 16 B per element for the head tail + 4 B per element for each consumer.  Here ONE pass reads the
// logits (4 or 2 B per element) and produces the focal sum and / or the decoder's keys.
//
// In NCHW a row (image, location, anchor) has its C classes at stride H*W, and the same class of
// neighbouring locations is contiguous: a thread owns one row and loops over its classes, a warp's
// loads are coalesced 128-byte lines, and the arg-max needs no cross-thread step at all.
//
// Arithmetic.
//  * Decoder score / class: exactly the reference's.  sigmoid is monotonic, so the scan tracks the
//    two largest LOGITS; the exact float32 sigmoid, 1 / (1 + expf(-x)) (bit-identical to torch's
//    CUDA kernel, tests/test_gpu_heads.py), is evaluated for those two only.  If they are more
//    than 4 ulp apart in probability space the winner is the reference's arg-max and its probability
//    the reference's score; otherwise (ties after rounding, saturation at 1.0) the row is rescanned
//    with the exact sigmoid of every class and np.argmax's first-maximum rule.
//  * Focal loss (tolerance 1e-5, north_star): background term as a function of t = e^x.  With
//    p = t / (1 + t):  p^2 * -log(1 - p) = t^3 * g(t),  g(t) = log(1 + t) / (t * (1 + t)^2), a
//    degree-7 polynomial on t <= 1/3 (p <= 0.25), max relative error 3.4e-7 in float32.  The clamp
//    p >= 1e-4 becomes t >= 1e-4 / (1 - 1e-4).  One MUFU.EX2 per element, no reciprocal; two
//    classes per FFMA2.  Target classes, p > 0.25, gamma != 2 take the exact-form functions of
//    focal_terms.cuh on the exact sigmoid.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "common.cuh"
#include "focal_terms.cuh"

namespace b200det {

constexpr int kLgThreads = 128;
constexpr int kLgMinCtas = 8;           // 64 registers: 1024 threads x 16 loads in flight per SM
constexpr int kLgUnroll = 16;            // loads in flight per thread
// The probability clamp [1e-4, .] and the polynomial's range p <= 0.25 as clamps of y = x * log2(e):
// t = 2^y in [1e-4 / (1 - 1e-4), 1/3]
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kYLo = -13.287568f;      // log2(1.00010001e-4)
constexpr float kYHi = -1.5849625f;      // log2(1/3)
constexpr float kXHi = -1.0986123f;      // ln(1/3): logits above take the exact-form path

struct LogitsArgs {
    PtrTab cls;                      // [B, A*C, HW] logits per level
    PtrTab ctr;                      // FCOS centre-ness logits [B, HW] (float32) or null
    long long row_base[kMaxLevels];  // level-major row base (B * off_l)
    int hw[kMaxLevels];              // H * W
    int tiles[kMaxLevels];           // kLgThreads-wide tiles per image
    int block_off[kMaxLevels + 1];
    int n_levels, A, C;
    const int *labels;               // level-major labels or null (every row is background)
    float alpha, gamma, min_score, thr_lo;
    long long *focal_slots;
    uint32_t *keys;
    int *classes;
};

template <typename T>
__device__ __forceinline__ float lg_load(const T *p);
template <>
__device__ __forceinline__ float lg_load<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float lg_load<__half>(const __half *p) { return __half2float(__ldg(p)); }
template <>
__device__ __forceinline__ float lg_load<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __bfloat162float(__ushort_as_bfloat16(__ldg(reinterpret_cast<const unsigned short *>(p))));
}

// torch's CUDA sigmoid for float32 (accurate expf, IEEE division)
__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + expf(-x)); }

// t^3 * g(t) for two classes (packed FP32), accumulated into acc
__device__ __forceinline__ float2 bg_term2_acc(float2 t, float2 acc) {
#define B200DET_C2(v) make_float2(v, v)
    float2 s = B200DET_C2(-4.1271257400512695f);
    s = __ffma2_rn(s, t, B200DET_C2(8.810138702392578f));
    s = __ffma2_rn(s, t, B200DET_C2(-9.959955215454102f));
    s = __ffma2_rn(s, t, B200DET_C2(8.53051471710205f));
    s = __ffma2_rn(s, t, B200DET_C2(-6.403131008148193f));
    s = __ffma2_rn(s, t, B200DET_C2(4.33279275894165f));
    s = __ffma2_rn(s, t, B200DET_C2(-2.4999916553497314f));
    s = __ffma2_rn(s, t, B200DET_C2(1.0f));
#undef B200DET_C2
    const float2 t3 = __fmul2_rn(__fmul2_rn(t, t), t);
    return __ffma2_rn(t3, s, acc);
}

// 2^y with one MUFU.EX2
__device__ __forceinline__ float ex2_fast(float y) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
    return r;
}
// t = e^x clamped to the polynomial's range; NaN stays NaN (and poisons the sum like the reference's)
__device__ __forceinline__ float lg_t_of_y(float y) {
    return ex2_fast(fmin_nan(fmax_nan(y, kYLo), kYHi));
}

// background term of one class outside the packed fast path (p > 0.25, gamma != 2): exact form
__device__ __noinline__ float lg_slow_bg(float x, float gamma, bool gamma2) {
    return neg_term_slow(sigmoid_exact(x), gamma, gamma2);
}

// the background term the sweep adds for logit x, WITHOUT (1 - alpha): fast polynomial or exact form
__device__ __forceinline__ float lg_bg_term(float x, float gamma, bool gamma2) {
    if (gamma2 && !(x > kXHi))
        return bg_term2_acc(make_float2(lg_t_of_y(x * kLog2e), 0.f), make_float2(0.f, 0.f)).x;
    return lg_slow_bg(x, gamma, gamma2);
}

// p + k * stride (bytes) as ONE 64-bit multiply-add (IMAD.WIDE.U32), independent of the other loads
template <typename T>
__device__ __forceinline__ const T *lg_plane(const T *p, unsigned stride_bytes, unsigned k) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(stride_bytes), "r"(k),
        "l"(reinterpret_cast<unsigned long long>(p)));
    return reinterpret_cast<const T *>(r);
}

// One CTA owns kLgThreads consecutive locations of one image and level, and ALL their anchors: the
// A * kLgThreads rows it produces are contiguous in the level-major row order, so labels are read
// and keys / classes written as full coalesced runs through shared memory.  (With one anchor per
// CTA every 32-byte sector of keys / classes was written by 9 different CTAs at different times:
// ncu showed 8x the key bytes in DRAM writes and as many fill reads.)
template <typename T, bool FOCAL, bool ARGMAX>
__global__ void __launch_bounds__(kLgThreads, kLgMinCtas)
    logits_sweep_kernel(LogitsArgs a) {
    // parked winning groups | [kLgThreads * A] labels, then keys | [kLgThreads * A] classes
    extern __shared__ __align__(16) int lg_smem[];
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i)
        if (i < a.n_levels && (int)blockIdx.x >= a.block_off[i]) l = i;
    const int rel = blockIdx.x - a.block_off[l];
    const int b = rel / a.tiles[l];
    const int hw0 = (rel - b * a.tiles[l]) * kLgThreads;
    const int hw = hw0 + threadIdx.x;
    const int HW = a.hw[l], C = a.C, A = a.A;
    const bool live = hw < HW;
    float4 *s_win = reinterpret_cast<float4 *>(lg_smem);   // [kLgUnroll / 4][kLgThreads]
    int *s_key = lg_smem + kLgUnroll * kLgThreads;
    int *s_cls = s_key + kLgThreads * A;
    const long long row0 = a.row_base[l] + ((long long)b * HW + hw0) * A;   // first row of the tile
    const int tile_rows = min(kLgThreads, HW - hw0) * A;
    if (FOCAL && a.labels) {
        for (int i = threadIdx.x; i < tile_rows; i += kLgThreads) s_key[i] = __ldg(a.labels + row0 + i);
        __syncthreads();
    }

    const bool gamma2 = a.gamma == 2.f;
    const float one_m_alpha = 1.f - a.alpha;
    const float ninf = -__int_as_float(0x7f800000);
    const int off0 = live ? (int)threadIdx.x : 0;
    const unsigned plane_bytes = (unsigned)HW * (unsigned)sizeof(T);
    float cta_total = 0.f;
    float cp = 0.f;   // FCOS centre-ness probability of this location
    if (ARGMAX && live && a.ctr.p[l])
        cp = sigmoid_exact(__ldg(static_cast<const float *>(a.ctr.p[l]) + (size_t)b * HW + hw));

    for (int anchor = 0; anchor < A; ++anchor) {
        // CTA-uniform base of this (image, anchor, tile) block of C class planes + a 32-bit element
        // offset (C * HW < 2^31 is checked on the host): one integer add and one address per load
        const T *__restrict__ src =
            static_cast<const T *>(a.cls.p[l]) + ((size_t)b * A + anchor) * C * HW + hw0 + off0;
        const T *ptr = src;   // walks the class planes: one 64-bit add per load
        const int slot = threadIdx.x * A + anchor;   // odd A: conflict-free; even A: 2-way at worst

        int target = -1;        // class index of the row's target, -1: none
        bool counted = live;    // ignored rows take no part in the focal loss (losses.py:228-230)
        if (FOCAL && live && a.labels) {
            const int label = s_key[slot];
            counted = label >= 0;
            target = label - 1;
        }
        // arg-max bookkeeping per GROUP of kLgUnroll classes (one max tree + 5 instructions per
        // group instead of 5 per class): m1 / m2 are the two largest group maxima, ig the first
        // group that holds the maximum; the winning group is rescanned after the loop
        float m1 = ninf, m2 = ninf;
        int ig = 0;
        bool any_nan = false;   // a NaN logit: np.argmax stops there, the row's score is NaN
        float2 acc2 = make_float2(0.f, 0.f);
        float acc_slow = 0.f;   // exact-form background terms, without (1 - alpha)

        // one group of kLgUnroll classes; FULL: no class-count predicate.  Dead threads of a ragged
        // tile read location 0 of the tile (off0) and discard the result: no load is predicated
        auto group = [&](int c0, auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            float x[kLgUnroll];
#pragma unroll
            for (int k = 0; k < kLgUnroll; ++k) {
                x[k] = ninf;
                if (FULL || c0 + k < C) x[k] = lg_load<T>(lg_plane(ptr, plane_bytes, k));
            }
            ptr = lg_plane(ptr, plane_bytes, kLgUnroll);
            float g = x[0];   // largest logit of the group; NaN if the group holds one (max.NaN)
#pragma unroll
            for (int k = 1; k < kLgUnroll; ++k) g = fmax_nan(g, x[k]);
            if (ARGMAX) {
                if (g != g) {
                    any_nan = true;
                    g = ninf;
                }
                m2 = fmaxf(m2, fminf(m1, g));
                if (g > m1) {   // strict: the first group that reaches the maximum; park its logits
                    ig = c0;
#pragma unroll
                    for (int k = 0; k < kLgUnroll; k += 4)
                        s_win[(k / 4) * kLgThreads + threadIdx.x] =
                            make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
                }
                m1 = fmaxf(m1, g);
            }
            if (FOCAL && counted) {
                // every class as background here; the target class is corrected after the loop
                float t[kLgUnroll];
#pragma unroll
                for (int k = 0; k < kLgUnroll; k += 2) {
                    const float2 y = __fmul2_rn(make_float2(x[k], x[k + 1]), make_float2(kLog2e, kLog2e));
                    t[k] = lg_t_of_y(y.x);       // padding (x = -inf) gives 2^kYLo
                    t[k + 1] = lg_t_of_y(y.y);
                }
#pragma unroll
                for (int k = 0; k < kLgUnroll; k += 2)
                    acc2 = bg_term2_acc(make_float2(t[k], t[k + 1]), acc2);
                if (!FULL) {
                    // remove the padding classes' terms again
#pragma unroll
                    for (int k = 0; k < kLgUnroll; ++k)
                        if (c0 + k >= C)
                            acc2.x -= bg_term2_acc(make_float2(t[k], 0.f), make_float2(0.f, 0.f)).x;
                }
                if (g > kXHi || !gamma2) {   // rare: some class is outside the polynomial's range
#pragma unroll
                    for (int k = 0; k < kLgUnroll; ++k) {
                        if ((FULL || c0 + k < C) && (x[k] > kXHi || !gamma2)) {
                            acc2.x -= bg_term2_acc(make_float2(t[k], 0.f), make_float2(0.f, 0.f)).x;
                            acc_slow += lg_slow_bg(x[k], a.gamma, gamma2);
                        }
                    }
                }
            }
        };
        int c0 = 0;
        for (; c0 + kLgUnroll <= C; c0 += kLgUnroll) group(c0, std::true_type());
        if (c0 < C) group(c0, std::false_type());

        if (FOCAL) {
            float total = one_m_alpha * ((acc2.x + acc2.y) + acc_slow);
            if (counted && target >= 0 && target < C) {
                // the target class was counted as background: swap in the positive term
                const float xt = lg_load<T>(src + (size_t)target * HW);
                total += a.alpha * pos_term(sigmoid_exact(xt), a.gamma, gamma2) -
                         one_m_alpha * lg_bg_term(xt, a.gamma, gamma2);
            }
            cta_total += total;
        }

        if (ARGMAX && live) {
            float p1 = sigmoid_exact(m1);
            int i1 = ig;
            if (p1 > a.thr_lo) {
                // the parked winning group: first class equal to the maximum, and the largest of
                // the group's other classes -> second-largest logit of the whole row
                float w[kLgUnroll];
#pragma unroll
                for (int k = 0; k < kLgUnroll; k += 4) {
                    const float4 v = s_win[(k / 4) * kLgThreads + threadIdx.x];
                    w[k] = v.x, w[k + 1] = v.y, w[k + 2] = v.z, w[k + 3] = v.w;
                }
                if (m1 != ninf) {   // a row of -inf never parks a group: class 0, p1 == p2 == 0
                    bool found = false;
#pragma unroll
                    for (int k = 0; k < kLgUnroll; ++k) {
                        if (!found && w[k] == m1) {
                            found = true;
                            i1 = ig + k;
                        } else {
                            m2 = fmaxf(m2, w[k]);
                        }
                    }
                } else {
                    m2 = m1;
                }
                const float p2 = sigmoid_exact(m2);
                if ((int)(__float_as_uint(p1) - __float_as_uint(p2)) <= 4) {
                    // the two best classes are (almost) tied in probability space: exact rescan
                    float best = ninf;
                    int bi = 0;
                    for (int c = 0; c < C; ++c) {
                        const float pe = sigmoid_exact(lg_load<T>(src + (size_t)c * HW));
                        if (pe > best) {
                            best = pe;
                            bi = c;
                        }
                    }
                    p1 = best;
                    i1 = bi;
                }
            }
            float score = p1;
            // np.sqrt(cls_scores * center_preds)  (decode.py:338) on the exact probabilities
            if (a.ctr.p[l]) score = __fsqrt_rn(__fmul_rn(p1, cp));
            // sigmoid(NaN) is NaN and np.argmax returns the first NaN (decode.py:230-238): the
            // row's score is NaN and fails the threshold whatever the other classes hold
            if (any_nan) score = __int_as_float(0x7fc00000);
            s_key[slot] = (int)((score > a.min_score) ? ((__float_as_uint(score) & 0x80000000u)
                                                             ? ~__float_as_uint(score)
                                                             : (__float_as_uint(score) | 0x80000000u))
                                                      : 0u);   // strict '>' (decode.py:133-138)
            s_cls[slot] = i1;
        }
    }

    if (ARGMAX) {
        __syncthreads();
        for (int i = threadIdx.x; i < tile_rows; i += kLgThreads) {
            a.keys[row0 + i] = (uint32_t)s_key[i];
            a.classes[row0 + i] = s_cls[i];
        }
    }
    if (FOCAL) sweep_accumulate<kLgThreads>(cta_total, a.focal_slots);
}

}  // namespace b200det

using namespace b200det;

// One sweep over the NCHW classification logits of every level.
//   labels      level-major labels from b200det_*_assign (NULL: every row is background)
//   focal_ws    loss workspace (sweep accumulators), NULL: no focal sum
//   keys/classes  decoder keys (NULL: no arg-max)
extern "C" int b200det_logits_sweep(const b200det_geometry *geo, const void *const *cls_logits,
                                    int cls_dtype, const void *const *ctr_logits,
                                    const int32_t *labels, float alpha, float gamma,
                                    void *loss_workspace, size_t loss_workspace_bytes,
                                    float min_score, uint32_t *keys, int32_t *classes,
                                    void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!cls_logits) return B200DET_EINVAL;
    if (cls_dtype != B200DET_F32 && cls_dtype != B200DET_F16 && cls_dtype != B200DET_BF16)
        return B200DET_EINVAL;
    const bool focal = loss_workspace != nullptr;
    const bool argmax = keys != nullptr;
    if (!focal && !argmax) return B200DET_EINVAL;
    if (argmax && !classes) return B200DET_EINVAL;
    LogitsArgs a;
    a.n_levels = g.n_levels;
    a.A = g.per_loc;
    a.C = g.num_classes;
    a.labels = labels;
    a.alpha = alpha;
    a.gamma = gamma;
    a.min_score = min_score;
    // rows whose best probability cannot pass the threshold need no exact tie handling
    a.thr_lo = ctr_logits ? min_score * min_score * 0.99999f : min_score * 0.999999f;
    if (!(min_score > 0.f)) a.thr_lo = -1.f;
    a.focal_slots = nullptr;
    a.keys = keys;
    a.classes = classes;
    if (focal) {
        const LossWs ws = loss_ws_layout(g);
        if (loss_workspace_bytes < ws.total) return B200DET_EWORKSPACE;
        a.focal_slots = reinterpret_cast<long long *>(static_cast<char *>(loss_workspace) + ws.off_focal);
        if (!g_skip_memset) {
            cudaError_t e = cudaMemsetAsync(a.focal_slots, 0, kSweepWords * sizeof(long long),
                                            (cudaStream_t)stream);
            if (e != cudaSuccess) return (int)e;
        }
    }
    int blocks = 0;
    const uintptr_t amask = cls_dtype == B200DET_F32 ? 3 : 1;
    for (int l = 0; l < kMaxLevels; ++l) {
        a.cls.p[l] = a.ctr.p[l] = nullptr;
        a.row_base[l] = 0;
        a.hw[l] = a.tiles[l] = 0;
    }
    for (int l = 0; l < g.n_levels; ++l) {
        if (!cls_logits[l] || (ctr_logits && !ctr_logits[l])) return B200DET_EINVAL;
        if (reinterpret_cast<uintptr_t>(cls_logits[l]) & amask) return B200DET_EALIGN;
        a.cls.p[l] = cls_logits[l];
        a.ctr.p[l] = ctr_logits ? ctr_logits[l] : nullptr;
        a.row_base[l] = (long long)g.batch * g.off[l];
        a.hw[l] = g.H[l] * g.W[l];
        a.tiles[l] = (a.hw[l] + kLgThreads - 1) / kLgThreads;
        if ((long long)g.num_classes * a.hw[l] + a.hw[l] >= (1ll << 31)) return B200DET_ERANGE;
        a.block_off[l] = blocks;
        const long long nb = (long long)g.batch * a.tiles[l];
        if (blocks + nb > 0x7fffffffLL) return B200DET_ERANGE;
        blocks += (int)nb;
    }
    for (int l = g.n_levels; l <= kMaxLevels; ++l) a.block_off[l] = blocks;
    const unsigned grid = (unsigned)blocks;
    const size_t smem = sizeof(int) * kLgThreads * (2 * (size_t)g.per_loc + kLgUnroll);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(kKernLogits, stream);
#define B200DET_LG(T)                                                                         \
    do {                                                                                      \
        if (focal && argmax) logits_sweep_kernel<T, true, true><<<grid, kLgThreads, smem, st>>>(a);   \
        else if (focal) logits_sweep_kernel<T, true, false><<<grid, kLgThreads, smem, st>>>(a);       \
        else logits_sweep_kernel<T, false, true><<<grid, kLgThreads, smem, st>>>(a);                  \
    } while (0)
    if (cls_dtype == B200DET_F32) B200DET_LG(float);
    else if (cls_dtype == B200DET_F16) B200DET_LG(__half);
    else B200DET_LG(__nv_bfloat16);
#undef B200DET_LG
    count_launch();
    return (int)cudaGetLastError();
}

// Evaluation step from logits: assignment -> box (and centre-ness) loss of the positives -> ONE
// sweep over the classification logits (label-aware focal sum + decoder keys) -> reduce / finish ->
// select + NMS.  Argument meaning as b200det_eval_step; cls are NCHW logits [B, A*C, H, W] of dtype
// cls_dtype; FCOS heads also pass their float32 centre-ness logits [B, 1, H, W].
extern "C" int b200det_logits_eval_step(const b200det_geometry *geo, const b200det_loss_params *lp,
                                        const b200det_decode_params *dp, const float *annotations,
                                        int max_gt, const void *const *cls_logits, int cls_dtype,
                                        const void *const *reg, const void *const *ctr_logits,
                                        int32_t *labels, void *loss_workspace,
                                        size_t loss_workspace_bytes, double *sums, float *losses,
                                        uint32_t *keys, int32_t *classes, float *out,
                                        void *decode_workspace, size_t decode_workspace_bytes,
                                        void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!lp || !dp || !annotations || !cls_logits || !reg || !labels || !loss_workspace || !sums ||
        !keys || !classes || !out)
        return B200DET_EINVAL;
    const bool fcos = lp->is_fcos != 0;
    if ((dp->is_fcos != 0) != fcos || (fcos && !ctr_logits)) return B200DET_EINVAL;
    const LossWs ws = loss_ws_layout(g);
    if (loss_workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    char *base = static_cast<char *>(loss_workspace);
    cudaError_t e = cudaMemsetAsync(base + ws.off_focal, 0,
                                    ws.off_counters + 2 * sizeof(int) - ws.off_focal,
                                    (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    g_skip_memset = true;
    rc = fcos ? b200det_fcos_assign(geo, annotations, max_gt, lp->use_center_sample, labels, nullptr,
                                    nullptr, loss_workspace, loss_workspace_bytes, stream)
              : b200det_retina_assign(geo, annotations, max_gt, lp->iou_neg, lp->iou_pos, labels,
                                      nullptr, loss_workspace, loss_workspace_bytes, stream);
    if (!rc)   // losses of the positives only: the sweep below is label-aware, nothing to correct
        rc = b200det_sparse_losses(geo, fcos ? (1 | B200DET_FCOS_CTR_LOGITS) : 0, annotations,
                                   max_gt, labels, reg, lp->reg_dtype, fcos ? ctr_logits : nullptr,
                                   lp->box_loss, lp->beta, nullptr, lp->alpha, lp->gamma, nullptr,
                                   nullptr, loss_workspace, loss_workspace_bytes, stream);
    if (!rc)
        rc = b200det_logits_sweep(geo, cls_logits, cls_dtype, fcos ? ctr_logits : nullptr, labels,
                                  lp->alpha, lp->gamma, loss_workspace, loss_workspace_bytes,
                                  dp->min_score, keys, classes, stream);
    g_skip_memset = false;
    if (!rc)   // reduction and normalisation in one launch unless the caller all-reduces in between
        rc = losses ? b200det_loss_reduce_finish(geo, loss_workspace, loss_workspace_bytes, lp->w_cls,
                                                 lp->w_box, lp->w_ctr, sums, losses, stream)
                    : b200det_loss_reduce(geo, 3, loss_workspace, loss_workspace_bytes, sums, stream);
    if (!rc)
        rc = select_decode_nms_impl(geo, keys, classes, reg, dp->reg_dtype, fcos ? 1 : 0,
                                    dp->min_score, dp->topn, dp->max_out, dp->nms_type,
                                    dp->nms_threshold, dp->scales, dp->sizes, dp->to_xywh, out,
                                    nullptr, nullptr, nullptr, dp->half_exp_table, stream);
    return rc;
}
