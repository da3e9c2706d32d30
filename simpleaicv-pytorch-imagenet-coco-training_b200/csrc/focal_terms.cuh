// focal_terms.cuh -- per-element focal-loss terms shared by focal.cu (the streaming sweep) and
// assign.cu (corrections for positive / ignored rows).
//
// Reference (losses.py:245-259, identical at :532-546), p already clamped to [1e-4, 1-1e-4]:
//   target class   : alpha     * (1 - p)^gamma        * -log(p)
//   other classes  : (1-alpha) * (1 - (1 - p))^gamma  * -log(1 - p)
// The functions below return the terms WITHOUT the alpha factors.
#pragma once
#include <cuda_runtime.h>

namespace b200det {

constexpr float kClampLo = 1e-4f;    // float32(1e-4)      (losses.py:196, :493)
constexpr float kClampHi = 0.9999f;  // float32(1. - 1e-4)
constexpr float kFastMax = 0.25f;

// torch.clamp propagates NaN (fmaxf / fminf do not): a NaN probability must reach the loss value,
// because the caller's training loop skips steps whose loss is NaN (tools/scripts.py:922-930).
// max.NaN / min.NaN are single FMNMX.NAN instructions.
__device__ __forceinline__ float fmax_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float fmin_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float clamp_prob(float p) {
    return fmin_nan(fmax_nan(p, kClampLo), kClampHi);
}

// S(x) = -log(1 - x) / x on [0, 0.25]; Chebyshev-node fit, max rel err 1.3e-7 in float32 Horner
__device__ __forceinline__ float neg_log1m_over_x(float x) {
    float s = 0.3386436402797699f;
    s = fmaf(s, x, 0.14463114738464355f);
    s = fmaf(s, x, 0.2576442062854767f);
    s = fmaf(s, x, 0.33287033438682556f);
    s = fmaf(s, x, 0.500010073184967f);
    s = fmaf(s, x, 0.9999999403953552f);
    return s;
}

// background term for gamma == 2 and max(p, 1e-4) <= 0.25 (no upper clamp needed):
//   xr = 1 - (1 - x) is the reference's (1 - pt);  -log(1 - xr) = xr * S(xr) because 1 - xr == 1 - x
//   exactly (Sterbenz), so the polynomial sees the same rounded argument the reference's log does.
__device__ __forceinline__ float neg_term_fast(float x_clamped_lo, float &xr_out, float &xs_out) {
    const float q = 1.f - x_clamped_lo;
    const float xr = 1.f - q;
    const float xs = xr * neg_log1m_over_x(xr);  // -log(q)
    xr_out = xr;
    xs_out = xs;
    return (xr * xr) * xs;
}

// Two elements at a time with Blackwell's packed FP32 pipe (fma.rn.f32x2 -> FFMA2): same
// arithmetic as neg_term_fast, half the issue slots.  1 - x is computed as fma(x, -1, 1), which
// rounds once exactly like the subtraction.
__device__ __forceinline__ float2 neg_term_fast2_acc(float2 x_clamped_lo, float2 acc,
                                                     float2 &xr_out, float2 &xs_out) {
    const float2 one = make_float2(1.f, 1.f), neg1 = make_float2(-1.f, -1.f);
    const float2 q = __ffma2_rn(x_clamped_lo, neg1, one);
    const float2 xr = __ffma2_rn(q, neg1, one);
    float2 s = make_float2(0.3386436402797699f, 0.3386436402797699f);
    s = __ffma2_rn(s, xr, make_float2(0.14463114738464355f, 0.14463114738464355f));
    s = __ffma2_rn(s, xr, make_float2(0.2576442062854767f, 0.2576442062854767f));
    s = __ffma2_rn(s, xr, make_float2(0.33287033438682556f, 0.33287033438682556f));
    s = __ffma2_rn(s, xr, make_float2(0.500010073184967f, 0.500010073184967f));
    s = __ffma2_rn(s, xr, make_float2(0.9999999403953552f, 0.9999999403953552f));
    const float2 xs = __fmul2_rn(xr, s);
    xr_out = xr;
    xs_out = xs;
    return __ffma2_rn(__fmul2_rn(xr, xr), xs, acc);   // acc + xr^2 * (-log q)
}

// exact-form background term (any p, any gamma): accurate logf / powf
static __device__ __noinline__ float neg_term_slow(float p, float gamma, bool gamma2) {
    const float pc = clamp_prob(p);
    const float q = 1.f - pc;  // pt
    const float x = 1.f - q;   // 1 - pt
    const float w = gamma2 ? x * x : powf(x, gamma);
    return w * (-logf(q));
}

__device__ __forceinline__ float neg_term(float p, float gamma, bool gamma2) {
    const float x = fmax_nan(p, kClampLo);
    if (gamma2 && x <= kFastMax) {
        float a, b;
        return neg_term_fast(x, a, b);
    }
    return neg_term_slow(p, gamma, gamma2);
}

// target-class term
__device__ __forceinline__ float pos_term(float p, float gamma, bool gamma2) {
    const float pc = clamp_prob(p);
    const float om = 1.f - pc;  // 1 - pt
    const float w = gamma2 ? om * om : powf(om, gamma);
    return w * (-logf(pc));
}

// CTA sum -> one 64-bit fixed-point atomic into one of kSweepSlots accumulators.  Integer adds
// commute, so the total does not depend on CTA scheduling order (deterministic), and the final
// reduction reads kSweepSlots values instead of one partial per CTA (300k at batch 256).
// Scale 2^36: resolution 1.5e-11 per CTA sum (typical CTA sums are ~1e-2), capacity 1.3e8 per slot.
// slots[kSweepSlots] is a poison flag: set when a CTA sum is NaN / inf (the reduction then reports
// NaN like the reference's float sum would).
constexpr int kSweepSlots = 1024;
constexpr int kSweepWords = kSweepSlots + 1;
constexpr double kFxSweep = 68719476736.0;   // 2^36
template <int THREADS>
__device__ __forceinline__ void sweep_accumulate(float value, long long *slots) {
    __shared__ float sweep_red[THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float w = value;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane == 0) sweep_red[warp] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < THREADS / 32; ++i) s += sweep_red[i];
        if (!(fabsf(s) <= 3.402823466e38f)) {
            slots[kSweepSlots] = 1;
        } else {
            const long long fx = __double2ll_rn((double)s * kFxSweep);
            atomicAdd(reinterpret_cast<unsigned long long *>(slots + (blockIdx.x & (kSweepSlots - 1))),
                      (unsigned long long)fx);
        }
    }
}

// Warp-level variant: no shared memory, no barrier; 8x the atomics of the CTA version (still a few
// hundred thousand per launch spread over kSweepSlots addresses).
__device__ __forceinline__ void sweep_accumulate_warp(float value, long long *slots) {
    float w = value;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0) {
        if (!(fabsf(w) <= 3.402823466e38f)) {
            slots[kSweepSlots] = 1;
        } else {
            const long long fx = __double2ll_rn((double)w * kFxSweep);
            const unsigned slot = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (kSweepSlots - 1);
            atomicAdd(reinterpret_cast<unsigned long long *>(slots + slot), (unsigned long long)fx);
        }
    }
}

}  // namespace b200det
