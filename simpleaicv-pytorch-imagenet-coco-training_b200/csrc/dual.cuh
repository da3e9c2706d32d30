// dual.cuh -- forward-mode dual numbers (value + 4 partials) used to differentiate the
// IoU-family box losses w.r.t. the four regression outputs of one positive row, following
// torch autograd's conventions for the reference expression (losses.py:54-123):
//   * elementwise max/min split the gradient 0.5/0.5 on ties,
//   * clamp(min=c) passes the gradient where x >= c (inclusive),
//   * CIoU's alpha is a constant (torch.no_grad, losses.py:104-105).
// Only positives (a few thousand rows per batch) evaluate this, so clarity beats speed.
//   * max / min / clamp propagate NaN like torch does (fmax_nan / fmin_nan), so a NaN regression
//     output yields a NaN loss instead of silently disappearing.
#pragma once
#include <cuda_runtime.h>

#include "focal_terms.cuh"

namespace b200det {

struct Dual {
    float v;
    float d[4];
};

__device__ __forceinline__ Dual dconst(float v) {
    Dual r;
    r.v = v;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = 0.f;
    return r;
}
__device__ __forceinline__ Dual dvar(float v, int k) {
    Dual r = dconst(v);
    r.d[k] = 1.f;
    return r;
}
__device__ __forceinline__ Dual operator+(const Dual &a, const Dual &b) {
    Dual r;
    r.v = __fadd_rn(a.v, b.v);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
__device__ __forceinline__ Dual operator-(const Dual &a, const Dual &b) {
    Dual r;
    r.v = __fsub_rn(a.v, b.v);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
__device__ __forceinline__ Dual operator+(const Dual &a, float b) {
    Dual r = a;
    r.v = __fadd_rn(a.v, b);
    return r;
}
__device__ __forceinline__ Dual operator-(const Dual &a, float b) {
    Dual r = a;
    r.v = __fsub_rn(a.v, b);
    return r;
}
__device__ __forceinline__ Dual operator-(float a, const Dual &b) {
    Dual r;
    r.v = __fsub_rn(a, b.v);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = -b.d[i];
    return r;
}
__device__ __forceinline__ Dual operator*(const Dual &a, float b) {
    Dual r;
    r.v = __fmul_rn(a.v, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] * b;
    return r;
}
__device__ __forceinline__ Dual operator*(const Dual &a, const Dual &b) {
    Dual r;
    r.v = __fmul_rn(a.v, b.v);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
__device__ __forceinline__ Dual operator/(const Dual &a, const Dual &b) {
    Dual r;
    r.v = __fdiv_rn(a.v, b.v);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
    return r;
}
__device__ __forceinline__ Dual operator/(const Dual &a, float b) {
    Dual r;
    r.v = __fdiv_rn(a.v, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] / b;
    return r;
}
__device__ __forceinline__ Dual dsq(const Dual &a) {
    Dual r;
    r.v = __fmul_rn(a.v, a.v);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = 2.f * a.v * a.d[i];
    return r;
}
// mode: reg_dtype incl. B200DET_REG_EXP_ROUNDED (eager half arithmetic: the result of exp is rounded
// to the tensor's precision and autograd multiplies by that rounded result)
__device__ __forceinline__ Dual dexp(const Dual &a, int mode = 0) {
    Dual r;
    r.v = round_like(expf(a.v), mode);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = r.v * a.d[i];
    return r;
}
__device__ __forceinline__ Dual datan(const Dual &a) {
    Dual r;
    r.v = atanf(a.v);
    const float s = 1.f / (1.f + a.v * a.v);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = s * a.d[i];
    return r;
}
// max / min against a constant
__device__ __forceinline__ Dual dmax(const Dual &a, float b) {
    Dual r;
    r.v = fmax_nan(a.v, b);
    const float w = a.v > b ? 1.f : (a.v == b ? 0.5f : 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = w * a.d[i];
    return r;
}
__device__ __forceinline__ Dual dmin(const Dual &a, float b) {
    Dual r;
    r.v = fmin_nan(a.v, b);
    const float w = a.v < b ? 1.f : (a.v == b ? 0.5f : 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = w * a.d[i];
    return r;
}
// torch.clamp(x, min=c): gradient passes where x >= c
__device__ __forceinline__ Dual dclamp_min(const Dual &a, float c) {
    Dual r;
    r.v = fmax_nan(a.v, c);
    const float w = a.v >= c ? 1.f : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = w * a.d[i];
    return r;
}

// IoU family between a predicted (differentiable) box p[4] and a constant box g[4].
// Operation order of IoUMethod.__call__ (losses.py:54-123).
__device__ __forceinline__ Dual iou_family(const Dual p[4], const float g[4], int type) {
    const Dual ltx = dmax(p[0], g[0]), lty = dmax(p[1], g[1]);
    const Dual rbx = dmin(p[2], g[2]), rby = dmin(p[3], g[3]);
    const Dual iw = dclamp_min(rbx - ltx, 0.f), ih = dclamp_min(rby - lty, 0.f);
    const Dual inter = iw * ih;
    const Dual w1 = dclamp_min(p[2] - p[0], 0.f), h1 = dclamp_min(p[3] - p[1], 0.f);
    const float w2 = fmaxf(__fsub_rn(g[2], g[0]), 0.f), h2 = fmaxf(__fsub_rn(g[3], g[1]), 0.f);
    const Dual area1 = w1 * h1;
    const float area2 = __fmul_rn(w2, h2);
    const Dual uni = dclamp_min((area1 + area2) - inter, 1e-4f);
    const Dual iou = inter / uni;
    if (type == B200DET_BOX_IOU) return iou;
    const Dual ex1 = dmin(p[0], g[0]), ey1 = dmin(p[1], g[1]);
    const Dual ex2 = dmax(p[2], g[2]), ey2 = dmax(p[3], g[3]);
    const Dual ew = dclamp_min(ex2 - ex1, 0.f), eh = dclamp_min(ey2 - ey1, 0.f);
    if (type == B200DET_BOX_GIOU) {
        const Dual enc = dclamp_min(ew * eh, 1e-4f);
        return iou - (enc - uni) / enc;
    }
    const Dual c2 = dclamp_min(dsq(ew) + dsq(eh), 1e-4f);
    const Dual pcx = (p[2] + p[0]) / 2.f, pcy = (p[3] + p[1]) / 2.f;
    const float gcx = __fdiv_rn(__fadd_rn(g[2], g[0]), 2.f), gcy = __fdiv_rn(__fadd_rn(g[3], g[1]), 2.f);
    const Dual p2 = dsq(pcx - gcx) + dsq(pcy - gcy);
    if (type == B200DET_BOX_DIOU) return iou - p2 / c2;
    if (type == B200DET_BOX_CIOU) {
        const float k = 0.40528473456935109f;  // float32(4 / pi^2)
        const Dual dat = atanf(__fdiv_rn(w2, h2)) - datan(w1 / h1);
        const Dual v = dsq(dat) * k;
        const float alpha = __fdiv_rn(v.v, fmaxf(__fadd_rn(__fsub_rn(1.f, iou.v), v.v), 1e-4f));
        return iou - (p2 / c2 + v * alpha);
    }
    // EIoU
    const Dual pw2 = dsq(w2 - w1), ph2 = dsq(h2 - h1);
    const Dual cw2 = dclamp_min(dsq(ew), 1e-4f), ch2 = dclamp_min(dsq(eh), 1e-4f);
    return iou - (p2 / c2 + pw2 / cw2 + ph2 / ch2);
}

}  // namespace b200det
