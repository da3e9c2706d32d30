// iou.cu -- IoUMethod.__call__ as a stand-alone operator (reference
// simpleAICV/detection/losses.py:28-123): IoU / GIoU / DIoU / CIoU / EIoU between two box sets,
// xyxy or xywh, element-wise or as the [N,1,4] x [1,M,4] broadcast of the assignment
// (losses.py:350-353), with the Jacobians autograd needs.
//
// The assignment and loss kernels carry their own copies of this arithmetic (assign.cu,
// dual.cuh: one box constant); this file is the operator a caller of `IoUMethod()` gets.  Both
// boxes are dual numbers here so that NaN coordinates of either side propagate like torch's
// max / min / clamp do, and the same function yields d/d(boxes1) and d/d(boxes2).
// Compiled with -fmad=false: every value is one correctly rounded float32 op per reference op.
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"
#include "dual.cuh"

namespace b200det {

// torch.max / torch.min of two tensors: NaN propagates, ties split the gradient 0.5 / 0.5
__device__ __forceinline__ Dual dmax2(const Dual &a, const Dual &b) {
    Dual r;
    r.v = fmax_nan(a.v, b.v);
    const float wa = a.v > b.v ? 1.f : (a.v == b.v ? 0.5f : 0.f);
    const float wb = a.v < b.v ? 1.f : (a.v == b.v ? 0.5f : 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = wa * a.d[i] + wb * b.d[i];
    return r;
}
__device__ __forceinline__ Dual dmin2(const Dual &a, const Dual &b) {
    Dual r;
    r.v = fmin_nan(a.v, b.v);
    const float wa = a.v < b.v ? 1.f : (a.v == b.v ? 0.5f : 0.f);
    const float wb = a.v > b.v ? 1.f : (a.v == b.v ? 0.5f : 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) r.d[i] = wa * a.d[i] + wb * b.d[i];
    return r;
}

// losses.py:54-123, op for op; a = boxes1, b = boxes2 (xyxy)
__device__ __forceinline__ Dual iou_method(const Dual a[4], const Dual b[4], int type) {
    const Dual ltx = dmax2(a[0], b[0]), lty = dmax2(a[1], b[1]);
    const Dual rbx = dmin2(a[2], b[2]), rby = dmin2(a[3], b[3]);
    const Dual iw = dclamp_min(rbx - ltx, 0.f), ih = dclamp_min(rby - lty, 0.f);
    const Dual inter = iw * ih;
    const Dual w1 = dclamp_min(a[2] - a[0], 0.f), h1 = dclamp_min(a[3] - a[1], 0.f);
    const Dual w2 = dclamp_min(b[2] - b[0], 0.f), h2 = dclamp_min(b[3] - b[1], 0.f);
    const Dual uni = dclamp_min((w1 * h1 + w2 * h2) - inter, 1e-4f);
    const Dual iou = inter / uni;
    if (type == B200DET_BOX_IOU) return iou;
    const Dual ex1 = dmin2(a[0], b[0]), ey1 = dmin2(a[1], b[1]);
    const Dual ex2 = dmax2(a[2], b[2]), ey2 = dmax2(a[3], b[3]);
    const Dual ew = dclamp_min(ex2 - ex1, 0.f), eh = dclamp_min(ey2 - ey1, 0.f);
    if (type == B200DET_BOX_GIOU) {
        const Dual enc = dclamp_min(ew * eh, 1e-4f);
        return iou - (enc - uni) / enc;
    }
    const Dual c2 = dclamp_min(dsq(ew) + dsq(eh), 1e-4f);
    const Dual c1x = (a[2] + a[0]) / 2.f, c1y = (a[3] + a[1]) / 2.f;
    const Dual c2x = (b[2] + b[0]) / 2.f, c2y = (b[3] + b[1]) / 2.f;
    const Dual p2 = dsq(c1x - c2x) + dsq(c1y - c2y);
    if (type == B200DET_BOX_DIOU) return iou - p2 / c2;
    if (type == B200DET_BOX_CIOU) {
        const float k = 0.40528473456935109f;  // float32(4 / pi^2)
        const Dual v = dsq(datan(w2 / h2) - datan(w1 / h1)) * k;
        // alpha is computed under torch.no_grad (losses.py:104-105): a constant
        const float alpha = __fdiv_rn(v.v, fmax_nan(__fadd_rn(__fsub_rn(1.f, iou.v), v.v), 1e-4f));
        return iou - (p2 / c2 + v * alpha);
    }
    const Dual pw2 = dsq(w2 - w1), ph2 = dsq(h2 - h1);
    const Dual cw2 = dclamp_min(dsq(ew), 1e-4f), ch2 = dclamp_min(dsq(eh), 1e-4f);
    return iou - (p2 / c2 + pw2 / cw2 + ph2 / ch2);
}

// the four inputs of one box as duals; `var`: they are the differentiation variables
__device__ __forceinline__ void load_box(const float *__restrict__ p, bool xywh, bool var, Dual out[4]) {
    Dual in[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) in[k] = var ? dvar(__ldg(p + k), k) : dconst(__ldg(p + k));
    if (!xywh) {
#pragma unroll
        for (int k = 0; k < 4; ++k) out[k] = in[k];
        return;
    }
    // losses.py:44-52: x1y1 = ctr - wh / 2, x2y2 = ctr + wh / 2
    const Dual hw = in[2] / 2.f, hh = in[3] / 2.f;
    out[0] = in[0] - hw;
    out[1] = in[1] - hh;
    out[2] = in[0] + hw;
    out[3] = in[1] + hh;
}

struct IouArgs {
    const float *b1, *b2;
    long long s1n, s1m, s2n, s2m, n, m;
    int type, xywh;
    float *out, *jac1, *jac2;
};

__global__ void __launch_bounds__(128) iou_method_kernel(IouArgs a) {
    const long long total = a.n * a.m;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / a.m, j = e - i * a.m;
        const float *p1 = a.b1 + 4 * (i * a.s1n + j * a.s1m);
        const float *p2 = a.b2 + 4 * (i * a.s2n + j * a.s2m);
        Dual x[4], y[4];
        load_box(p1, a.xywh, true, x);
        load_box(p2, a.xywh, false, y);
        const Dual r = iou_method(x, y, a.type);
        a.out[e] = r.v;
        if (a.jac1) reinterpret_cast<float4 *>(a.jac1)[e] = make_float4(r.d[0], r.d[1], r.d[2], r.d[3]);
        if (a.jac2) {
            load_box(p1, a.xywh, false, x);
            load_box(p2, a.xywh, true, y);
            const Dual q = iou_method(x, y, a.type);
            reinterpret_cast<float4 *>(a.jac2)[e] = make_float4(q.d[0], q.d[1], q.d[2], q.d[3]);
        }
    }
}

}  // namespace b200det

using namespace b200det;

extern "C" int b200det_iou_method(const float *boxes1, long long s1n, long long s1m,
                                  const float *boxes2, long long s2n, long long s2m, long long n,
                                  long long m, int iou_type, int xywh, float *out, float *jac1,
                                  float *jac2, void *stream) {
    if (n < 0 || m < 0) return B200DET_EINVAL;
    if (iou_type < B200DET_BOX_IOU || iou_type > B200DET_BOX_EIOU) return B200DET_EINVAL;
    if (n == 0 || m == 0) return 0;
    if (!boxes1 || !boxes2 || !out) return B200DET_EINVAL;
    if (s1n < 0 || s1m < 0 || s2n < 0 || s2m < 0) return B200DET_EINVAL;
    if (n > (1ll << 40) / m) return B200DET_ERANGE;
    if ((reinterpret_cast<uintptr_t>(boxes1) | reinterpret_cast<uintptr_t>(boxes2) |
         reinterpret_cast<uintptr_t>(out)) & 3)
        return B200DET_EALIGN;
    if ((reinterpret_cast<uintptr_t>(jac1) | reinterpret_cast<uintptr_t>(jac2)) & 15)
        return B200DET_EALIGN;
    IouArgs a{boxes1, boxes2, s1n, s1m, s2n, s2m, n, m, iou_type, xywh ? 1 : 0, out, jac1, jac2};
    long long blocks = (n * m + 127) / 128;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ProfScope prof(kKernOther, stream);
    iou_method_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    return (int)cudaGetLastError();
}
