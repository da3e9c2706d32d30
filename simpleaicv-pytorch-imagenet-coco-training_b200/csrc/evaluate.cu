// evaluate.cu -- the per-batch part of the reference's VOC evaluator:
//   compute_ious                         tools/scripts.py:487-508
//   detection <-> ground-truth matching  tools/scripts.py:626-651 (inside evaluate_voc_detection)
// The reference runs it on the host over the whole test set, one Python loop per (threshold, class,
// image, detection).  Here one CTA per (image, threshold) walks the image's detections in the
// decoder's order and marks the true positives; AP / mAP stay host NumPy (b200det.evaluation).
//
// Arithmetic (float32, the reference's op order, no FMA contraction: compiled with -fmad=false):
//   w = max(0, min(a.x2, b.x2) - max(a.x1, b.x1)), h likewise, overlap = w * h
//   area_a = (a.x2 - a.x1) * (a.y2 - a.y1), area_b likewise          (no clamps)
//   iou = overlap / ((area_a + area_b) - overlap)                    (0/0 -> NaN like NumPy)
// np.argmax over the ground truth of the detection's class: first maximum, NaN counts as the
// maximum (first NaN); true positive iff iou >= threshold (float32 compare, NaN fails) and that
// ground-truth box has not been taken by an earlier detection.
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"

namespace b200det {

constexpr int kEvalThreads = 128;

__device__ __forceinline__ float pair_iou(const float4 a, const float4 b) {
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    // np.maximum(0.0, NaN) is NaN; fmaxf would drop it
    const float dw = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float dh = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    const float ww = dw != dw ? dw : w, hh = dh != dh ? dh : h;
    const float overlap = __fmul_rn(ww, hh);
    const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    return __fdiv_rn(overlap, __fsub_rn(__fadd_rn(area_a, area_b), overlap));
}

// is (x, i) a better arg-max candidate than (y, j)?  NaN is the maximum; ties -> lower index
__device__ __forceinline__ bool better(float x, int i, float y, int j) {
    if (j < 0) return i >= 0;
    if (i < 0) return false;
    const bool xn = x != x, yn = y != y;
    if (xn || yn) return xn && (!yn || i < j);
    return x > y || (x == y && i < j);
}

__global__ void __launch_bounds__(256)
    pair_ious_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b, int n, int m,
                     float *__restrict__ out) {
    const long long total = (long long)n * m;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = pair_iou(__ldg(a + i / m), __ldg(b + i % m));
}

__global__ void __launch_bounds__(kEvalThreads)
    voc_match_kernel(const float4 *__restrict__ pred_boxes, const float *__restrict__ pred_classes,
                     const float4 *__restrict__ gt_boxes, const float *__restrict__ gt_classes,
                     const float *__restrict__ thresholds, int batch, int max_det, int max_gt,
                     unsigned char *__restrict__ tp) {
    extern __shared__ unsigned char taken[];   // [max_gt]
    __shared__ float red_v[kEvalThreads / 32];
    __shared__ int red_i[kEvalThreads / 32];
    const int b = blockIdx.x, t = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float thr = __ldg(thresholds + t);
    for (int g = tid; g < max_gt; g += kEvalThreads) taken[g] = 0;
    __syncthreads();
    const float4 *gb = gt_boxes + (size_t)b * max_gt;
    const float *gc = gt_classes + (size_t)b * max_gt;
    unsigned char *out = tp + ((size_t)t * batch + b) * max_det;
    for (int d = 0; d < max_det; ++d) {
        const float cls = __ldg(pred_classes + (size_t)b * max_det + d);
        if (!(cls > -1.f)) {   // padding (tools/scripts.py:570-575 drops classes <= -1)
            if (tid == 0) out[d] = 0;
            continue;
        }
        const float4 pb = __ldg(pred_boxes + (size_t)b * max_det + d);
        float best = 0.f;
        int best_g = -1;
        for (int g = tid; g < max_gt; g += kEvalThreads) {
            if (__ldg(gc + g) == cls) {   // ground truth of the detection's class, in order
                const float v = pair_iou(__ldg(gb + g), pb);
                if (better(v, g, best, best_g)) {
                    best = v;
                    best_g = g;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int og = __shfl_xor_sync(0xffffffffu, best_g, o);
            if (better(ov, og, best, best_g)) {
                best = ov;
                best_g = og;
            }
        }
        if (lane == 0) {
            red_v[warp] = best;
            red_i[warp] = best_g;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kEvalThreads / 32; ++w)
                if (better(red_v[w], red_i[w], best, best_g)) {
                    best = red_v[w];
                    best_g = red_i[w];
                }
            unsigned char hit = 0;
            if (best_g >= 0 && best >= thr && !taken[best_g]) {
                hit = 1;
                taken[best_g] = 1;
            }
            out[d] = hit;
        }
        __syncthreads();
    }
}

}  // namespace b200det

using namespace b200det;

extern "C" int b200det_pair_ious(const float *a, int n, const float *b, int m, float *out,
                                 void *stream) {
    if (n < 0 || m < 0) return B200DET_EINVAL;
    if (n == 0 || m == 0) return 0;
    if (!a || !b || !out) return B200DET_EINVAL;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) return B200DET_EALIGN;
    const long long total = (long long)n * m;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    ProfScope prof(kKernOther, stream);
    pair_ious_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(a), reinterpret_cast<const float4 *>(b), n, m, out);
    count_launch();
    return (int)cudaGetLastError();
}

extern "C" int b200det_voc_match(const float *pred_boxes, const float *pred_classes, int max_det,
                                 const float *gt_boxes, const float *gt_classes, int max_gt,
                                 const float *thresholds, int n_thresholds, int batch,
                                 unsigned char *tp, void *stream) {
    if (!pred_boxes || !pred_classes || !gt_boxes || !gt_classes || !thresholds || !tp)
        return B200DET_EINVAL;
    if (batch < 1 || max_det < 1 || max_gt < 1 || n_thresholds < 1) return B200DET_EINVAL;
    if (max_gt > 32768 || n_thresholds > 65535) return B200DET_ERANGE;
    if ((reinterpret_cast<uintptr_t>(pred_boxes) | reinterpret_cast<uintptr_t>(gt_boxes)) & 15)
        return B200DET_EALIGN;
    const dim3 grid((unsigned)batch, (unsigned)n_thresholds);
    ProfScope prof(kKernOther, stream);
    voc_match_kernel<<<grid, kEvalThreads, (size_t)max_gt, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(pred_boxes), pred_classes,
        reinterpret_cast<const float4 *>(gt_boxes), gt_classes, thresholds, batch, max_det, max_gt,
        tp);
    count_launch();
    return (int)cudaGetLastError();
}
