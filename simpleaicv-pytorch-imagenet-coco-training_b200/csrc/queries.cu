// queries.cu -- front end of the query-based decoders (DETRDecoder / DINODETRDecoder,
// decode.py:367-594) and of a stand-alone DecodeMethod (decode.py:107-172).
//
// The reference takes the last decoder layer's [B, Q, C] class logits, applies softmax (DETR) or
// sigmoid (DINO-DETR) with torch on the model's device, copies everything to the host and does the
// arg-max, the score gather, the class / score filters, the cxcywh -> xyxy box transform and the
// per-image sort / top-n / NMS in NumPy.  Here one warp per (image, query) row produces the same
// order-preserving score key + class the dense decoders use, plus the xyxy box, and
// select_nms_kernel (decode.cu) does the rest on the device.
//
// Arithmetic (bit-exact targets):
//  * SIGMOID  1 / (1 + expf(-x)) with IEEE division: torch's CUDA sigmoid (tests/test_gpu_heads.py).
//  * SOFTMAX  torch's CUDA softmax for rows of <= 1024 elements (softmax_warp_forward in ATen's
//             PersistentSoftmax.cuh): with P = next power of two >= C and W = min(32, P), lane i
//             owns the elements i, i + W, i + 2W, ...; the row maximum is a per-lane scan followed
//             by an xor-butterfly; every element becomes expf(x - max); the per-lane sums (in
//             element order, starting from 0) are added by the same butterfly; p = e / sum.
//             Reproduced op for op, so scores match a CUDA run of the reference bit for bit.
//  * arg-max  np.argmax's first maximum of the PROBABILITIES (decode.py:397, :521).
//  * boxes    (cx - 0.5 w) * W etc. as separate float32 operations (decode.py:411-416, :474-482).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"

namespace b200det {

constexpr int kQThreads = 256;
constexpr int kQWarps = kQThreads / 32;

struct QueryArgs {
    const void *cls;
    const float *boxes;      // [rows, 4] cx,cy,w,h or null
    const float *sizes_hw;   // [B, 2]
    long long rows;          // B * Q
    int queries, channels, num_classes, dtype;
    int lanes;               // W = min(32, next_pow2(channels))
    float min_score;
    uint32_t *keys;
    int *classes;
    float *boxes_out;
};

__device__ __forceinline__ float q_load(const void *base, int dtype, long long i) {
    if (dtype == B200DET_F32) return __ldg(static_cast<const float *>(base) + i);
    if (dtype == B200DET_F16) return __half2float(__ldg(static_cast<const __half *>(base) + i));
    return __bfloat162float(__ushort_as_bfloat16(__ldg(static_cast<const unsigned short *>(base) + i)));
}

__device__ __forceinline__ uint32_t q_flip_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <int MODE>
__global__ void __launch_bounds__(kQThreads)
    query_scores_kernel(QueryArgs a) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kQWarps + (threadIdx.x >> 5);
    if (row >= a.rows) return;   // whole warps leave together
    const int C = a.channels, W = a.lanes;
    const long long base = row * C;
    const float ninf = -__int_as_float(0x7f800000);

    float row_max = 0.f, row_sum = 1.f;
    if (MODE == B200DET_SCORES_SOFTMAX) {
        // lanes >= W (rows shorter than 32) hold the identities and mirror their partner lanes
        float m = ninf;
        if (lane < W) {
            m = lane < C ? q_load(a.cls, a.dtype, base + lane) : ninf;
            for (int c = lane + W; c < C; c += W) {
                const float e = q_load(a.cls, a.dtype, base + c);
                m = (m > e) ? m : e;
            }
        }
        for (int o = W >> 1; o > 0; o >>= 1) {
            const float other = __shfl_xor_sync(0xffffffffu, m, o);
            m = (m < other) ? other : m;
        }
        float s = 0.f;
        if (lane < W)
            for (int c = lane; c < C; c += W) s = __fadd_rn(s, expf(__fsub_rn(q_load(a.cls, a.dtype, base + c), m)));
        for (int o = W >> 1; o > 0; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        row_max = m;
        row_sum = s;
    }

    // first maximum of the probabilities: strict '>' inside the lane (ascending classes), then
    // larger value / lower class between lanes
    float best = ninf;
    int best_c = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        const float x = q_load(a.cls, a.dtype, base + c);
        float p;
        if (MODE == B200DET_SCORES_SOFTMAX) p = __fdiv_rn(expf(__fsub_rn(x, row_max)), row_sum);
        else if (MODE == B200DET_SCORES_SIGMOID) p = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
        else p = x;
        if (best_c == 0x7fffffff || p > best) {
            best = p;
            best_c = c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
        if (oc != 0x7fffffff && (best_c == 0x7fffffff || ob > best || (ob == best && oc < best_c))) {
            best = ob;
            best_c = oc;
        }
    }
    if (lane == 0) {
        // decode.py:423-431 (class filter, DETR only) and :133-138 / :433-438 (strict threshold)
        const bool ok = best_c < a.num_classes && best > a.min_score;
        a.keys[row] = ok ? q_flip_key(best) : 0u;
        a.classes[row] = best_c;
    }
    if (a.boxes && lane == 1) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(a.boxes) + row);
        const int b = (int)(row / a.queries);
        const float h = __ldg(a.sizes_hw + 2 * b), w = __ldg(a.sizes_hw + 2 * b + 1);
        const float hw = __fmul_rn(0.5f, t.z), hh = __fmul_rn(0.5f, t.w);
        float4 o;
        o.x = __fmul_rn(__fsub_rn(t.x, hw), w);
        o.y = __fmul_rn(__fsub_rn(t.y, hh), h);
        o.z = __fmul_rn(__fadd_rn(t.x, hw), w);
        o.w = __fmul_rn(__fadd_rn(t.y, hh), h);
        reinterpret_cast<float4 *>(a.boxes_out)[row] = o;
    }
}

}  // namespace b200det

using namespace b200det;

extern "C" int b200det_query_scores(const void *cls, int cls_dtype, int mode,
                                    const float *boxes_cxcywh, const float *sizes_hw, int batch,
                                    int queries, int channels, int num_classes, float min_score,
                                    uint32_t *keys, int32_t *classes, float *boxes_xyxy,
                                    void *stream) {
    if (!cls || !keys || !classes) return B200DET_EINVAL;
    if (batch < 1 || queries < 1 || channels < 1 || num_classes < 1) return B200DET_EINVAL;
    if (mode < B200DET_SCORES_PROBS || mode > B200DET_SCORES_SOFTMAX) return B200DET_EINVAL;
    if (cls_dtype != B200DET_F32 && cls_dtype != B200DET_F16 && cls_dtype != B200DET_BF16)
        return B200DET_EINVAL;
    // the reference's softmax runs in the logits' own dtype; only float32 is reproduced, and only
    // the row lengths torch gives to its warp-softmax kernel
    if (mode == B200DET_SCORES_SOFTMAX && (cls_dtype != B200DET_F32 || channels > 1024))
        return B200DET_ERANGE;
    if (mode == B200DET_SCORES_PROBS && cls_dtype != B200DET_F32) return B200DET_EINVAL;
    if ((long long)batch * queries >= (1ll << 31)) return B200DET_ERANGE;
    if ((long long)batch * queries * channels >= (1ll << 40)) return B200DET_ERANGE;
    if (boxes_cxcywh) {
        if (!sizes_hw || !boxes_xyxy) return B200DET_EINVAL;
        if ((reinterpret_cast<uintptr_t>(boxes_cxcywh) | reinterpret_cast<uintptr_t>(boxes_xyxy)) & 15)
            return B200DET_EALIGN;
    }
    const uintptr_t amask = cls_dtype == B200DET_F32 ? 3 : 1;
    if (reinterpret_cast<uintptr_t>(cls) & amask) return B200DET_EALIGN;
    QueryArgs a;
    a.cls = cls;
    a.boxes = boxes_cxcywh;
    a.sizes_hw = sizes_hw;
    a.rows = (long long)batch * queries;
    a.queries = queries;
    a.channels = channels;
    a.num_classes = num_classes;
    a.dtype = cls_dtype;
    int p2 = 1;
    while (p2 < channels) p2 <<= 1;
    a.lanes = p2 < 32 ? p2 : 32;
    a.min_score = min_score;
    a.keys = keys;
    a.classes = classes;
    a.boxes_out = boxes_xyxy;
    const long long blocks = (a.rows + kQWarps - 1) / kQWarps;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(kKernOther, stream);
    if (mode == B200DET_SCORES_SOFTMAX)
        query_scores_kernel<B200DET_SCORES_SOFTMAX><<<(unsigned)blocks, kQThreads, 0, st>>>(a);
    else if (mode == B200DET_SCORES_SIGMOID)
        query_scores_kernel<B200DET_SCORES_SIGMOID><<<(unsigned)blocks, kQThreads, 0, st>>>(a);
    else
        query_scores_kernel<B200DET_SCORES_PROBS><<<(unsigned)blocks, kQThreads, 0, st>>>(a);
    count_launch();
    return (int)cudaGetLastError();
}
