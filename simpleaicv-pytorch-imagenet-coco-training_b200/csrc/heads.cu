// heads.cu -- the tail of the detection heads: sigmoid + NCHW -> NHWC in one pass.
//
// Reference: RetinaClsHead.forward `x = x.float(); x = self.sigmoid(x)` (models/head.py:46-50),
// FCOSClsRegCntHead.forward (head.py:176-179), then `permute(0, 2, 3, 1).contiguous()` in
// RetinaNet.forward / FCOS.forward (models/retinanet.py:73-77, models/fcos.py:70-79).  torch runs
// that as two elementwise-sized kernels (sigmoid: read + write, permute copy: read + write =
// 16 B per element); here it is one transposing pass (read 4 or 2 B, write 4 B per element), and
// the backward is one pass too (grad_in = grad_out * (1 - y) * y, transposed back, cast to the
// convolution's dtype -- the derivative torch's sigmoid_backward and `.float()` produce).
//
// Layout: per image the source is a [CH, HW] matrix (HW contiguous), the destination [HW, CH].
// A CTA of 256 threads moves a 64 (hw) x 64 (ch) tile: every thread issues 16 independent scalar
// loads (coalesced along hw, 128 B per warp instruction; rows are only 4-byte aligned because HW
// is odd on most pyramid levels), applies the activation in registers, parks float4s in shared
// memory with a 17-float4 pitch (conflict-free for both the 128-bit stores and the transposed
// 128-bit loads) and writes 256-byte row segments with 128-bit stores.  HBM-bound: 8 B/element.
//
// Arithmetic: y = 1 / (1 + expf(-x)) in float32, the expression torch's CUDA sigmoid evaluates
// (accurate expf, IEEE division), so the result is bit-identical to running the reference's ops on
// the GPU; against torch-CPU (vectorised exp) it differs by <= 2 ulp.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/b200det.h"
#include "common.cuh"

namespace b200det {

constexpr int kTile = 64;           // tile edge in both hw and ch
constexpr int kTileThreads = 256;
constexpr int kPitch4 = kTile / 4 + 1;  // float4 units per smem row

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__device__ __forceinline__ T load_stream(const T *p) { return __ldcs(p); }
template <>
__device__ __forceinline__ __nv_bfloat16 load_stream<__nv_bfloat16>(const __nv_bfloat16 *p) {
    const unsigned short raw = __ldcs(reinterpret_cast<const unsigned short *>(p));
    return __ushort_as_bfloat16(raw);
}
template <typename T>
__device__ __forceinline__ void store_stream(T *p, T v) { __stcs(p, v); }
template <>
__device__ __forceinline__ void store_stream<__nv_bfloat16>(__nv_bfloat16 *p, __nv_bfloat16 v) {
    __stcs(reinterpret_cast<unsigned short *>(p), __bfloat16_as_ushort(v));
}

__device__ __forceinline__ float sigmoid_like_torch(float x) { return 1.f / (1.f + expf(-x)); }

// src [B, CH, HW] (T) -> dst [B, HW, CH] float32 probabilities.  VEC: CH % 4 == 0 (16-byte rows)
template <typename T, bool VEC>
__global__ void __launch_bounds__(kTileThreads)
    sigmoid_permute_kernel(const T *__restrict__ src, float *__restrict__ dst, int CH, int HW) {
    __shared__ float4 tile[kTile * kPitch4];
    const int t = threadIdx.x;
    const int hw_l = t & (kTile - 1), grp = t >> 6;  // 4 groups of 64 threads
    const int hw0 = blockIdx.x * kTile, ch0 = blockIdx.y * kTile;
    const size_t img = (size_t)blockIdx.z * CH * HW;
    const T *__restrict__ s = src + img;
    float *__restrict__ d = dst + img;

    const int hw = hw0 + hw_l;
    const bool hw_ok = hw < HW;
    T raw[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ch = ch0 + (grp + 4 * j) * 4 + e;
            raw[j][e] = from_f32<T>(0.f);
            if (hw_ok && ch < CH) raw[j][e] = load_stream(s + (size_t)ch * HW + hw);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 y;
        y.x = sigmoid_like_torch(to_f32(raw[j][0]));
        y.y = sigmoid_like_torch(to_f32(raw[j][1]));
        y.z = sigmoid_like_torch(to_f32(raw[j][2]));
        y.w = sigmoid_like_torch(to_f32(raw[j][3]));
        tile[hw_l * kPitch4 + grp + 4 * j] = y;
    }
    __syncthreads();
    const int c4 = t & 15, r0 = t >> 4;
    const int ch = ch0 + c4 * 4;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int r = r0 + 16 * it;
        const int hw_g = hw0 + r;
        if (hw_g >= HW) continue;
        const float4 y = tile[r * kPitch4 + c4];
        float *row = d + (size_t)hw_g * CH + ch;
        if (VEC) {
            if (ch < CH) __stcs(reinterpret_cast<float4 *>(row), y);
        } else {
            if (ch + 0 < CH) __stcs(row + 0, y.x);
            if (ch + 1 < CH) __stcs(row + 1, y.y);
            if (ch + 2 < CH) __stcs(row + 2, y.z);
            if (ch + 3 < CH) __stcs(row + 3, y.w);
        }
    }
}

// grad_in[b, ch, hw] = T(grad_out[b, hw, ch] * (1 - y[b, hw, ch]) * y[b, hw, ch])
template <typename T, bool VEC>
__global__ void __launch_bounds__(kTileThreads)
    sigmoid_permute_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ probs,
                               T *__restrict__ gin, int CH, int HW) {
    __shared__ float4 tile[kTile * kPitch4];
    const int t = threadIdx.x;
    const int hw0 = blockIdx.x * kTile, ch0 = blockIdx.y * kTile;
    const size_t img = (size_t)blockIdx.z * CH * HW;
    const float *__restrict__ g = gout + img;
    const float *__restrict__ p = probs + img;
    T *__restrict__ o = gin + img;

    const int c4 = t & 15, r0 = t >> 4;
    const int ch = ch0 + c4 * 4;
    float4 gv[4], pv[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int hw_g = hw0 + r0 + 16 * it;
        gv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        pv[it] = gv[it];
        if (hw_g < HW) {
            const size_t at = (size_t)hw_g * CH + ch;
            if (VEC) {
                if (ch < CH) {
                    gv[it] = __ldcs(reinterpret_cast<const float4 *>(g + at));
                    pv[it] = __ldcs(reinterpret_cast<const float4 *>(p + at));
                }
            } else {
                if (ch + 0 < CH) { gv[it].x = __ldcs(g + at + 0); pv[it].x = __ldcs(p + at + 0); }
                if (ch + 1 < CH) { gv[it].y = __ldcs(g + at + 1); pv[it].y = __ldcs(p + at + 1); }
                if (ch + 2 < CH) { gv[it].z = __ldcs(g + at + 2); pv[it].z = __ldcs(p + at + 2); }
                if (ch + 3 < CH) { gv[it].w = __ldcs(g + at + 3); pv[it].w = __ldcs(p + at + 3); }
            }
        }
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        float4 dx;  // torch sigmoid_backward: (grad * (1 - y)) * y
        dx.x = (gv[it].x * (1.f - pv[it].x)) * pv[it].x;
        dx.y = (gv[it].y * (1.f - pv[it].y)) * pv[it].y;
        dx.z = (gv[it].z * (1.f - pv[it].z)) * pv[it].z;
        dx.w = (gv[it].w * (1.f - pv[it].w)) * pv[it].w;
        tile[(r0 + 16 * it) * kPitch4 + c4] = dx;
    }
    __syncthreads();
    const int hw_l = t & (kTile - 1), grp = t >> 6;
    const int hw = hw0 + hw_l;
    if (hw >= HW) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 dx = tile[hw_l * kPitch4 + grp + 4 * j];
        const int c = ch0 + (grp + 4 * j) * 4;
        if (c + 0 < CH) store_stream(o + (size_t)(c + 0) * HW + hw, from_f32<T>(dx.x));
        if (c + 1 < CH) store_stream(o + (size_t)(c + 1) * HW + hw, from_f32<T>(dx.y));
        if (c + 2 < CH) store_stream(o + (size_t)(c + 2) * HW + hw, from_f32<T>(dx.z));
        if (c + 3 < CH) store_stream(o + (size_t)(c + 3) * HW + hw, from_f32<T>(dx.w));
    }
}

static int check_tail_args(const void *a, const void *b, int batch, int channels, long long hw,
                           int dtype, dim3 *grid) {
    if (!a || !b) return B200DET_EINVAL;
    if (batch < 0 || channels < 1 || hw < 0) return B200DET_EINVAL;
    if (dtype != B200DET_F32 && dtype != B200DET_F16 && dtype != B200DET_BF16) return B200DET_EINVAL;
    if (batch > 65535 || hw > 0x7fffffffLL || (long long)channels * hw > 0x7fffffffLL)
        return B200DET_ERANGE;
    const long long tiles_ch = (channels + kTile - 1) / kTile;
    if (tiles_ch > 65535) return B200DET_ERANGE;
    *grid = dim3((unsigned)((hw + kTile - 1) / kTile), (unsigned)tiles_ch, (unsigned)batch);
    return 0;
}

}  // namespace b200det

using namespace b200det;

extern "C" int b200det_head_sigmoid_permute(const void *src, int src_dtype, int batch, int channels,
                                            long long hw, float *dst, void *stream) {
    dim3 grid;
    if (batch == 0 || hw == 0) return 0;
    const int rc = check_tail_args(src, dst, batch, channels, hw, src_dtype, &grid);
    if (rc) return rc;
    const bool vec = (channels & 3) == 0;
    if (vec && (reinterpret_cast<uintptr_t>(dst) & 15)) return B200DET_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(kKernHeadTail, stream);
    const int CH = channels, HW = (int)hw;
#define B200DET_TAIL(T)                                                                              \
    do {                                                                                             \
        if (vec) sigmoid_permute_kernel<T, true><<<grid, kTileThreads, 0, st>>>(                     \
            static_cast<const T *>(src), dst, CH, HW);                                               \
        else sigmoid_permute_kernel<T, false><<<grid, kTileThreads, 0, st>>>(                        \
            static_cast<const T *>(src), dst, CH, HW);                                               \
    } while (0)
    if (src_dtype == B200DET_F32) B200DET_TAIL(float);
    else if (src_dtype == B200DET_F16) B200DET_TAIL(__half);
    else B200DET_TAIL(__nv_bfloat16);
#undef B200DET_TAIL
    count_launch();
    return (int)cudaGetLastError();
}

extern "C" int b200det_head_sigmoid_permute_backward(const float *grad_out, const float *probs,
                                                     int batch, int channels, long long hw,
                                                     void *grad_in, int grad_dtype, void *stream) {
    dim3 grid;
    if (batch == 0 || hw == 0) return 0;
    if (!grad_out) return B200DET_EINVAL;
    const int rc = check_tail_args(probs, grad_in, batch, channels, hw, grad_dtype, &grid);
    if (rc) return rc;
    const bool vec = (channels & 3) == 0;
    if (vec && ((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(probs)) & 15))
        return B200DET_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(kKernHeadTail, stream);
    const int CH = channels, HW = (int)hw;
#define B200DET_TAIL_BWD(T)                                                                          \
    do {                                                                                             \
        if (vec) sigmoid_permute_bwd_kernel<T, true><<<grid, kTileThreads, 0, st>>>(                 \
            grad_out, probs, static_cast<T *>(grad_in), CH, HW);                                     \
        else sigmoid_permute_bwd_kernel<T, false><<<grid, kTileThreads, 0, st>>>(                    \
            grad_out, probs, static_cast<T *>(grad_in), CH, HW);                                     \
    } while (0)
    if (grad_dtype == B200DET_F32) B200DET_TAIL_BWD(float);
    else if (grad_dtype == B200DET_F16) B200DET_TAIL_BWD(__half);
    else B200DET_TAIL_BWD(__nv_bfloat16);
#undef B200DET_TAIL_BWD
    count_launch();
    return (int)cudaGetLastError();
}
