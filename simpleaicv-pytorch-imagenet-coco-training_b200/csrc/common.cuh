// common.cuh -- shared device/host helpers for libb200det (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/b200det.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200det is written for sm_100a (B200) only"
#endif

// Debug builds (B200DET_NVCC_EXTRA=-DB200DET_DEBUG python -m b200det._build --force) check the
// capacities of the shared-memory lists / queues on the device (compute-sanitizer is not available on
// the GPU pool); release builds compile the checks away.
#ifdef B200DET_DEBUG
#include <assert.h>
#define B200DET_ASSERT(cond) assert(cond)
#else
#define B200DET_ASSERT(cond) ((void)0)
#endif

namespace b200det {

constexpr int kMaxLevels = B200DET_MAX_LEVELS;
constexpr int kMaxPerLoc = B200DET_MAX_PER_LOC;

// Device-side geometry: the POD the kernels take by value (no H2D copies, no globals).
struct Geo {
    int n_levels, batch, per_loc, num_classes;
    int H[kMaxLevels], W[kMaxLevels];
    int rows[kMaxLevels];     // rows of one image at level l  (H*W*per_loc)
    int off[kMaxLevels + 1];  // rows of one image before level l; off[n_levels] = N
    float stride[kMaxLevels];
};

struct BaseAnchors {
    float v[kMaxLevels][kMaxPerLoc][4];
};

struct FcosTab {
    float mi_lo[kMaxLevels], mi_hi[kMaxLevels], radius[kMaxLevels];
};

struct PtrTab {
    const void *p[kMaxLevels];
};
struct MutPtrTab {
    void *p[kMaxLevels];
};

// host: validate + derive.  Returns 0 or a B200DET_E* code.
int make_geo(const b200det_geometry *g, Geo *out);
void count_launch();
unsigned long long launches();

// Optional per-kernel timing with CUDA events on the launching stream (b200det_profile_*):
// bench.py reads the dominant kernel's average duration from the same launches it times.
enum KernelId {
    kKernFocal = 0,
    kKernAssign,
    kKernSparse,
    kKernReduce,
    kKernFinish,
    kKernArgmax,
    kKernSelect,
    kKernOther,
    kKernHeadTail,
    kKernLogits,
    kKernCount
};
struct ProfScope {   // records start on construction, stop on destruction (no-op when disabled)
    int slot;
    cudaStream_t st;
    ProfScope(int kernel_id, void *stream);
    ~ProfScope();
};

// Loss workspace (caller-owned scratch): per-CTA partials, reduced in fixed order by
// loss_reduce_kernel, plus the matched annotation row of every row (assign -> sparse kernel).
struct SparsePartial {
    double box, ctr, focal;
};
struct LossWs {
    size_t assign_blocks_per_image, assign_blocks, sparse_blocks, focal_chunks;
    size_t off_assign /* int npos [assign_blocks] */, off_sparse /* SparsePartial [sparse_blocks] */,
        off_focal /* int64 fixed-point [1024] */, off_counters /* int [2] */,
        off_pos_queue /* int2 [B*N] */, off_ign_queue /* int [B*N] */, total;
};
LossWs loss_ws_layout(const Geo &g);
int score_argmax_impl(const b200det_geometry *geo, const void *const *cls, const void *const *ctr,
                      float min_score, uint32_t *keys, int32_t *classes, float alpha, float gamma,
                      long long *focal_slots, void *stream);   // decode.cu
int select_decode_nms_impl(const b200det_geometry *geo, const uint32_t *keys, const int32_t *classes,
                           const void *const *reg, int reg_dtype, int is_fcos, float min_score,
                           int topn, int max_out, int nms_type, double nms_threshold,
                           const float *scales, const float *sizes, int to_xywh, float *out,
                           int32_t *order, int32_t *keep, int32_t *counts,
                           const uint16_t *half_exp_table, void *stream,
                           const void *const *verify_cls = nullptr, const void *const *verify_ctr = nullptr,
                           int32_t *stale = nullptr, bool inputs_complete = false);   // decode.cu
// set by the fused entry points after they have cleared all accumulators with ONE memset
extern thread_local bool g_skip_memset;
// set by the overlapped forward: images per assignment launch (0 = the whole batch in one launch)
extern thread_local int g_assign_chunk;
int assign_blocks_per_image(const Geo &g);  // assign.cu
int sparse_blocks(const Geo &g);            // assign.cu

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int level_of_row(const Geo &g, int row) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i)
        if (i < g.n_levels && row >= g.off[i]) l = i;
    return l;
}

// index of (image b, image-major row r at level l) in a level-major per-row array
__device__ __forceinline__ long long lm_index(const Geo &g, int b, int l, int local) {
    return (long long)g.batch * g.off[l] + (long long)b * g.rows[l] + local;
}

// (x+0.5)*stride: the reference computes it in float64 and rounds once to float32
// (models/anchor.py:65-71, :116-123); one DMUL per row keeps that exact for any stride.
__device__ __forceinline__ float shift_of(int i, float stride) {
    return (float)(((double)i + 0.5) * (double)stride);
}

__device__ __forceinline__ float4 anchor_of(const Geo &g, const BaseAnchors &ba, int l,
                                            int local) {
    const int a = local % g.per_loc;
    const int loc = local / g.per_loc;
    const int x = loc % g.W[l];
    const int y = loc / g.W[l];
    const float sx = shift_of(x, g.stride[l]);
    const float sy = shift_of(y, g.stride[l]);
    float4 r;  // base + shift in float32 (anchor.py:80)
    r.x = __fadd_rn(ba.v[l][a][0], sx);
    r.y = __fadd_rn(ba.v[l][a][1], sy);
    r.z = __fadd_rn(ba.v[l][a][2], sx);
    r.w = __fadd_rn(ba.v[l][a][3], sy);
    return r;
}

__device__ __forceinline__ float2 point_of(const Geo &g, int l, int local) {
    const int x = local % g.W[l];
    const int y = local / g.W[l];
    return make_float2(shift_of(x, g.stride[l]), shift_of(y, g.stride[l]));
}

// regression head element loader: 4 values of row `row` of a [rows,4] tensor, upcast to f32
__device__ __forceinline__ float4 load_reg4(const void *base, int dtype, long long row) {
    dtype &= 0xf;   // (B200DET_REG_EXP_ROUNDED rides in the high bits)
    if (dtype == B200DET_F32) {
        return __ldg(reinterpret_cast<const float4 *>(base) + row);
    } else if (dtype == B200DET_F16) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(base) + row);
        const __half2 lo = *reinterpret_cast<const __half2 *>(&raw.x);
        const __half2 hi = *reinterpret_cast<const __half2 *>(&raw.y);
        const float2 a = __half22float2(lo), b = __half22float2(hi);
        return make_float4(a.x, a.y, b.x, b.y);
    } else {
        const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(base) + row);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162 *>(&raw.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162 *>(&raw.y);
        const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
        return make_float4(a.x, a.y, b.x, b.y);
    }
}

// Programmatic dependent launch (sm_90+): a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may start once every CTA of its predecessor has
// executed launch_dependents (or exited); it must execute wait before touching anything the
// predecessor wrote (wait = the predecessor grid has completed and its writes are visible).
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// exp() of a regression value as eager half arithmetic sees it: computed in float32, then rounded to
// the tensor's own precision (B200DET_REG_EXP_ROUNDED); mode = reg_dtype incl. the flag
__device__ __forceinline__ float round_like(float v, int mode) {
    if (!(mode & B200DET_REG_EXP_ROUNDED)) return v;
    if ((mode & 0xf) == B200DET_F16) return __half2float(__float2half_rn(v));
    if ((mode & 0xf) == B200DET_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// NumPy's float32 exp (AVX512F / AVX2+FMA3 kernel), operation for operation; see oracle/npexp.c.
// Needed because RetinaDecoder / FCOSDecoder truncate exp-derived coordinates to int32
// (decode.py:260-268, :356-361): a correctly rounded exp would flip ~5e-7 of the coordinates.
__device__ __forceinline__ float npexp(float x) {
    const float xmax = 88.72283935546875f;
    const float xmin = -103.97208404541015625f;
    if (x != x) return x;
    if (x >= xmax) return __int_as_float(0x7f800000);
    if (x <= xmin) return 0.0f;
    const float magic = 12582912.0f;  // 1.5 * 2^23
    float q = __fmul_rn(x, 1.442695040888963407359924681001892137f);
    q = __fadd_rn(q, magic);
    q = __fsub_rn(q, magic);
    float r = __fmaf_rn(q, -6.93145752e-1f, x);
    r = __fmaf_rn(q, -1.42860677e-6f, r);
    float num = __fmaf_rn(5.082762527590693718096e-04f, r, 6.757896990527504603057e-03f);
    num = __fmaf_rn(num, r, 5.114512081637298353406e-02f);
    num = __fmaf_rn(num, r, 2.473615434895520810817e-01f);
    num = __fmaf_rn(num, r, 7.257664613233124478488e-01f);
    num = __fmaf_rn(num, r, 9.999999999980870924916e-01f);
    float den = __fmaf_rn(2.159509375685829852307e-02f, r, -2.742335390411667452936e-01f);
    den = __fmaf_rn(den, r, 1.0f);
    return ldexpf(__fdiv_rn(num, den), (int)q);
}

// x86 cvttss2si semantics of ndarray.astype(np.int32): truncate toward zero; NaN and
// out-of-range values become INT_MIN ("integer indefinite"), NOT saturated as in CUDA.
__device__ __forceinline__ int x86_f2i(float v) {
    if (!(v > -2147483904.0f && v < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rz(v);
}

// ---------------------------------------------------------------------------------------
// Staging of a contiguous float range, global -> shared memory, via one TMA bulk copy
// (cp.async.bulk + mbarrier; needs 16-byte aligned ends) or plain loads.  Users: the annotation rows
// of the assignment kernels, the raw class tiles of score_argmax_raw_kernel.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void stage_rows_begin(float *dst, const float *src, int n_floats,
                                                 uint64_t *mbar, bool bulk) {
    if (bulk) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t bytes = (uint32_t)n_floats * 4u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                             smem_u32(mbar)),
                         "r"(bytes)
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                    "r"(smem_u32(dst)),
                "l"(src), "r"(bytes), "r"(smem_u32(mbar))
                : "memory");
        }
    } else {
        for (int i = threadIdx.x; i < n_floats; i += blockDim.x) dst[i] = __ldg(src + i);
    }
}

__device__ __forceinline__ void stage_rows_wait(uint64_t *mbar, bool bulk) {
    __syncthreads();  // mbarrier init visible to all waiters / plain stores visible
    if (bulk) {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(done)
                : "r"(smem_u32(mbar))
                : "memory");
        }
    }
}


}  // namespace b200det
