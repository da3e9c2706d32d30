// exchange.cu -- the loss normaliser's cross-GPU exchange fused into the reduction kernel.
//
// The path shards by image; the only coupling between ranks is the whole-batch positive
// count and the three loss sums (losses.py:231-259: 4 doubles).  With NCCL that is: reduce kernel ->
// host call -> all-reduce kernel(s) -> finish kernel.  Here ONE kernel per rank reduces its block
// partials, stores its 4 doubles straight into every peer's exchange buffer over NVLink (peer
// memory mapped with CUDA IPC), waits for the peers' stores to land in its own buffer, adds the
// W contributions in rank order (so every rank gets bit-identical totals) and applies the weights /
// normalisation: no host call, no extra launch, 40 bytes per peer on the wire.
//
// Exchange buffer of a rank (4 KB, zero-initialised): 2 sets (epoch parity) x 16 slots x 64 bytes;
// slot r of set s = {sums[4], epoch} written by rank r.  A rank can not run two epochs ahead of a
// peer (finishing epoch e+1 needs every peer's e+1 store, which a peer issues only after it has read
// epoch e), so two sets are enough.  Data and flag are written by the same thread, the flag with
// st.release.sys; the reader spins with ld.acquire.sys on the flag, bounded by a wall-clock budget
// (%globaltimer; default 120 s): on expiry the sums become NaN and the status word is set -- the
// kernel never hangs.
//
// The EPOCH lives in device memory (byte 2048 of the rank's own buffer) and is advanced by the
// kernel itself: every rank runs the same sequence of exchanges, so the counters move in lock-step,
// and a captured CUDA graph that is replayed advances them like eager calls do (a host-side counter
// passed by value would be frozen into the graph and a replay would match the previous replay's
// flags).  px->epoch != 0 still selects a caller-numbered exchange (tests).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "focal_terms.cuh"

namespace b200det {

constexpr int kMaxPeers = B200DET_MAX_PEERS;
constexpr int kSlotDoubles = 8;   // 64-byte slots
constexpr size_t kPeerBufferBytes = 4096;

constexpr size_t kEpochOffset = 2048;      // device-side epoch counter of the owning rank

struct ExchangeArgs {
    double *peer[kMaxPeers];
    int rank, world;
    unsigned long long epoch;        // 0: take (and advance) the counter in the rank's own buffer
    unsigned long long timeout_ns;
};
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(double *p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Same reduction as loss_reduce_kernel (focal.cu): fixed order, fp64; result valid in thread 0.
__device__ __forceinline__ void reduce_partials(const int *__restrict__ npos, long long n_assign,
                                                const SparsePartial *__restrict__ sp,
                                                long long n_sparse, const long long *__restrict__ fp,
                                                long long n_focal, double (&out)[4]) {
    __shared__ double red[4][32];
    double s_pos = 0.0, s_cls = 0.0, s_box = 0.0, s_ctr = 0.0;
    for (long long i = threadIdx.x; i < n_assign; i += blockDim.x) s_pos += (double)npos[i];
    for (long long i = threadIdx.x; i < n_sparse; i += blockDim.x) {
        const SparsePartial p = sp[i];
        s_box += p.box;
        s_ctr += p.ctr;
        s_cls += p.focal;
    }
    for (long long i = threadIdx.x; i < n_focal; i += blockDim.x) s_cls += (double)fp[i] / kFxSweep;
    if (threadIdx.x == 0 && fp[n_focal] != 0) s_cls = __longlong_as_double(0x7ff8000000000000ll);
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
        out[0] = a, out[1] = b, out[2] = c, out[3] = d;
    }
}

// After `local` (shared) holds this rank's 4 doubles: store them into every peer's buffer, wait for
// the peers', add in rank order, write totals / losses / status.  All threads of the CTA call it.
__device__ __forceinline__ void exchange_tail(const double *local, const ExchangeArgs &x, float w_cls,
                                              float w_box, float w_ctr, double *__restrict__ sums,
                                              float *__restrict__ losses, int *__restrict__ status) {
    __shared__ double gathered[kMaxPeers][4];
    __shared__ int failed;
    __shared__ unsigned long long s_epoch;
    if (threadIdx.x == 0) {
        failed = 0;
        unsigned long long e = x.epoch;
        if (e == 0) {   // this rank's exchange number, kept on the device (graph replays advance it)
            unsigned long long *ctr = reinterpret_cast<unsigned long long *>(
                reinterpret_cast<char *>(x.peer[x.rank]) + kEpochOffset);
            e = *ctr + 1ull;
            *ctr = e;
        }
        s_epoch = e;
    }
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    const int set = (int)(epoch & 1ull);
    const int t = threadIdx.x;
    if (t < x.world) {
        // my contribution into slot [rank] of peer t's buffer (t == rank: my own buffer)
        double *dst = x.peer[t] + ((size_t)set * kMaxPeers + x.rank) * kSlotDoubles;
#pragma unroll
        for (int k = 0; k < 4; ++k) st_relaxed_sys(dst + k, local[k]);
        st_release_sys(reinterpret_cast<unsigned long long *>(dst + 4), epoch);
        // peer t's contribution from slot [t] of my own buffer
        const double *src = x.peer[x.rank] + ((size_t)set * kMaxPeers + t) * kSlotDoubles;
        const unsigned long long t0 = globaltimer_ns();
        bool ok = true;
        while (ld_acquire_sys(reinterpret_cast<const unsigned long long *>(src + 4)) != epoch) {
            if (globaltimer_ns() - t0 > x.timeout_ns) {
                ok = false;
                break;
            }
            __nanosleep(64);
        }
        if (ok) {
#pragma unroll
            for (int k = 0; k < 4; ++k) gathered[t][k] = ld_relaxed_sys(src + k);
        } else {
            atomicExch(&failed, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r = 0; r < x.world; ++r)   // rank order: identical totals on every rank
            for (int k = 0; k < 4; ++k) tot[k] += gathered[r][k];
        if (failed) {
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            tot[0] = tot[1] = tot[2] = tot[3] = nan;
        }
        if (status) *status = failed;   // written on every call: the caller need not clear it
        for (int k = 0; k < 4; ++k) sums[k] = tot[k];
        if (losses) {
            // as loss_finish_kernel: float32 sum / count, then * weight (losses.py:259, :293, :318,
            // :210-211); 0 without positives (:234-235); NaN after a failed exchange
            const float w[3] = {w_cls, w_box, w_ctr};
            for (int i = 0; i < 3; ++i) {
                float v = 0.f;
                if (tot[0] > 0.0) v = w[i] * ((float)tot[1 + i] / (float)tot[0]);
                if (failed) v = __int_as_float(0x7fc00000);
                losses[i] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(1024)
    loss_reduce_exchange_kernel(const int *__restrict__ npos, long long n_assign,
                                const SparsePartial *__restrict__ sp, long long n_sparse,
                                const long long *__restrict__ fp, long long n_focal, ExchangeArgs x,
                                float w_cls, float w_box, float w_ctr, double *__restrict__ sums,
                                float *__restrict__ losses, int *__restrict__ status) {
    __shared__ double local[4];
    // the wait for the peers below may take a while: a dependent launched programmatically (the
    // decoder's select kernel on handed-over keys, which reads nothing this kernel writes) runs beside it
    pdl_launch_dependents();
    double mine[4] = {0.0, 0.0, 0.0, 0.0};
    reduce_partials(npos, n_assign, sp, n_sparse, fp, n_focal, mine);
    if (threadIdx.x == 0) local[0] = mine[0], local[1] = mine[1], local[2] = mine[2], local[3] = mine[3];
    __syncthreads();
    exchange_tail(local, x, w_cls, w_box, w_ctr, sums, losses, status);
}

// The exchange alone: sums[0..3] (this rank's) -> totals over the ranks, in place (+ losses).
__global__ void __launch_bounds__(32)
    sums_exchange_kernel(ExchangeArgs x, float w_cls, float w_box, float w_ctr,
                         double *__restrict__ sums, float *__restrict__ losses,
                         int *__restrict__ status) {
    __shared__ double local[4];
    if (threadIdx.x < 4) local[threadIdx.x] = sums[threadIdx.x];
    __syncthreads();
    exchange_tail(local, x, w_cls, w_box, w_ctr, sums, losses, status);
}

}  // namespace b200det

using namespace b200det;

extern "C" int b200det_peer_buffer_create(void **buffer, unsigned char *handle64) {
    if (!buffer || !handle64) return B200DET_EINVAL;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, kPeerBufferBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, kPeerBufferBytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return (int)e;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    *buffer = p;
    return 0;
}

extern "C" int b200det_peer_buffer_open(const unsigned char *handle64, void **mapped) {
    if (!handle64 || !mapped) return B200DET_EINVAL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    // enables peer access between the current device and the exporting device if needed
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return (int)e;
    *mapped = p;
    return 0;
}

extern "C" int b200det_peer_buffer_close(void *mapped) {
    return mapped ? (int)cudaIpcCloseMemHandle(mapped) : B200DET_EINVAL;
}

extern "C" int b200det_peer_buffer_destroy(void *buffer) {
    return buffer ? (int)cudaFree(buffer) : B200DET_EINVAL;
}

static int fill_exchange(const b200det_peer_exchange *px, ExchangeArgs *out) {
    if (!px) return B200DET_EINVAL;
    if (px->world < 1 || px->world > kMaxPeers || px->rank < 0 || px->rank >= px->world)
        return B200DET_ERANGE;
    ExchangeArgs &x = *out;
    for (int r = 0; r < kMaxPeers; ++r) x.peer[r] = nullptr;
    for (int r = 0; r < px->world; ++r) {
        if (!px->peer[r]) return B200DET_EINVAL;
        if (reinterpret_cast<uintptr_t>(px->peer[r]) & 63) return B200DET_EALIGN;
        x.peer[r] = static_cast<double *>(px->peer[r]);
    }
    x.rank = px->rank;
    x.world = px->world;
    x.epoch = px->epoch;
    // the field keeps its r01 name; since r02 it is a wall-clock budget in nanoseconds
    x.timeout_ns = px->timeout_cycles ? px->timeout_cycles : 120000000000ull;      // 120 s
    return 0;
}

extern "C" int b200det_sums_exchange(const b200det_peer_exchange *px, float w_cls, float w_box,
                                     float w_ctr, double *sums, float *losses, int32_t *status,
                                     void *stream) {
    if (!sums) return B200DET_EINVAL;
    ExchangeArgs x;
    int rc = fill_exchange(px, &x);
    if (rc) return rc;
    ProfScope prof(kKernReduce, stream);
    sums_exchange_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(x, w_cls, w_box, w_ctr, sums, losses,
                                                             status);
    count_launch();
    return (int)cudaGetLastError();
}

extern "C" int b200det_loss_reduce_exchange(const b200det_geometry *geo, const void *workspace,
                                            size_t workspace_bytes,
                                            const b200det_peer_exchange *px, float w_cls,
                                            float w_box, float w_ctr, double *sums, float *losses,
                                            int32_t *status, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!workspace || !px || !sums) return B200DET_EINVAL;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    ExchangeArgs x;
    if ((rc = fill_exchange(px, &x))) return rc;
    const char *base = static_cast<const char *>(workspace);
    ProfScope prof(kKernReduce, stream);
    loss_reduce_exchange_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const int *>(base + ws.off_assign), (long long)ws.assign_blocks,
        reinterpret_cast<const SparsePartial *>(base + ws.off_sparse), (long long)ws.sparse_blocks,
        reinterpret_cast<const long long *>(base + ws.off_focal), (long long)kSweepSlots, x, w_cls,
        w_box, w_ctr, sums, losses, status);
    count_launch();
    return (int)cudaGetLastError();
}

// The whole no-grad forward.  px == NULL: reduce (+ finish when `losses`); px != NULL: reduce +
// cross-rank exchange + finish in one kernel.  side / ev_fork / ev_join != NULL: the assignment and
// the sparse losses -- which do not depend on the focal sweep -- run on `side` beside it
// (fork after the memset, join before the reduction).  The caller owns the stream and the events.
namespace b200det {
int loss_forward_impl(const b200det_geometry *geo, const b200det_loss_params *p,
                      const float *annotations, int max_gt, const void *const *cls,
                      const void *const *reg, const void *const *ctr, int32_t *labels,
                      void *workspace, size_t workspace_bytes, const b200det_peer_exchange *px,
                      double *sums, float *losses, int32_t *status, void *side, void *ev_fork,
                      void *ev_join, void *stream, int phase, float keys_min_score, uint32_t *keys,
                      int32_t *classes) {
    // keys != NULL (b200det_loss_forward_keys): the sweep is the fused one, which also writes the
    // decoder's keys / classes.
    // phase: 0 = everything; 1 = only memset + fork + focal sweep (needs geo, p, cls, workspace);
    //        2 = the rest of a call whose phase 1 has been enqueued (same arguments)
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!p || !cls || !workspace) return B200DET_EINVAL;
    if (phase != 1 && (!annotations || !labels || !sums)) return B200DET_EINVAL;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    const bool fork = side && ev_fork && ev_join;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (phase != 2) {
        // sweep accumulators and queue counters are adjacent in the workspace: one memset
        char *base = static_cast<char *>(workspace);
        e = cudaMemsetAsync(base + ws.off_focal, 0, ws.off_counters + 2 * sizeof(int) - ws.off_focal, st);
        if (e != cudaSuccess) return (int)e;
        if (fork) {
            if ((e = cudaEventRecord((cudaEvent_t)ev_fork, st)) != cudaSuccess) return (int)e;
            if ((e = cudaStreamWaitEvent((cudaStream_t)side, (cudaEvent_t)ev_fork, 0)) != cudaSuccess)
                return (int)e;
        }
        g_skip_memset = true;
        // the long HBM-bound sweep first: the host prepares the remaining launches behind it
        if (keys) {
            if (!classes || (p->is_fcos && !ctr)) rc = B200DET_EINVAL;
            else
                rc = score_argmax_impl(geo, cls, p->is_fcos ? ctr : nullptr, keys_min_score, keys, classes,
                                       p->alpha, p->gamma,
                                       reinterpret_cast<long long *>(base + ws.off_focal), stream);
        } else {
            rc = b200det_focal_loss(geo, cls, nullptr, p->alpha, p->gamma, nullptr, nullptr, 0.f, workspace,
                                    workspace_bytes, stream);
        }
        g_skip_memset = false;
        if (phase == 1) return rc;
    }
    void *st_sparse = stream;
    if (fork) {
        st_sparse = side;
        static const int env_chunk = getenv("B200DET_ASSIGN_CHUNK") ? atoi(getenv("B200DET_ASSIGN_CHUNK")) : -1;
        g_assign_chunk = env_chunk >= 0 ? env_chunk : 0;
    }
    g_skip_memset = true;
    if (!rc) {
        rc = p->is_fcos ? b200det_fcos_assign(geo, annotations, max_gt, p->use_center_sample,
                                              labels, nullptr, nullptr, workspace,
                                              workspace_bytes, st_sparse)
                        : b200det_retina_assign(geo, annotations, max_gt, p->iou_neg, p->iou_pos,
                                                labels, nullptr, workspace, workspace_bytes,
                                                st_sparse);
    }
    if (!rc)
        rc = b200det_sparse_losses(geo, p->is_fcos, annotations, max_gt, labels, reg, p->reg_dtype,
                                   ctr, p->box_loss, p->beta, cls, p->alpha, p->gamma, nullptr,
                                   nullptr, workspace, workspace_bytes, st_sparse);
    g_skip_memset = false;
    g_assign_chunk = 0;
    if (fork) {
        // always re-join, also after a failed launch: the side stream must not stay forked (capture)
        e = cudaEventRecord((cudaEvent_t)ev_join, (cudaStream_t)side);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, (cudaEvent_t)ev_join, 0);
        if (!rc && e != cudaSuccess) rc = (int)e;
    }
    if (rc) return rc;
    if (px)
        return b200det_loss_reduce_exchange(geo, workspace, workspace_bytes, px, p->w_cls, p->w_box,
                                            p->w_ctr, sums, losses, status, stream);
    // reduction and normalisation in one launch unless the caller all-reduces in between
    return losses ? b200det_loss_reduce_finish(geo, workspace, workspace_bytes, p->w_cls, p->w_box,
                                               p->w_ctr, sums, losses, stream)
                  : b200det_loss_reduce(geo, 3, workspace, workspace_bytes, sums, stream);
}
}  // namespace b200det

// b200det_loss_forward with the reduction, the cross-rank exchange and the normalisation in one
// kernel (see above) instead of reduce -> [caller all-reduces] -> finish.
extern "C" int b200det_loss_forward_exchange(const b200det_geometry *geo,
                                             const b200det_loss_params *p,
                                             const float *annotations, int max_gt,
                                             const void *const *cls, const void *const *reg,
                                             const void *const *ctr, int32_t *labels,
                                             void *workspace, size_t workspace_bytes,
                                             const b200det_peer_exchange *px, double *sums,
                                             float *losses, int32_t *status, void *stream) {
    if (!px) return B200DET_EINVAL;
    return loss_forward_impl(geo, p, annotations, max_gt, cls, reg, ctr, labels, workspace,
                             workspace_bytes, px, sums, losses, status, nullptr, nullptr, nullptr,
                             stream, 0, 0.f, nullptr, nullptr);
}

extern "C" int b200det_loss_forward_overlap(const b200det_geometry *geo,
                                            const b200det_loss_params *p, const float *annotations,
                                            int max_gt, const void *const *cls,
                                            const void *const *reg, const void *const *ctr,
                                            int32_t *labels, void *workspace,
                                            size_t workspace_bytes,
                                            const b200det_peer_exchange *px, double *sums,
                                            float *losses, int32_t *status, void *side_stream,
                                            void *ev_fork, void *ev_join, void *stream, int phase) {
    if (phase < 0 || phase > 2) return B200DET_EINVAL;
    return loss_forward_impl(geo, p, annotations, max_gt, cls, reg, ctr, labels, workspace,
                             workspace_bytes, px, sums, losses, status, side_stream, ev_fork,
                             ev_join, stream, phase, 0.f, nullptr, nullptr);
}

// b200det_loss_forward_overlap whose sweep ALSO produces what the decoder's sweep would: the fused
// sweep of b200det_eval_step (focal sum + first-maximum class / score key per row) instead of the
// focal-only one.  A decoder call on the same head outputs then only needs b200det_select_decode_nms
// on these keys -- cls is read once per evaluation step inside the reference's two-call structure
// (tools/scripts.py:733-740).  FCOS: ctr is needed in phase 1 as well.
extern "C" int b200det_loss_forward_keys(const b200det_geometry *geo, const b200det_loss_params *p,
                                         const float *annotations, int max_gt,
                                         const void *const *cls, const void *const *reg,
                                         const void *const *ctr, int32_t *labels, void *workspace,
                                         size_t workspace_bytes, const b200det_peer_exchange *px,
                                         double *sums, float *losses, int32_t *status,
                                         void *side_stream, void *ev_fork, void *ev_join,
                                         void *stream, int phase, float min_score, uint32_t *keys,
                                         int32_t *classes) {
    if (phase < 0 || phase > 2 || !keys || !classes) return B200DET_EINVAL;
    return loss_forward_impl(geo, p, annotations, max_gt, cls, reg, ctr, labels, workspace,
                             workspace_bytes, px, sums, losses, status, side_stream, ev_fork,
                             ev_join, stream, phase, min_score, keys, classes);
}

// Caller-owned helper objects of b200det_loss_forward_overlap (the library keeps none itself).
extern "C" int b200det_stream_create(void **stream, int high_priority) {
    if (!stream) return B200DET_EINVAL;
    int lo = 0, hi = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (e != cudaSuccess) return (int)e;
    cudaStream_t s = nullptr;
    e = cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo);
    if (e != cudaSuccess) return (int)e;
    *stream = s;
    return 0;
}
extern "C" int b200det_stream_destroy(void *stream) {
    return stream ? (int)cudaStreamDestroy((cudaStream_t)stream) : B200DET_EINVAL;
}
extern "C" int b200det_event_create(void **event) {
    if (!event) return B200DET_EINVAL;
    cudaEvent_t ev = nullptr;
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return (int)e;
    *event = ev;
    return 0;
}
extern "C" int b200det_event_destroy(void *event) {
    return event ? (int)cudaEventDestroy((cudaEvent_t)event) : B200DET_EINVAL;
}
