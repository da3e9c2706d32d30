// assign.cu -- target assignment of RetinaLoss / FCOSLoss and the sparse losses that depend on it.
// Compiled with -fmad=false: every float op on the assignment path is one IEEE float32 operation
// in the reference's order, so labels and matched indices are bit-exact.
//
// Two kernels per loss call:
//
//  *_assign_kernel   pure ALU scan, no head tensors touched.  Per image it reads G <= 2048
//      annotation rows (20 B each, staged once per CTA into shared memory with one cp.async.bulk +
//      mbarrier), generates anchors / points in registers and writes 4 B (label) + 2 B (matched
//      annotation row) per row.  Instruction count per (anchor, GT) pair is what matters:
//        - a WARP owns an 8x4 patch of locations of one pyramid level, a THREAD one location; for
//          RetinaNet's 9 anchors per location the 9 anchors live in registers, so a GT box read
//          from shared memory is reused 9 times and level/x/y/shift arithmetic is paid once;
//        - GT boxes that cannot overlap the CTA's patch row are culled once per CTA (ordered ballot
//          compaction, so "first maximum / first minimum" survive), then per WARP with one ballot
//          per 32 candidates; a culled GT has IoU exactly 0 with every anchor of the warp and can
//          never beat best >= 0 under the reference's strict '>' scan;
//        - the IEEE divide is skipped when the intersection is empty.
//
//  sparse_loss_kernel   driven by the labels: for the ~1 % positive rows it gathers the matched
//      annotation, the regression row (and FCOS centre-ness) and evaluates the box / centre-ness
//      loss and their gradients; when asked, it also produces the focal-loss CORRECTIONS (target
//      class of positives, every class of ignored rows) that let the classification sweep in
//      focal.cu run label-free and concurrently.  Sums are accumulated in 64-bit fixed point, so
//      they do not depend on the order in which rows are visited (deterministic).
#include <stdlib.h>
#include <atomic>
#include "common.cuh"
#include "dual.cuh"
#include <algorithm>
#include "focal_terms.cuh"

namespace b200det {

constexpr int kAssignThreads = 128;
constexpr int kAssignWarps = kAssignThreads / 32;
constexpr int kTileW = 8, kTileH = 4;   // locations per warp: 8 wide x 4 high
#ifndef B200DET_SPARSE_THREADS
#define B200DET_SPARSE_THREADS 256
#endif
constexpr int kSparseThreads = B200DET_SPARSE_THREADS;

// Shared-memory carve-up (dynamic): raw rows, then the compacted candidate arrays.
struct GtSmem {
    float *raw;    // [G*5]
    float4 *box;   // [G]
    float *area;   // [G]
    int *label;    // [G]  class + 1
    short *fidx;   // [G]  index in the filtered (class >= 0) list
    short *ridx;   // [G]  annotation row (unfiltered)
};
__device__ __forceinline__ GtSmem carve(unsigned char *base, int G) {
    GtSmem s;
    const int raw_floats = (G * 5 + 3) & ~3;
    s.raw = reinterpret_cast<float *>(base);
    s.box = reinterpret_cast<float4 *>(s.raw + raw_floats);
    s.area = reinterpret_cast<float *>(s.box + G);
    s.label = reinterpret_cast<int *>(s.area + G);
    s.fidx = reinterpret_cast<short *>(s.label + G);
    s.ridx = s.fidx + G;
    return s;
}
__host__ __device__ static inline size_t gt_smem_bytes(int G) {
    const int raw_floats = (G * 5 + 3) & ~3;
    return (size_t)raw_floats * 4 + (size_t)G * (16 + 4 + 4 + 2 + 2);
}

enum CullMode { kCullNone = 0, kCullOverlap = 1, kCullContain = 2 };

// Ordered compaction of annotation rows: keeps rows with class >= 0 (filtered index = rank
// among them, losses.py:338-339 / :667-668) and, depending on `mode`, only those whose box can
// matter for the CTA's region [rx1,ry1,rx2,ry2]:
//   kCullOverlap : the box has a non-empty intersection with the region (Retina IoU > 0)
//   kCullContain : the box can strictly contain a point of the region (FCOS min(l,t,r,b) > 0)
// Returns {#kept, #valid}.  Order is preserved, which keeps the reference's "first maximum /
// first minimum" tie rules exact.
template <int THREADS = kAssignThreads>
__device__ __forceinline__ int2 compact_gt(const GtSmem &s, int G, int mode, float rx1, float ry1,
                                           float rx2, float ry2, bool fcos_area,
                                           int *warp_cnt /* [2 * THREADS / 32] smem */) {
    constexpr int kAssignThreads = THREADS, kAssignWarps = THREADS / 32;   // of THIS instantiation
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base_valid = 0, base_keep = 0;
    for (int j0 = 0; j0 < G; j0 += kAssignThreads) {
        const int j = j0 + threadIdx.x;
        float x1 = 0, y1 = 0, x2 = 0, y2 = 0, c = -1;
        if (j < G) {
            x1 = s.raw[j * 5 + 0];
            y1 = s.raw[j * 5 + 1];
            x2 = s.raw[j * 5 + 2];
            y2 = s.raw[j * 5 + 3];
            c = s.raw[j * 5 + 4];
        }
        const bool valid = (j < G) && (c >= 0.f);
        bool keep = valid;
        if (valid && mode == kCullOverlap)
            keep = (fminf(rx2, x2) > fmaxf(rx1, x1)) && (fminf(ry2, y2) > fmaxf(ry1, y1));
        if (valid && mode == kCullContain)
            keep = (x1 < rx2) && (x2 > rx1) && (y1 < ry2) && (y2 > ry1);
        const unsigned bv = __ballot_sync(0xffffffffu, valid);
        const unsigned bk = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) {
            warp_cnt[warp] = __popc(bv);
            warp_cnt[kAssignWarps + warp] = __popc(bk);
        }
        __syncthreads();
        int pv = base_valid, pk = base_keep, tv = 0, tk = 0;
#pragma unroll
        for (int w = 0; w < kAssignWarps; ++w) {
            const int cv = warp_cnt[w], ck = warp_cnt[kAssignWarps + w];
            if (w < warp) {
                pv += cv;
                pk += ck;
            }
            tv += cv;
            tk += ck;
        }
        const unsigned lower = (1u << lane) - 1u;
        pv += __popc(bv & lower);
        pk += __popc(bk & lower);
        if (keep) {
            s.box[pk] = make_float4(x1, y1, x2, y2);
            if (fcos_area) {
                // FCOS: plain (x2-x1)*(y2-y1), losses.py:785-788
                s.area[pk] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
            } else {
                // IoU: clamp(w,0)*clamp(h,0), losses.py:62-65
                s.area[pk] = __fmul_rn(fmaxf(__fsub_rn(x2, x1), 0.f), fmaxf(__fsub_rn(y2, y1), 0.f));
            }
            s.label[pk] = (int)(c + 1.f);
            s.fidx[pk] = (short)pv;
            s.ridx[pk] = (short)j;
        }
        base_valid += tv;
        base_keep += tk;
        __syncthreads();
    }
    return make_int2(base_keep, base_valid);
}

// CTA-wide min/max of a per-thread box -> region[4] in shared memory; also returns the WARP's box.
__device__ __forceinline__ float4 reduce_region(bool active, float x1, float y1, float x2, float y2,
                                                float *red /* [4*kAssignWarps] */,
                                                float *region /* [4] */) {
    const float big = 3.0e38f;
    float mnx = active ? x1 : big, mny = active ? y1 : big;
    float mxx = active ? x2 : -big, mxy = active ? y2 : -big;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[warp * 4 + 0] = mnx;
        red[warp * 4 + 1] = mny;
        red[warp * 4 + 2] = mxx;
        red[warp * 4 + 3] = mxy;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = red[0], b = red[1], c = red[2], d = red[3];
        for (int w = 1; w < kAssignWarps; ++w) {
            a = fminf(a, red[w * 4 + 0]);
            b = fminf(b, red[w * 4 + 1]);
            c = fmaxf(c, red[w * 4 + 2]);
            d = fmaxf(d, red[w * 4 + 3]);
        }
        region[0] = a;
        region[1] = b;
        region[2] = c;
        region[3] = d;
    }
    return make_float4(mnx, mny, mxx, mxy);
}

// Location owned by this thread: warp -> 8x4 patch of one level, lane -> location in the patch.
struct TileTab {
    int tile_off[kMaxLevels + 1];   // patches of one image before level l
    int tiles_x[kMaxLevels];
    // Retina: extent of a level's base anchors {min x1, min y1, max x2, max y2}.  Rounding is
    // monotonic, so min_a fl(base_a + shift) == fl(min_a base_a + shift): four adds give the exact
    // bounding box of a location's anchors.
    float ext[kMaxLevels][4];
};
struct Loc {
    bool active;
    int l, x, y, loc;   // level, position, y*W + x
};
__device__ __forceinline__ Loc my_location(const Geo &g, const TileTab &tt) {
    Loc r;
    r.active = false;
    r.l = r.x = r.y = r.loc = 0;
    const int tile = blockIdx.x * kAssignWarps + (threadIdx.x >> 5);
    if (tile >= tt.tile_off[g.n_levels]) return r;
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i)
        if (i < g.n_levels && tile >= tt.tile_off[i]) l = i;
    const int tl = tile - tt.tile_off[l];
    const int ty = tl / tt.tiles_x[l], tx = tl - ty * tt.tiles_x[l];
    const int lane = threadIdx.x & 31;
    r.l = l;
    r.x = tx * kTileW + (lane & (kTileW - 1));
    r.y = ty * kTileH + (lane / kTileW);
    r.active = r.x < g.W[l] && r.y < g.H[l];
    r.loc = r.y * g.W[l] + r.x;
    return r;
}

// Work queues handed from the assignment kernels to sparse_loss_kernel (global memory, in the
// caller's workspace).  Order inside the queues is arbitrary (atomics); every consumer
// accumulates in fixed point, so results do not depend on it.
struct Queues {
    int *counters;   // [0] = positives queued, [1] = ignored rows queued
    int2 *pos;       // (level-major row, annotation row)
    int *ign;        // level-major row
};

// CTA-local queues in shared memory -> one global atomic per queue and CTA, coalesced copy-out.
// Also stores the CTA's positive count (deterministic integer partial).
__device__ __forceinline__ void flush_queues(const Queues &q, const int2 *pos_q, int n_pos,
                                             const int *ign_q, int n_ign, int *npos_dst,
                                             int *bases /* [2] smem */) {
    if (threadIdx.x == 0) {
        *npos_dst = n_pos;
        bases[0] = n_pos ? atomicAdd(q.counters + 0, n_pos) : 0;
        bases[1] = n_ign ? atomicAdd(q.counters + 1, n_ign) : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_pos; i += kAssignThreads) q.pos[bases[0] + i] = pos_q[i];
    for (int i = threadIdx.x; i < n_ign; i += kAssignThreads) q.ign[bases[1] + i] = ign_q[i];
}

// ---------------------------------------------------------------------------------------
// Retina: anchor <-> GT IoU, max / arg-max, label  (losses.py:322-388)
//   PL > 0 : per_loc == PL anchors of the location held in registers (scan unrolled over them)
//   PL == 0: any per_loc, anchors processed one after the other
// ---------------------------------------------------------------------------------------
// EXACT = false (production: the arg-max index of non-positive rows is not an output): a pair
// whose IoU cannot reach the 0.4 "background" threshold can change neither a label nor a
// positive's matched box, so it is dropped early with conservative float32 bounds
// (IoU <= min(area)/max(area), and ov < floor*union, floor = 0.38 for RetinaLoss) and the IEEE
// divide only runs for the few pairs above it.  EXACT = true reproduces the reference's
// max/arg-max for every row.
struct IouThresholds {
    float neg;     // best IoU <  neg -> background   (RetinaLoss 0.4, RetinaFaceLoss 0.35)
    float pos;     // best IoU >= pos -> positive     (RetinaLoss 0.5, RetinaFaceLoss 0.35)
    float floor;   // 0.95 * min(neg, pos): pairs that cannot reach it are dropped (production scan)
};

template <int NA, bool EXACT>
__device__ __forceinline__ void scan_candidates(const GtSmem &s, int n_cand, const float4 wreg,
                                                const float4 (&A)[NA], float (&best)[NA],
                                                int (&best_slot)[NA], const float kIouFloor) {
    float area_a[NA], area_lo[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) {
        area_a[a] = __fmul_rn(fmaxf(__fsub_rn(A[a].z, A[a].x), 0.f),
                              fmaxf(__fsub_rn(A[a].w, A[a].y), 0.f));
        area_lo[a] = kIouFloor * area_a[a];
        best[a] = 0.f;
        best_slot[a] = -1;
    }
    // Bounds over the location's anchors (identical in every lane: a warp's patch lies on one level)
    float amin = area_a[0], amax = area_a[0], amin_lo = area_lo[0];
    float awmax = __fsub_rn(A[0].z, A[0].x), ahmax = __fsub_rn(A[0].w, A[0].y);
#pragma unroll
    for (int a = 1; a < NA; ++a) {
        amin = fminf(amin, area_a[a]);
        amax = fmaxf(amax, area_a[a]);
        amin_lo = fminf(amin_lo, area_lo[a]);
        awmax = fmaxf(awmax, __fsub_rn(A[a].z, A[a].x));
        ahmax = fmaxf(ahmax, __fsub_rn(A[a].w, A[a].y));
    }
    const int lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < n_cand; c0 += 32) {
        // warp-level cull: one candidate per lane against the warp's anchor bounding box
        bool hit = false;
        if (c0 + lane < n_cand) {
            const float4 gt = s.box[c0 + lane];
            const float ox = __fsub_rn(fminf(wreg.z, gt.z), fmaxf(wreg.x, gt.x));
            const float oy = __fsub_rn(fminf(wreg.w, gt.w), fmaxf(wreg.y, gt.y));
            hit = ox > 0.f && oy > 0.f;
            if (!EXACT && hit) {
                // production scan: drop the candidate for the whole patch when no anchor of this
                // level can reach the IoU floor with it.  (a) area ratio, the per-anchor test below
                // taken over all anchors; (b) IoU <= min(ox, aw_max) * min(oy, ah_max) /
                // max(gt area, smallest anchor area), with the overlap measured against the patch's
                // bounding box.  The floor sits 5 % under the threshold, far above rounding error.
                const float garea = s.area[c0 + lane];
                if (garea < amin_lo || kIouFloor * garea > amax) hit = false;
                else if (fminf(ox, awmax) * fminf(oy, ahmax) < kIouFloor * fmaxf(garea, amin))
                    hit = false;
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        while (m) {
            const int k = c0 + __ffs(m) - 1;
            m &= m - 1;
            const float4 gt = s.box[k];
            const float garea = s.area[k];
            const float garea_lo = kIouFloor * garea;
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                if (!EXACT && (garea < area_lo[a] || garea_lo > area_a[a])) continue;
                // IoU (losses.py:54-70); strict '>' in GT order == first maximum (:357)
                const float mnx = fminf(A[a].z, gt.z), mxx = fmaxf(A[a].x, gt.x);
                const float mny = fminf(A[a].w, gt.w), mxy = fmaxf(A[a].y, gt.y);
                if (mnx > mxx && mny > mxy) {
                    const float ov = __fmul_rn(__fsub_rn(mnx, mxx), __fsub_rn(mny, mxy));
                    const float un = fmaxf(__fsub_rn(__fadd_rn(area_a[a], garea), ov), 1e-4f);
                    if (!EXACT && ov < kIouFloor * un) continue;
                    const float iou = __fdiv_rn(ov, un);
                    if (iou > best[a]) {
                        best[a] = iou;
                        best_slot[a] = k;
                    }
                }
            }
        }
    }
}

template <int PL>
__global__ void __launch_bounds__(kAssignThreads)
    retina_assign_kernel(Geo g, BaseAnchors ba, TileTab tt, IouThresholds thr,
                         const float *__restrict__ annots, int G, int *__restrict__ labels,
                         int *__restrict__ matched, Queues q, int *__restrict__ npos_partials,
                         int b0) {
    constexpr int NA = PL > 0 ? PL : 1;
    constexpr int kQueue = kAssignThreads * (PL > 0 ? PL : kMaxPerLoc);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ float red[4 * kAssignWarps];
    __shared__ int warp_cnt[2 * kAssignWarps];
    __shared__ float region[4];
    __shared__ int2 pos_q[kQueue];
    __shared__ int ign_q[kQueue];
    __shared__ int n_pos_s, n_ign_s, bases[2];
    if (threadIdx.x == 0) {
        n_pos_s = 0;
        n_ign_s = 0;
    }

    const int b = blockIdx.y + b0;
    const GtSmem s = carve(smem_raw, G);
    const float *src = annots + (size_t)b * G * 5;
    const bool bulk = (((G * 5 * 4) & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    stage_rows_begin(s.raw, src, G * 5, &mbar, bulk);

    const Loc me = my_location(g, tt);
    const int per_loc = PL > 0 ? PL : g.per_loc;
    const float sx = shift_of(me.x, g.stride[me.l]), sy = shift_of(me.y, g.stride[me.l]);
    // bounding box of this location's anchors (models/anchor.py:59-86: base + shift in float32)
    const float tx1 = __fadd_rn(tt.ext[me.l][0], sx), ty1 = __fadd_rn(tt.ext[me.l][1], sy);
    const float tx2 = __fadd_rn(tt.ext[me.l][2], sx), ty2 = __fadd_rn(tt.ext[me.l][3], sy);
    const float4 wreg = reduce_region(me.active, tx1, ty1, tx2, ty2, red, region);
    stage_rows_wait(&mbar, bulk);  // contains a __syncthreads(): region[] is visible too
    const int2 cnt = compact_gt(s, G, kCullOverlap, region[0], region[1], region[2], region[3],
                                false, warp_cnt);
    const int n_cand = cnt.x;
    const bool has_gt = cnt.y > 0;

    const long long lm0 = lm_index(g, b, me.l, me.loc * per_loc);
    for (int a0 = 0; a0 < per_loc; a0 += NA) {   // one trip when PL > 0
        float4 A[NA];
        float best[NA];
        int best_slot[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            A[a].x = __fadd_rn(ba.v[me.l][a0 + a][0], sx);
            A[a].y = __fadd_rn(ba.v[me.l][a0 + a][1], sy);
            A[a].z = __fadd_rn(ba.v[me.l][a0 + a][2], sx);
            A[a].w = __fadd_rn(ba.v[me.l][a0 + a][3], sy);
        }
        if (matched)
            scan_candidates<NA, true>(s, n_cand, wreg, A, best, best_slot, thr.floor);
        else
            scan_candidates<NA, false>(s, n_cand, wreg, A, best, best_slot, thr.floor);
        // Most patches see no GT above the IoU floor: every anchor is background (label 0)
        bool any_hit = false;
#pragma unroll
        for (int a = 0; a < NA; ++a) any_hit |= best_slot[a] >= 0;
        if (!matched && has_gt && !__any_sync(0xffffffffu, any_hit)) {
            if (me.active) {
#pragma unroll
                for (int a = 0; a < NA; ++a) labels[lm0 + a0 + a] = 0;
            }
            continue;
        }
        // labels (losses.py:358-365)
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            int label = -1, match = -1, grow = -1;
            if (has_gt) {
                match = best_slot[a] >= 0 ? s.fidx[best_slot[a]] : 0;
                if (best[a] < thr.neg) label = 0;
                if (best[a] >= thr.pos) {
                    label = s.label[best_slot[a]];
                    grow = s.ridx[best_slot[a]];
                }
            }
            if (me.active) {
                const int row = (int)(lm0 + a0 + a);
                labels[row] = label;
                if (matched) matched[row] = match;
                if (label > 0) {
                    const int slot = atomicAdd(&n_pos_s, 1);
                    B200DET_ASSERT(slot < kQueue);   // one slot per (thread, anchor)
                    pos_q[slot] = make_int2(row, grow);
                }
                if (label < 0) {
                    const int slot = atomicAdd(&n_ign_s, 1);
                    B200DET_ASSERT(slot < kQueue);
                    ign_q[slot] = row;
                }
            }
        }
    }
    __syncthreads();
    flush_queues(q, pos_q, n_pos_s, ign_q, n_ign_s,
                 npos_partials + (size_t)b * gridDim.x + blockIdx.x, bases);
}

// ---------------------------------------------------------------------------------------
// Retina, production scan (no `matched` output), r02: one CTA per TILE of one level, GT-centric.
//
// The anchor-centric kernel above costs ~100 thread instructions per anchor whatever the image holds
// (anchor generation, per-CTA staging + compaction of the GT rows repeated by 114 CTAs per image,
// per-warp culling): 1 us per image of ALU time that the HBM-bound sweep running beside it has to
// absorb.  Only anchors that reach the IoU floor (0.38) with some GT can end up anything but
// background, and those are ~1 % of the 120 087: so here a CTA owns a tile of <= 32 x 32 locations of
// one level, keeps one 64-bit key (IoU bits | ~candidate index) per anchor of the tile in shared
// memory, lets every WARP take GT boxes and visit only the (anchor shape, location) pairs that can
// reach the floor with it -- area ratio within the floor, overlap per axis >= floor * max(area) /
// min(height) -- and atomicMax the pair's exact IoU into the anchor's key (largest IoU wins, equal
// IoUs -> lowest GT index: the reference's first maximum, losses.py:357).  A final pass turns the keys
// into labels, two anchors per thread, and queues the few positive / ignored rows.  The IoU itself is
// the same float32 op sequence as in scan_candidates; the candidate ranges are conservative supersets
// (0.5 % slack on the floor, one location of slack per side), so labels are bit-identical.
// ---------------------------------------------------------------------------------------
#ifndef B200DET_TILE_THREADS
#define B200DET_TILE_THREADS 256
#endif
#ifdef B200DET_TILE_MINB   // resident CTAs per SM to compile for (register cap); default: no cap
#define B200DET_TILE_BOUNDS __launch_bounds__(B200DET_TILE_THREADS, B200DET_TILE_MINB)
#else
#define B200DET_TILE_BOUNDS __launch_bounds__(B200DET_TILE_THREADS)
#endif
constexpr int kTileThreads = B200DET_TILE_THREADS;
constexpr int kTileWarps = kTileThreads / 32;
#ifndef B200DET_TILE_SIDE
#define B200DET_TILE_SIDE 32
#endif
constexpr int kTileSide = B200DET_TILE_SIDE;   // largest tile side in locations
constexpr int kTileItems = 1024;   // (GT box, anchor shape) work items per round
constexpr int kTileSideMin = 16;          // smallest tile side the host may choose (sizes the workspace)
constexpr int kTileSideSmallBatch = 24;   // tile side for batches <= kTileSmallBatch (more, smaller CTAs)
constexpr int kTileSmallBatch = 4;        // measured (tools/r02c_run10.sh, r02c_run13.sh): criterion-only calls
                                          // 0.0495 -> 0.0422 ms at 1 image, 0.072 -> 0.063 at 4, but 0.066 -> 0.070
                                          // at 8 and 0.110 -> 0.127 at 16 (BASELINE configs[1]); beside the FUSED
                                          // sweep the small tiles also win at 16 images (0.229 -> 0.210 ms per step)

struct BigTiles {
    int tile_off[kMaxLevels + 1];        // tiles of one image before level l
    int nx[kMaxLevels];                  // tiles per row of level l
    int tw[kMaxLevels], th[kMaxLevels];  // tile size in locations
    float ext[kMaxLevels][4];            // extent of the level's base anchors (see TileTab)
    int max_anchors;                     // largest tw * th * per_loc
};

__device__ __forceinline__ float shift_f32(int i, float stride) {
    // (i + 0.5) * stride rounded once: i + 0.5 is exact in float32 and so is the product in float64,
    // hence this equals shift_of()'s float64 product rounded to float32
    return __fmul_rn(__fadd_rn(__int2float_rn(i), 0.5f), stride);
}

__global__ void B200DET_TILE_BOUNDS
    retina_assign_tile_kernel(Geo g, BaseAnchors ba, BigTiles bt, IouThresholds thr,
                              const float *__restrict__ annots, int G, int *__restrict__ labels,
                              Queues q, int *__restrict__ npos_partials, int blocks_per_image) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int warp_cnt[2 * kTileWarps];
    __shared__ int s_npos, s_items;
    __shared__ int2 items[kTileItems];
    __shared__ int touched[kTileSide * ((kTileSide * kMaxPerLoc + 127) / 128)];   // per 128-anchor run of a row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i)
        if (i < g.n_levels && (int)blockIdx.x >= bt.tile_off[i]) l = i;
    const int t = blockIdx.x - bt.tile_off[l];
    const int tyi = t / bt.nx[l], txi = t - tyi * bt.nx[l];
    const int x0 = txi * bt.tw[l], y0 = tyi * bt.th[l];
    const int tw = min(bt.tw[l], g.W[l] - x0), th = min(bt.th[l], g.H[l] - y0);
    const int per_loc = g.per_loc;
    const int row_len = tw * per_loc, n_anch = row_len * th;
    const float stride = g.stride[l];

    const GtSmem s = carve(smem_raw, G);
    unsigned long long *keys =
        reinterpret_cast<unsigned long long *>(smem_raw + ((gt_smem_bytes(G) + 15) & ~(size_t)15));
    const float *src = annots + (size_t)b * G * 5;
    const bool bulk = (((G * 5 * 4) & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    stage_rows_begin(s.raw, src, G * 5, &mbar, bulk);
    for (int i = tid; i < n_anch; i += kTileThreads) keys[i] = 0ull;
    const int runs_per_row = (row_len + 127) >> 7;
    for (int i = tid; i < th * runs_per_row; i += kTileThreads) touched[i] = 0;
    if (tid == 0) s_npos = 0;
    // exact bounding box of the tile's anchors (rounding is monotonic, see TileTab::ext)
    const float rx1 = __fadd_rn(bt.ext[l][0], shift_f32(x0, stride));
    const float ry1 = __fadd_rn(bt.ext[l][1], shift_f32(y0, stride));
    const float rx2 = __fadd_rn(bt.ext[l][2], shift_f32(x0 + tw - 1, stride));
    const float ry2 = __fadd_rn(bt.ext[l][3], shift_f32(y0 + th - 1, stride));
    stage_rows_wait(&mbar, bulk);   // contains a __syncthreads()
    const int2 cnt = compact_gt<kTileThreads>(s, G, kCullOverlap, rx1, ry1, rx2, ry2, false, warp_cnt);
    const int n_cand = cnt.x;
    const bool has_gt = cnt.y > 0;

    // ---- scatter, step 1: one THREAD per (GT box, anchor shape) works out the range of locations
    // that can reach the floor and queues the non-empty ones; step 2: one WARP per queued item,
    // lanes over the item's locations ----
    const float fl = thr.floor, fs = 0.995f * thr.floor;   // exact floor / floor with slack
    const float inv_stride = 1.f / stride;
    const int chunk = kTileItems / per_loc;                 // GT boxes per round
    for (int c0 = 0; c0 < n_cand; c0 += chunk) {
        const int nc = min(chunk, n_cand - c0);
        if (tid == 0) s_items = 0;
        __syncthreads();
        for (int idx = tid; idx < nc * per_loc; idx += kTileThreads) {
            const int ci = idx / per_loc, a = idx - ci * per_loc, c = c0 + ci;
            const float4 gt = s.box[c];
            const float garea = s.area[c];
            if (!(garea > 0.f)) continue;   // IoU is exactly 0 with everything
            const float gw = gt.z - gt.x, gh = gt.w - gt.y;
            const float bx = ba.v[l][a][0], by = ba.v[l][a][1], bz = ba.v[l][a][2], bw = ba.v[l][a][3];
            const float aw = bz - bx, ah = bw - by, aarea = aw * ah;   // (shifted anchors: same up to ulps)
            // IoU <= min(area) / max(area)
            if (garea < fs * aarea || fs * garea > aarea) continue;
            // ox * oy >= floor * max(area) with oy <= min(ah, gh)  =>  ox >= need / min(ah, gh)
            const float need = fs * fmaxf(aarea, garea);
            const float oxmin = need / fminf(ah, gh), oymin = need / fminf(aw, gw);
            // anchor at location x spans [sx + bx, sx + bz], sx = (x + 0.5) * stride
            const float sxl = (gt.x + oxmin - bz) * inv_stride - 0.5f, sxh = (gt.z - oxmin - bx) * inv_stride - 0.5f;
            const float syl = (gt.y + oymin - bw) * inv_stride - 0.5f, syh = (gt.w - oymin - by) * inv_stride - 0.5f;
            if (!(sxl <= sxh) || !(syl <= syh)) continue;   // also catches NaN boxes
            const int xlo = max(x0, (int)fminf(fmaxf(ceilf(sxl) - 1.f, -1.f), 2.0e6f));
            const int xhi = min(x0 + tw - 1, (int)fmaxf(fminf(floorf(sxh) + 1.f, 2.0e6f), -2.f));
            const int ylo = max(y0, (int)fminf(fmaxf(ceilf(syl) - 1.f, -1.f), 2.0e6f));
            const int yhi = min(y0 + th - 1, (int)fmaxf(fminf(floorf(syh) + 1.f, 2.0e6f), -2.f));
            if (xlo > xhi || ylo > yhi) continue;
            items[atomicAdd(&s_items, 1)] =
                make_int2(c | (a << 16), (xlo - x0) | ((xhi - x0) << 8) | ((ylo - y0) << 16) | ((yhi - y0) << 24));
        }
        __syncthreads();
        const int n_items = s_items;
        for (int it = warp; it < n_items; it += kTileWarps) {
            const int2 e = items[it];
            const int c = e.x & 0xffff, a = e.x >> 16;
            const int xlo = x0 + (e.y & 0xff), xhi = x0 + ((e.y >> 8) & 0xff);
            const int ylo = y0 + ((e.y >> 16) & 0xff), yhi = y0 + ((e.y >> 24) & 0xff);
            const float4 gt = s.box[c];
            const float garea = s.area[c];
            const float bx = ba.v[l][a][0], by = ba.v[l][a][1], bz = ba.v[l][a][2], bw = ba.v[l][a][3];
            const unsigned long long key_lo = 0xffffffffull - (unsigned long long)c;
            const int nx = xhi - xlo + 1, np = nx * (yhi - ylo + 1);
            const float rnx = 1.f / (float)nx;
            for (int p = lane; p < np; p += 32) {
                const int py = (int)(((float)p + 0.5f) * rnx);   // p / nx for these small integers
                const int xx = xlo + p - py * nx, yy = ylo + py;
                const float sx = shift_f32(xx, stride), sy = shift_f32(yy, stride);
                const float ax1 = __fadd_rn(bx, sx), ax2 = __fadd_rn(bz, sx);   // anchor.py:80
                const float ay1 = __fadd_rn(by, sy), ay2 = __fadd_rn(bw, sy);
                // IoU exactly as scan_candidates (losses.py:54-70)
                const float area_a = __fmul_rn(fmaxf(__fsub_rn(ax2, ax1), 0.f),
                                               fmaxf(__fsub_rn(ay2, ay1), 0.f));
                if (garea < fl * area_a || fl * garea > area_a) continue;
                const float mnx = fminf(ax2, gt.z), mxx = fmaxf(ax1, gt.x);
                const float mny = fminf(ay2, gt.w), mxy = fmaxf(ay1, gt.y);
                if (mnx > mxx && mny > mxy) {
                    const float ov = __fmul_rn(__fsub_rn(mnx, mxx), __fsub_rn(mny, mxy));
                    const float un = fmaxf(__fsub_rn(__fadd_rn(area_a, garea), ov), 1e-4f);
                    if (ov < fl * un) continue;
                    const float iou = __fdiv_rn(ov, un);
                    const int j = (xx - x0) * per_loc + a;
                    atomicMax(keys + (yy - y0) * row_len + j,
                              ((unsigned long long)__float_as_uint(iou) << 32) | key_lo);
                    touched[(yy - y0) * runs_per_row + (j >> 7)] = 1;   // (benign race: all write 1)
                }
            }
        }
        __syncthreads();
    }

    // ---- labels (losses.py:358-365): a warp per tile row, 4 x 32 anchors per lane round ----
    int n_pos_mine = 0;
    for (int r = warp; r < th; r += kTileWarps) {
        const long long grow0 = lm_index(g, b, l, ((y0 + r) * g.W[l] + x0) * per_loc);
        const unsigned long long *krow = keys + r * row_len;
        for (int j0 = 0; j0 < row_len; j0 += 128) {
            if (has_gt && !touched[r * runs_per_row + (j0 >> 7)]) {
                // no pair of this run reached the floor: background (the common case by far)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u + lane;
                    if (j < row_len) labels[grow0 + j] = 0;
                }
                continue;
            }
            int label[4], grow[4];
            bool quiet = true;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 32 * u + lane;
                label[u] = 0;
                grow[u] = -1;
                if (j < row_len) {
                    if (has_gt) {
                        const unsigned long long k = krow[j];
                        if (k != 0ull) {
                            const float best = __uint_as_float((uint32_t)(k >> 32));
                            if (!(best < thr.neg)) label[u] = -1;
                            if (best >= thr.pos) {
                                const uint32_t slot = 0xffffffffu - (uint32_t)k;
                                label[u] = s.label[slot];
                                grow[u] = s.ridx[slot];
                            }
                        }
                    } else {
                        label[u] = -1;   // image without GT: every anchor ignored (losses.py:341-345)
                    }
                    labels[grow0 + j] = label[u];
                    quiet = quiet && label[u] == 0;
                }
            }
            if (__all_sync(0xffffffffu, quiet)) continue;   // background only: the common case
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 32 * u + lane;
                const bool valid = j < row_len;
                const bool is_pos = valid && label[u] > 0, is_ign = valid && label[u] < 0;
                const unsigned mp = __ballot_sync(0xffffffffu, is_pos);
                const unsigned mi = __ballot_sync(0xffffffffu, is_ign);
                const unsigned lower = (1u << lane) - 1u;
                if (mp) {   // one global atomic per warp and visit
                    int base = 0;
                    if (lane == 0) base = atomicAdd(q.counters + 0, __popc(mp));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (is_pos) q.pos[base + __popc(mp & lower)] = make_int2((int)(grow0 + j), grow[u]);
                    n_pos_mine += is_pos;
                }
                if (mi) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(q.counters + 1, __popc(mi));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (is_ign) q.ign[base + __popc(mi & lower)] = (int)(grow0 + j);
                }
            }
        }
    }
    n_pos_mine = warp_sum_int(n_pos_mine);
    if (lane == 0 && n_pos_mine) atomicAdd(&s_npos, n_pos_mine);
    __syncthreads();
    int *part = npos_partials + (size_t)b * blocks_per_image;
    if (tid == 0) part[blockIdx.x] = s_npos;
    // the workspace has one slot per CTA of the anchor-centric kernel: clear the ones this grid lacks
    if (blockIdx.x == 0)
        for (int i = gridDim.x + tid; i < blocks_per_image; i += kTileThreads) part[i] = 0;
}

// ---------------------------------------------------------------------------------------
// FCOS: point <-> GT with centre sampling, scale range, min-area  (losses.py:612-833)
// ---------------------------------------------------------------------------------------
struct FcosPick {
    int slot;
    float l, t, r, b;
};

// losses.py:688-735 (candidate tests) for one (point, GT) pair
__device__ __forceinline__ bool fcos_candidate(const float2 pt, const float4 gt, float m0, float m1,
                                               float rad, int use_center_sample, float &cl,
                                               float &ct, float &cr, float &cb) {
    cl = __fsub_rn(pt.x, gt.x);
    ct = __fsub_rn(pt.y, gt.y);
    cr = __fsub_rn(gt.z, pt.x);
    cb = __fsub_rn(gt.w, pt.y);
    const float mn = fminf(fminf(cl, ct), fminf(cr, cb));
    bool ok = mn > 0.f;
    if (ok && use_center_sample) {
        const float cx = __fdiv_rn(__fadd_rn(gt.z, gt.x), 2.f);
        const float cy = __fdiv_rn(__fadd_rn(gt.w, gt.y), 2.f);
        const float dx = __fsub_rn(pt.x, cx), dy = __fsub_rn(pt.y, cy);
        const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        ok = d < rad;
    }
    if (ok) {
        const float mx = fmaxf(fmaxf(cl, ct), fmaxf(cr, cb));
        ok = (mx > m0) && (mx < m1);
    }
    return ok;
}

// centre-ness target, losses.py:822-824
__device__ __forceinline__ float fcos_centerness(float l, float t, float r, float b) {
    return __fsqrt_rn(__fmul_rn(__fdiv_rn(fminf(l, r), fmaxf(l, r)),
                                __fdiv_rn(fminf(t, b), fmaxf(t, b))));
}

__global__ void __launch_bounds__(kAssignThreads)
    fcos_assign_kernel(Geo g, FcosTab ft, TileTab tt, const float *__restrict__ annots, int G,
                       int use_center_sample, int *__restrict__ labels,
                       int *__restrict__ matched, float *__restrict__ targets, Queues q,
                       int *__restrict__ npos_partials, int b0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ float red[4 * kAssignWarps];
    __shared__ int warp_cnt[2 * kAssignWarps];
    __shared__ float region[4];
    __shared__ int2 pos_q[kAssignThreads];
    __shared__ int n_pos_s, bases[2];
    if (threadIdx.x == 0) n_pos_s = 0;

    const int b = blockIdx.y + b0;
    const GtSmem s = carve(smem_raw, G);
    const float *src = annots + (size_t)b * G * 5;
    const bool bulk = (((G * 5 * 4) & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    stage_rows_begin(s.raw, src, G * 5, &mbar, bulk);

    const Loc me = my_location(g, tt);
    const float2 pt = make_float2(shift_of(me.x, g.stride[me.l]), shift_of(me.y, g.stride[me.l]));
    const float m0 = ft.mi_lo[me.l], m1 = ft.mi_hi[me.l], rad = ft.radius[me.l];
    const float4 wreg = reduce_region(me.active, pt.x, pt.y, pt.x, pt.y, red, region);
    stage_rows_wait(&mbar, bulk);
    // only GT boxes that can strictly contain one of the CTA's points can be candidates
    const int2 cnt = compact_gt(s, G, kCullContain, region[0], region[1], region[2], region[3],
                                true, warp_cnt);
    const int n_cand = cnt.x;

    // smallest area among the candidates, first minimum (losses.py:785-808)
    float best_area = __int_as_float(0x7f800000);
    int best = -1;
    float bl = 0.f, bt = 0.f, br = 0.f, bb = 0.f;
    const int lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < n_cand; c0 += 32) {
        bool hit = false;
        if (c0 + lane < n_cand) {
            const float4 gt = s.box[c0 + lane];
            hit = (gt.x < wreg.z) && (gt.z > wreg.x) && (gt.y < wreg.w) && (gt.w > wreg.y);
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        while (m) {
            const int k = c0 + __ffs(m) - 1;
            m &= m - 1;
            float cl, ct, cr, cb;
            const bool ok = fcos_candidate(pt, s.box[k], m0, m1, rad, use_center_sample, cl, ct, cr, cb);
            if (ok && s.area[k] < best_area) {
                best_area = s.area[k];
                best = k;
                bl = cl;
                bt = ct;
                br = cr;
                bb = cb;
            }
        }
    }
    const bool pos = me.active && best >= 0;
    if (me.active) {
        const long long lm = lm_index(g, b, me.l, me.loc);
        const int label = pos ? s.label[best] : 0;
        labels[lm] = label;
        if (pos) pos_q[atomicAdd(&n_pos_s, 1)] = make_int2((int)lm, s.ridx[best]);
        if (matched) matched[lm] = pos ? s.fidx[best] : -1;
        if (targets) {
            float *t = targets + lm * 6;
            t[0] = bl;
            t[1] = bt;
            t[2] = br;
            t[3] = bb;
            t[4] = (float)label;
            t[5] = pos ? fcos_centerness(bl, bt, br, bb) : 0.f;
        }
    }
    __syncthreads();
    flush_queues(q, pos_q, n_pos_s, nullptr, 0,
                 npos_partials + (size_t)b * gridDim.x + blockIdx.x, bases);
}

// ---------------------------------------------------------------------------------------
// sparse losses: box / centre-ness loss of the positives (+ gradients), focal corrections
// ---------------------------------------------------------------------------------------
// Box loss of one positive anchor (losses.py:263-320) and d(loss)/d(reg row).
__device__ __forceinline__ float retina_box_term(const float4 a, const float4 gt, const float4 t,
                                                 int box_loss, float beta, float4 &grad,
                                                 int exp_mode) {
    const float awx = __fsub_rn(a.z, a.x), awy = __fsub_rn(a.w, a.y);
    const float acx = __fadd_rn(a.x, __fmul_rn(0.5f, awx));
    const float acy = __fadd_rn(a.y, __fmul_rn(0.5f, awy));
    float term = 0.f;
    if (box_loss == B200DET_BOX_SMOOTHL1) {
        // targets (losses.py:390-409)
        const float gwx = fmaxf(__fsub_rn(gt.z, gt.x), 1e-4f);
        const float gwy = fmaxf(__fsub_rn(gt.w, gt.y), 1e-4f);
        const float gcx = __fadd_rn(gt.x, __fmul_rn(0.5f, gwx));
        const float gcy = __fadd_rn(gt.y, __fmul_rn(0.5f, gwy));
        const float tg[4] = {__fdiv_rn(__fsub_rn(gcx, acx), awx), __fdiv_rn(__fsub_rn(gcy, acy), awy),
                             logf(__fdiv_rn(gwx, awx)), logf(__fdiv_rn(gwy, awy))};
        const float pr[4] = {t.x, t.y, t.z, t.w};
        float gr[4];
        const float half_beta = 0.5f * beta;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float d = __fsub_rn(pr[i], tg[i]);
            const float x = fabsf(d);
            if (x >= beta) {
                term += __fsub_rn(x, half_beta);
                gr[i] = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            } else {
                term += __fdiv_rn(__fmul_rn(0.5f, __fmul_rn(x, x)), beta);
                gr[i] = d / beta;
            }
        }
        grad = make_float4(gr[0], gr[1], gr[2], gr[3]);
    } else {
        // decode (losses.py:411-429) then 1 - IoU-family (losses.py:286-293)
        const Dual tx = dvar(t.x, 0), ty = dvar(t.y, 1), tw = dvar(t.z, 2), th = dvar(t.w, 3);
        const Dual bw = dexp(tw, exp_mode) * awx, bh = dexp(th, exp_mode) * awy;
        const Dual cx = tx * awx + acx, cy = ty * awy + acy;
        const Dual hw = bw * 0.5f, hh = bh * 0.5f;
        const Dual p[4] = {cx - hw, cy - hh, cx + hw, cy + hh};
        const float gg[4] = {gt.x, gt.y, gt.z, gt.w};
        const Dual iou = iou_family(p, gg, box_loss);
        term = __fsub_rn(1.f, iou.v);
        grad = make_float4(-iou.d[0], -iou.d[1], -iou.d[2], -iou.d[3]);
    }
    return term;
}

struct SparseArgs {
    Geo g;
    BaseAnchors ba;
    PtrTab reg, ctr, cls;        // cls.p[0] == nullptr: no focal corrections
    MutPtrTab reg_grad, ctr_grad;
    int reg_dtype, box_loss, is_fcos, G, C;
    int ctr_logits;              // FCOS: ctr holds the head's logits (the logits path), not probabilities
    float beta, alpha, gamma;
};

// 64-bit fixed point: integer adds commute, so the sums are independent of the visiting order
constexpr double kFxBox = 4294967296.0;          // 2^32
constexpr double kFxFocal = 1099511627776.0;     // 2^40
// NaN / inf terms cannot be represented: they set `bad` and the CTA's partial becomes NaN, which the
// fp64 reduction propagates (the reference's float sum would be NaN or inf; the caller's training
// loop skips the step either way, tools/scripts.py:922-930).
__device__ __forceinline__ long long to_fx(float v, double scale, unsigned &bad, unsigned bit) {
    if (!(fabsf(v) <= 3.402823466e38f)) {
        bad |= bit;
        return 0;
    }
    return __double2ll_rn((double)v * scale);
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// level-major row -> level and row inside that level's [B*rows_l] tensor
// (batch * off[k] < 2^31 is checked by make_geo: 32-bit products and compares)
__device__ __forceinline__ int split_row(const Geo &g, int i, long long &rrow) {
    int l = 0, base = 0;
#pragma unroll
    for (int k = 1; k < kMaxLevels; ++k) {
        const int bk = g.batch * g.off[k];
        if (k < g.n_levels && i >= bk) {
            l = k;
            base = bk;
        }
    }
    rrow = (long long)(i - base);
    return l;
}

#ifndef B200DET_SPARSE_MINB
#define B200DET_SPARSE_MINB (1024 / B200DET_SPARSE_THREADS)   // 64 registers
#endif
__global__ void __launch_bounds__(kSparseThreads, B200DET_SPARSE_MINB)
    sparse_loss_kernel(SparseArgs a, const float *__restrict__ annots,
                       const int *__restrict__ labels, Queues q,
                       SparsePartial *__restrict__ partials) {
    const Geo &g = a.g;
    const int lane = threadIdx.x & 31;
    const bool want_fix = a.cls.p[0] != nullptr;
    const bool want_loss = a.box_loss != B200DET_BOX_NONE;
    const bool gamma2 = a.gamma == 2.f;
    const int n_pos = q.counters[0], n_ign = q.counters[1];
    const int gtid = blockIdx.x * kSparseThreads + threadIdx.x;
    const int gsize = gridDim.x * kSparseThreads;
    long long box_fx = 0, ctr_fx = 0, foc_fx = 0;
    unsigned bad = 0;  // bit 0 box, bit 1 centre-ness, bit 2 focal corrections

    // ---- positives: one thread each ----
    for (int k = gtid; k < n_pos; k += gsize) {
        const int2 e = q.pos[k];
        long long rrow;
        const int l = split_row(g, e.x, rrow);
        const int label = __ldg(labels + e.x);
        if (want_loss) {
            const int b = (int)(rrow / g.rows[l]);
            const int local = (int)(rrow - (long long)b * g.rows[l]);
            const float *gp = annots + ((size_t)b * a.G + e.y) * 5;
            const float4 gt = make_float4(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), __ldg(gp + 3));
            const float4 t = load_reg4(a.reg.p[l], a.reg_dtype, rrow);
            float4 grad = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!a.is_fcos) {
                const float4 an = anchor_of(g, a.ba, l, local);
                box_fx += to_fx(retina_box_term(an, gt, t, a.box_loss, a.beta, grad, a.reg_dtype),
                                kFxBox, bad, 1u);
            } else {
                // IoU loss, losses.py:550-586: boxes rebuilt around the point from l,t,r,b
                const float2 pt = point_of(g, l, local);
                const float bl = __fsub_rn(pt.x, gt.x), bt = __fsub_rn(pt.y, gt.y);
                const float br = __fsub_rn(gt.z, pt.x), bb = __fsub_rn(gt.w, pt.y);
                const float ctr_t = fcos_centerness(bl, bt, br, bb);
                const Dual e0 = dexp(dvar(t.x, 0), a.reg_dtype), e1 = dexp(dvar(t.y, 1), a.reg_dtype);
                const Dual e2 = dexp(dvar(t.z, 2), a.reg_dtype), e3 = dexp(dvar(t.w, 3), a.reg_dtype);
                const Dual p[4] = {pt.x - e0, pt.y - e1, e2 + pt.x, e3 + pt.y};
                const float gg[4] = {__fsub_rn(pt.x, bl), __fsub_rn(pt.y, bt), __fadd_rn(pt.x, br),
                                     __fadd_rn(pt.y, bb)};
                const Dual iou = iou_family(p, gg, a.box_loss);
                box_fx += to_fx(__fmul_rn(__fsub_rn(1.f, iou.v), ctr_t), kFxBox, bad, 1u);
                grad = make_float4(-iou.d[0] * ctr_t, -iou.d[1] * ctr_t, -iou.d[2] * ctr_t,
                                   -iou.d[3] * ctr_t);
                // centre-ness BCE, losses.py:588-610 (prob clamped at losses.py:494)
                float craw = __ldg(static_cast<const float *>(a.ctr.p[l]) + rrow);
                // logits path: torch's CUDA sigmoid of the head's centre-ness logit (models/head.py:176-179)
                if (a.ctr_logits) craw = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-craw)));
                const float cp = clamp_prob(craw);
                const float one_m = __fsub_rn(1.f, cp), one_t = __fsub_rn(1.f, ctr_t);
                ctr_fx += to_fx(-__fadd_rn(__fmul_rn(ctr_t, logf(cp)), __fmul_rn(one_t, logf(one_m))),
                                kFxBox, bad, 2u);
                if (a.ctr_grad.p[0] != nullptr) {
                    const float cgrad =
                        (craw >= kClampLo && craw <= kClampHi) ? -(ctr_t / cp - one_t / one_m) : 0.f;
                    static_cast<float *>(a.ctr_grad.p[l])[rrow] = cgrad;
                }
            }
            if (a.reg_grad.p[0] != nullptr) reinterpret_cast<float4 *>(a.reg_grad.p[l])[rrow] = grad;
        }
        if (want_fix) {
            // the sweep counted the target class as background: swap the term
            const float p = __ldg(static_cast<const float *>(a.cls.p[l]) + rrow * a.C + (label - 1));
            foc_fx += to_fx(a.alpha * pos_term(p, a.gamma, gamma2) -
                                (1.f - a.alpha) * neg_term(p, a.gamma, gamma2),
                            kFxFocal, bad, 4u);
        }
    }

    // ---- ignored rows (Retina: 0.4 <= IoU < 0.5, or image without GT) take no part in the focal
    // loss (losses.py:228-230): remove what the label-free sweep added.  Flat grid-stride loop over
    // (ignored row, 16-byte unit) pairs, four independent pairs in flight per thread; every unit's
    // four terms are summed in a fixed order and converted to fixed point on their own, so the
    // total does not depend on the queue order or on which thread visits which unit.
    if (want_fix) {
        const bool vec = (a.C & 3) == 0;
        const int U = vec ? (a.C >> 2) : a.C;              // units per row
        const long long n_pairs = (long long)n_ign * U;
        const float one_m_alpha = 1.f - a.alpha;
        constexpr int kInFlight = 4;
        // p / U by multiply-high while the pair index fits 31 bits (it does up to ~100 M ignored
        // class scores; the 64-bit division was 10 % of this kernel's instructions)
        const bool small = n_pairs < (1ll << 31);
        int sft = 0;
        while ((1u << sft) < (unsigned)U) ++sft;
        const unsigned magic = U > 1 ? (unsigned)(((1ull << (31 + sft)) / (unsigned long long)U) + 1ull) : 0u;
        for (long long p0 = gtid; p0 < n_pairs; p0 += (long long)kInFlight * gsize) {
            float4 v[kInFlight];
            bool live[kInFlight];
#pragma unroll
            for (int t = 0; t < kInFlight; ++t) {
                const long long p = p0 + (long long)t * gsize;
                live[t] = p < n_pairs;
                v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live[t]) {
                    const int r = U == 1 ? (int)p
                                         : (small ? (int)(__umulhi((unsigned)p, magic) >> (sft - 1)) : (int)(p / U));
                    const int u = (int)(p - (long long)r * U);
                    long long rr;
                    const int l = split_row(g, __ldg(q.ign + r), rr);
                    const float *row = static_cast<const float *>(a.cls.p[l]) + rr * a.C;
                    if (vec) v[t] = __ldg(reinterpret_cast<const float4 *>(row) + u);
                    else v[t].x = __ldg(row + u);
                }
            }
#pragma unroll
            for (int t = 0; t < kInFlight; ++t) {
                if (!live[t]) continue;
                float sum = neg_term(v[t].x, a.gamma, gamma2);
                if (vec) {
                    sum += neg_term(v[t].y, a.gamma, gamma2);
                    sum += neg_term(v[t].z, a.gamma, gamma2);
                    sum += neg_term(v[t].w, a.gamma, gamma2);
                }
                foc_fx -= to_fx(one_m_alpha * sum, kFxFocal, bad, 4u);
            }
        }
    }

    __shared__ long long red[3][kSparseThreads / 32];
    const int warp = threadIdx.x >> 5;
    box_fx = warp_sum_ll(box_fx);
    ctr_fx = warp_sum_ll(ctr_fx);
    foc_fx = warp_sum_ll(foc_fx);
    if (lane == 0) {
        red[0][warp] = box_fx;
        red[1][warp] = ctr_fx;
        red[2][warp] = foc_fx;
    }
    // __syncthreads_or returns a truth value, not the bitwise OR: one vote per loss
    const bool bad_box = __syncthreads_or((int)(bad & 1u)) != 0;
    const bool bad_ctr = __syncthreads_or((int)(bad & 2u)) != 0;
    const bool bad_foc = __syncthreads_or((int)(bad & 4u)) != 0;
    if (threadIdx.x == 0) {
        long long sb = 0, sc = 0, sf = 0;
        for (int w = 0; w < kSparseThreads / 32; ++w) {
            sb += red[0][w];
            sc += red[1][w];
            sf += red[2][w];
        }
        SparsePartial p;
        p.box = (double)sb / kFxBox;
        p.ctr = (double)sc / kFxBox;
        p.focal = (double)sf / kFxFocal;
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        if (bad_box) p.box = nan;
        if (bad_ctr) p.ctr = nan;
        if (bad_foc) p.focal = nan;
        partials[blockIdx.x] = p;
    }
}

// ---------------------------------------------------------------------------------------
// utilities: materialise rows (tests), level-major -> image-major
// ---------------------------------------------------------------------------------------
// Autograd backward of the box / centre-ness losses: the forward wrote their gradients at the rows of
// the positives only (into zeroed buffers); scaling them by upstream * weight / positives therefore
// only has to visit the positive-row queue the assignment left in the workspace -- a few thousand
// rows instead of a pass over the whole [B*N, 4] tensor (492 MB read + written at batch 256).
struct ScaleRowsArgs {
    Geo g;
    MutPtrTab reg_grad, ctr_grad;
    const float *g_box, *g_ctr;   // upstream scalars (device), NULL: that head is not scaled
    const double *sums;           // sums[0] = positives
    float w_box, w_ctr;
};
__global__ void scale_pos_rows_kernel(ScaleRowsArgs a, Queues q) {
    const double npos = a.sums[0];
    // as scale_levels_kernel: 0 without positives, NaN after a failed cross-rank exchange
    const float kb = a.g_box ? (npos > 0.0 ? (float)((double)*a.g_box * (double)a.w_box / npos)
                                           : (npos != npos ? (float)npos : 0.f))
                             : 1.f;
    const float kc = a.g_ctr ? (npos > 0.0 ? (float)((double)*a.g_ctr * (double)a.w_ctr / npos)
                                           : (npos != npos ? (float)npos : 0.f))
                             : 1.f;
    const int n_pos = q.counters[0];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_pos; k += gridDim.x * blockDim.x) {
        long long rrow;
        const int l = split_row(a.g, q.pos[k].x, rrow);
        if (a.g_box) {
            float4 *p = reinterpret_cast<float4 *>(a.reg_grad.p[l]) + rrow;
            float4 v = *p;
            v.x *= kb, v.y *= kb, v.z *= kb, v.w *= kb;
            *p = v;
        }
        if (a.g_ctr) static_cast<float *>(a.ctr_grad.p[l])[rrow] *= kc;
    }
}

__global__ void generate_rows_kernel(Geo g, BaseAnchors ba, int is_fcos, float *out) {
    const int N = g.off[g.n_levels];
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= N) return;
    const int l = level_of_row(g, row);
    const int local = row - g.off[l];
    if (is_fcos) {
        const float2 p = point_of(g, l, local);
        out[row * 2 + 0] = p.x;
        out[row * 2 + 1] = p.y;
    } else {
        const float4 a = anchor_of(g, ba, l, local);
        reinterpret_cast<float4 *>(out)[row] = a;
    }
}

__global__ void rows_to_image_major_kernel(Geo g, const uint32_t *src, uint32_t *dst, int width) {
    const int N = g.off[g.n_levels];
    const long long total = (long long)g.batch * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / N), row = (int)(i % N);
        const int l = level_of_row(g, row);
        const long long lm = lm_index(g, b, l, row - g.off[l]);
        for (int w = 0; w < width; ++w) dst[i * width + w] = src[lm * width + w];
    }
}

static TileTab make_tiles(const Geo &g) {
    TileTab t;
    int off = 0;
    for (int l = 0; l < kMaxLevels; ++l) {
        t.tile_off[l] = off;
        t.tiles_x[l] = 1;
        if (l < g.n_levels) {
            t.tiles_x[l] = (g.W[l] + kTileW - 1) / kTileW;
            off += t.tiles_x[l] * ((g.H[l] + kTileH - 1) / kTileH);
        }
    }
    for (int l = g.n_levels; l <= kMaxLevels; ++l) t.tile_off[l] = off;
    for (int l = 0; l < kMaxLevels; ++l)
        for (int k = 0; k < 4; ++k) t.ext[l][k] = 0.f;
    return t;
}
static void set_anchor_extents(TileTab *t, const Geo &g, const BaseAnchors &ba) {
    for (int l = 0; l < g.n_levels; ++l) {
        for (int k = 0; k < 4; ++k) t->ext[l][k] = ba.v[l][0][k];
        for (int a = 1; a < g.per_loc; ++a) {
            t->ext[l][0] = fminf(t->ext[l][0], ba.v[l][a][0]);
            t->ext[l][1] = fminf(t->ext[l][1], ba.v[l][a][1]);
            t->ext[l][2] = fmaxf(t->ext[l][2], ba.v[l][a][2]);
            t->ext[l][3] = fmaxf(t->ext[l][3], ba.v[l][a][3]);
        }
    }
}

// side: largest tile side in locations, <= kTileSide (the kernel's static arrays are sized for it)
static BigTiles make_big_tiles(const Geo &g, int side = kTileSide) {
    BigTiles t;
    int off = 0;
    t.max_anchors = 0;
    for (int l = 0; l < kMaxLevels; ++l) {
        t.tile_off[l] = off;
        t.nx[l] = t.tw[l] = t.th[l] = 1;
        for (int k = 0; k < 4; ++k) t.ext[l][k] = 0.f;
        if (l < g.n_levels) {
            // balanced tiles of at most kTileSide x kTileSide locations (100 -> 4 x 25, 50 -> 2 x 25)
            const int nx = (g.W[l] + side - 1) / side, ny = (g.H[l] + side - 1) / side;
            t.nx[l] = nx;
            t.tw[l] = (g.W[l] + nx - 1) / nx;
            t.th[l] = (g.H[l] + ny - 1) / ny;
            off += nx * ((g.H[l] + t.th[l] - 1) / t.th[l]);
            const int anchors = t.tw[l] * t.th[l] * g.per_loc;
            if (anchors > t.max_anchors) t.max_anchors = anchors;
        }
    }
    for (int l = g.n_levels; l <= kMaxLevels; ++l) t.tile_off[l] = off;
    return t;
}

static Queues queues_of(char *base, const LossWs &ws) {
    Queues q;
    q.counters = reinterpret_cast<int *>(base + ws.off_counters);
    q.pos = reinterpret_cast<int2 *>(base + ws.off_pos_queue);
    q.ign = reinterpret_cast<int *>(base + ws.off_ign_queue);
    return q;
}

// Optional residency cap for the assignment kernels (B200DET_ASSIGN_CTAS_PER_SM, default: none).
// When a caller overlaps them with the HBM-bound sweep on another stream, their register
// footprint (up to 92 x 128 per CTA, five CTAs per SM) can lock the sweep out of the SMs; padding
// the dynamic shared-memory request limits how many fit.  Measured on B200 (profiles/): running
// the kernels back to back is faster than any overlap, so the default is no cap.
constexpr size_t kAssignSmemBudget = 200 * 1024;
static size_t assign_dyn_smem(int max_gt) {
    static int ctas = 0;
    if (ctas == 0) {
        const char *e = getenv("B200DET_ASSIGN_CTAS_PER_SM");
        ctas = e ? atoi(e) : 16;
        if (ctas < 1) ctas = 1;
        if (ctas > 16) ctas = 16;
    }
    const size_t need = gt_smem_bytes(max_gt);
    const size_t pad = kAssignSmemBudget / (size_t)ctas;   // static smem (<= 26 KB) comes on top
    return need > pad ? need : pad;
}

// CTAs per image of the assignment kernels / CTAs of the sparse kernel (workspace layout)
int assign_blocks_per_image(const Geo &g) {
    // one positive-count slot per CTA of whichever assignment kernel runs (the tile kernel clears
    // the slots it does not use)
    const TileTab t = make_tiles(g);
    const int anchor_centric = (t.tile_off[g.n_levels] + kAssignWarps - 1) / kAssignWarps;
    // one partial-count slot per CTA of whichever kernel / tile side runs
    const int tiles = std::max(make_big_tiles(g).tile_off[g.n_levels],
                               make_big_tiles(g, kTileSideMin).tile_off[g.n_levels]);
    return anchor_centric > tiles ? anchor_centric : tiles;
}
int sparse_blocks(const Geo &g) {
    const long long total = (long long)g.batch * g.off[g.n_levels];
    const long long want = (total + kSparseThreads - 1) / kSparseThreads;
    const long long cap = 148 * 8 * (256 / kSparseThreads);   // 2048 threads per SM's worth
    return (int)(want < cap ? want : cap);
}

}  // namespace b200det

using namespace b200det;

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int fill_ptrs(const void *const *src, int n, PtrTab *dst, uintptr_t align_mask) {
    for (int i = 0; i < kMaxLevels; ++i) dst->p[i] = nullptr;
    if (!src) return 0;
    for (int i = 0; i < n; ++i) {
        if (!src[i]) return B200DET_EINVAL;
        if (reinterpret_cast<uintptr_t>(src[i]) & align_mask) return B200DET_EALIGN;
        dst->p[i] = src[i];
    }
    return 0;
}
static int fill_mut_ptrs(void *const *src, int n, MutPtrTab *dst, uintptr_t align_mask) {
    for (int i = 0; i < kMaxLevels; ++i) dst->p[i] = nullptr;
    if (!src) return 0;
    for (int i = 0; i < n; ++i) {
        if (!src[i]) return B200DET_EINVAL;
        if (reinterpret_cast<uintptr_t>(src[i]) & align_mask) return B200DET_EALIGN;
        dst->p[i] = src[i];
    }
    return 0;
}
static void copy_base(const b200det_geometry *geo, BaseAnchors *ba) {
    for (int l = 0; l < kMaxLevels; ++l)
        for (int a = 0; a < kMaxPerLoc; ++a)
            for (int k = 0; k < 4; ++k) ba->v[l][a][k] = geo->base_anchors[l][a][k];
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: `done` holds one bit per
// device ordinal (a process that drives several GPUs raises it on each)
template <typename K>
static int raise_smem_limit(K kernel, std::atomic<unsigned long long> *done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 64 && ((done->load(std::memory_order_relaxed) >> dev) & 1ull)) return 0;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kAssignSmemBudget);
    if (e != cudaSuccess) return (int)e;
    if (dev < 64) done->fetch_or(1ull << dev, std::memory_order_relaxed);
    return 0;
}

extern "C" int b200det_retina_assign(const b200det_geometry *geo, const float *annotations,
                                     int max_gt, float iou_neg, float iou_pos, int32_t *labels,
                                     int32_t *matched, void *workspace, size_t workspace_bytes,
                                     void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!annotations || !labels || !workspace) return B200DET_EINVAL;
    if (max_gt < 1 || max_gt > B200DET_MAX_GT) return B200DET_ERANGE;
    if (!(iou_neg > 0.f) || !(iou_pos >= iou_neg)) return B200DET_EINVAL;
    IouThresholds thr;
    thr.neg = iou_neg;
    thr.pos = iou_pos;
    thr.floor = 0.95f * iou_neg;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    BaseAnchors ba;
    copy_base(geo, &ba);
    TileTab tt = make_tiles(g);
    set_anchor_extents(&tt, g, ba);
    static std::atomic<unsigned long long> a9{0}, a0{0};
    if ((rc = raise_smem_limit(retina_assign_kernel<9>, &a9))) return rc;
    if ((rc = raise_smem_limit(retina_assign_kernel<0>, &a0))) return rc;
    char *base = static_cast<char *>(workspace);
    int *npos = reinterpret_cast<int *>(base + ws.off_assign);
    const Queues q = queues_of(base, ws);
    if (!g_skip_memset) {
        cudaError_t e = cudaMemsetAsync(q.counters, 0, 2 * sizeof(int), (cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
    }
    ProfScope prof(kKernAssign, stream);
    // production scan (labels + queues only): the GT-centric tile kernel
    static const bool no_tiles = getenv("B200DET_ASSIGN_ANCHOR_CENTRIC") != nullptr;   // A/B knob
    if (!matched && !no_tiles) {
        // Small batches: smaller tiles.  A CTA is a chain of short phases (stage + compact the GT rows,
        // queue the (box, shape) items, scatter, labels: ~12 us), so at a few dozen images the kernel is
        // bound by how many CTAs run at once -- 2 per SM with 32 x 32 tiles (74 KB of keys each) -- not
        // by its instruction count (B200DET_TILE_SIDE_SMALL / B200DET_TILE_SMALL_BATCH: A/B knobs)
        static const int side_small = getenv("B200DET_TILE_SIDE_SMALL") ? atoi(getenv("B200DET_TILE_SIDE_SMALL")) : kTileSideSmallBatch;
        static const int small_batch = getenv("B200DET_TILE_SMALL_BATCH") ? atoi(getenv("B200DET_TILE_SMALL_BATCH")) : kTileSmallBatch;
        const int side = (g.batch <= small_batch && side_small >= kTileSideMin && side_small <= kTileSide) ? side_small : kTileSide;
        BigTiles bt = make_big_tiles(g, side);
        for (int l = 0; l < g.n_levels; ++l)
            for (int k = 0; k < 4; ++k) bt.ext[l][k] = tt.ext[l][k];
        const size_t tile_smem = ((gt_smem_bytes(max_gt) + 15) & ~(size_t)15) + (size_t)bt.max_anchors * 8;
        if (tile_smem <= kAssignSmemBudget) {
            static std::atomic<unsigned long long> at{0};
            if ((rc = raise_smem_limit(retina_assign_tile_kernel, &at))) return rc;
            dim3 grid((unsigned)bt.tile_off[g.n_levels], (unsigned)g.batch);
            retina_assign_tile_kernel<<<grid, kTileThreads, tile_smem, (cudaStream_t)stream>>>(
                g, ba, bt, thr, annotations, max_gt, labels, q, npos, (int)ws.assign_blocks_per_image);
            count_launch();
            return (int)cudaGetLastError();
        }
    }
    const size_t smem = assign_dyn_smem(max_gt);
    // g_assign_chunk > 0 (set by the overlapped forward): a few images per launch, so that the
    // kernel's CTAs (96 registers each) displace only part of the HBM-bound sweep they run beside
    const int chunk = g_assign_chunk > 0 ? g_assign_chunk : g.batch;
    for (int b0 = 0; b0 < g.batch; b0 += chunk) {
        dim3 grid((unsigned)ws.assign_blocks_per_image, (unsigned)(g.batch - b0 < chunk ? g.batch - b0 : chunk));
        if (g.per_loc == 9)
            retina_assign_kernel<9><<<grid, kAssignThreads, smem, (cudaStream_t)stream>>>(
                g, ba, tt, thr, annotations, max_gt, labels, matched, q, npos, b0);
        else
            retina_assign_kernel<0><<<grid, kAssignThreads, smem, (cudaStream_t)stream>>>(
                g, ba, tt, thr, annotations, max_gt, labels, matched, q, npos, b0);
        count_launch();
    }
    return (int)cudaGetLastError();
}

extern "C" int b200det_fcos_assign(const b200det_geometry *geo, const float *annotations,
                                   int max_gt, int use_center_sample, int32_t *labels,
                                   int32_t *matched, float *targets, void *workspace,
                                   size_t workspace_bytes, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!annotations || !labels || !workspace) return B200DET_EINVAL;
    if (max_gt < 1 || max_gt > B200DET_MAX_GT) return B200DET_ERANGE;
    if (g.per_loc != 1) return B200DET_EINVAL;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    FcosTab ft;
    for (int l = 0; l < kMaxLevels; ++l) {
        ft.mi_lo[l] = geo->mi_lo[l];
        ft.mi_hi[l] = geo->mi_hi[l];
        ft.radius[l] = geo->radius[l];
    }
    const TileTab tt = make_tiles(g);
    static std::atomic<unsigned long long> done{0};
    if ((rc = raise_smem_limit(fcos_assign_kernel, &done))) return rc;
    char *base = static_cast<char *>(workspace);
    const Queues q = queues_of(base, ws);
    if (!g_skip_memset) {
        cudaError_t e = cudaMemsetAsync(q.counters, 0, 2 * sizeof(int), (cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
    }
    ProfScope prof(kKernAssign, stream);
    const int chunk = g_assign_chunk > 0 ? g_assign_chunk : g.batch;
    for (int b0 = 0; b0 < g.batch; b0 += chunk) {
        dim3 grid((unsigned)ws.assign_blocks_per_image, (unsigned)(g.batch - b0 < chunk ? g.batch - b0 : chunk));
        fcos_assign_kernel<<<grid, kAssignThreads, assign_dyn_smem(max_gt), (cudaStream_t)stream>>>(
            g, ft, tt, annotations, max_gt, use_center_sample, labels, matched, targets, q,
            reinterpret_cast<int *>(base + ws.off_assign), b0);
        count_launch();
    }
    return (int)cudaGetLastError();
}

extern "C" int b200det_sparse_losses(const b200det_geometry *geo, int is_fcos,
                                     const float *annotations, int max_gt, const int32_t *labels,
                                     const void *const *reg, int reg_dtype,
                                     const void *const *ctr, int box_loss, float beta,
                                     const void *const *cls, float alpha, float gamma,
                                     void *const *reg_grad, void *const *ctr_grad,
                                     void *workspace, size_t workspace_bytes, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!annotations || !labels || !workspace) return B200DET_EINVAL;
    if (max_gt < 1 || max_gt > B200DET_MAX_GT) return B200DET_ERANGE;
    if (box_loss < B200DET_BOX_NONE || box_loss > B200DET_BOX_EIOU) return B200DET_EINVAL;
    const bool ctr_logits = (is_fcos & B200DET_FCOS_CTR_LOGITS) != 0;
    is_fcos &= 1;
    if (ctr_logits && (!is_fcos || ctr_grad)) return B200DET_EINVAL;
    if (is_fcos && (box_loss == B200DET_BOX_SMOOTHL1 || g.per_loc != 1)) return B200DET_EINVAL;
    const bool with_loss = box_loss != B200DET_BOX_NONE;
    if (with_loss && (!reg || (is_fcos && !ctr))) return B200DET_EINVAL;
    const int reg_base = reg_dtype & 0xf;
    if ((reg_dtype & ~(0xf | B200DET_REG_EXP_ROUNDED)) ||
        (reg_base != B200DET_F32 && reg_base != B200DET_F16 && reg_base != B200DET_BF16))
        return B200DET_EINVAL;
    SparseArgs a;
    a.g = g;
    copy_base(geo, &a.ba);
    if ((rc = fill_ptrs(with_loss ? reg : nullptr, g.n_levels, &a.reg,
                        reg_base == B200DET_F32 ? 15 : 7)))
        return rc;
    if ((rc = fill_ptrs(with_loss && is_fcos ? ctr : nullptr, g.n_levels, &a.ctr, 3))) return rc;
    if ((rc = fill_ptrs(cls, g.n_levels, &a.cls, 3))) return rc;
    if ((rc = fill_mut_ptrs(with_loss ? reg_grad : nullptr, g.n_levels, &a.reg_grad, 15))) return rc;
    if ((rc = fill_mut_ptrs(with_loss && is_fcos ? ctr_grad : nullptr, g.n_levels, &a.ctr_grad, 3)))
        return rc;
    a.reg_dtype = reg_dtype;
    a.box_loss = box_loss;
    a.is_fcos = is_fcos;
    a.ctr_logits = ctr_logits ? 1 : 0;
    a.G = max_gt;
    a.C = g.num_classes;
    a.beta = beta;
    a.alpha = alpha;
    a.gamma = gamma;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    char *base = static_cast<char *>(workspace);
    ProfScope prof(kKernSparse, stream);
    sparse_loss_kernel<<<(unsigned)ws.sparse_blocks, kSparseThreads, 0, (cudaStream_t)stream>>>(
        a, annotations, labels, queues_of(base, ws),
        reinterpret_cast<SparsePartial *>(base + ws.off_sparse));
    count_launch();
    return (int)cudaGetLastError();
}

extern "C" int b200det_scale_pos_rows(const b200det_geometry *geo, const void *workspace,
                                      size_t workspace_bytes, void *const *reg_grad,
                                      void *const *ctr_grad, const float *g_box, const float *g_ctr,
                                      const double *sums, float w_box, float w_ctr, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!workspace || !sums || (!g_box && !g_ctr)) return B200DET_EINVAL;
    if ((g_box && !reg_grad) || (g_ctr && !ctr_grad)) return B200DET_EINVAL;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    ScaleRowsArgs a;
    a.g = g;
    if ((rc = fill_mut_ptrs(g_box ? reg_grad : nullptr, g.n_levels, &a.reg_grad, 15))) return rc;
    if ((rc = fill_mut_ptrs(g_ctr ? ctr_grad : nullptr, g.n_levels, &a.ctr_grad, 3))) return rc;
    a.g_box = g_box;
    a.g_ctr = g_ctr;
    a.sums = sums;
    a.w_box = w_box;
    a.w_ctr = w_ctr;
    char *base = const_cast<char *>(static_cast<const char *>(workspace));
    ProfScope prof(kKernOther, stream);
    scale_pos_rows_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(a, queues_of(base, ws));
    count_launch();
    return (int)cudaGetLastError();
}

extern "C" int b200det_generate_rows(const b200det_geometry *geo, int is_fcos, float *out,
                                     void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!out) return B200DET_EINVAL;
    BaseAnchors ba;
    copy_base(geo, &ba);
    const int N = g.off[g.n_levels];
    generate_rows_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, ba, is_fcos, out);
    count_launch();
    return (int)cudaGetLastError();
}

extern "C" int b200det_rows_to_image_major(const b200det_geometry *geo, const void *src,
                                           void *dst, int width, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!src || !dst || width < 1) return B200DET_EINVAL;
    const long long total = (long long)g.batch * g.off[g.n_levels];
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    rows_to_image_major_kernel<<<blocks > 0 ? blocks : 1, 256, 0, (cudaStream_t)stream>>>(
        g, static_cast<const uint32_t *>(src), static_cast<uint32_t *>(dst), width);
    count_launch();
    return (int)cudaGetLastError();
}
