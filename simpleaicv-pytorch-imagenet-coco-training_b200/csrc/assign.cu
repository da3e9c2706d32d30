// assign.cu -- target assignment + box / centre-ness losses (prediction-independent part of
// RetinaLoss / FCOSLoss).  Compiled with -fmad=false: every float op below is one IEEE
// float32 operation in the reference's order, so labels and matched
// indices are bit-exact.
//
// Layout / roofline: per image the kernels read G<=2048 annotation rows (20 B each, staged
// once per CTA into shared memory with one cp.async.bulk + mbarrier), generate anchors / points
// in registers, and write 4 (labels) [+4 matched, +24 targets] bytes per row; regression rows
// are read only for positives.  The kernels are ALU/latency work that must hide under the
// classification sweep (focal.cu); Retina culls GT boxes that cannot overlap the CTA's 256
// anchors before the pair loop, and skips the IEEE divide when the overlap is empty.
#include "common.cuh"
#include "dual.cuh"

namespace b200det {

constexpr int kAssignThreads = 256;

// ---------------------------------------------------------------------------------------
// GT staging: global [G,5] float rows -> shared memory, via TMA bulk copy when aligned.
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

// Shared-memory carve-up (dynamic): raw rows, then the compacted candidate arrays.
struct GtSmem {
    float *raw;    // [G*5]
    float4 *box;   // [G]
    float *area;   // [G]
    int *label;    // [G]  class + 1
    int *fidx;     // [G]  index in the filtered (class >= 0) list
};
__device__ __forceinline__ GtSmem carve(unsigned char *base, int G) {
    GtSmem s;
    const int raw_floats = (G * 5 + 3) & ~3;
    s.raw = reinterpret_cast<float *>(base);
    s.box = reinterpret_cast<float4 *>(s.raw + raw_floats);
    s.area = reinterpret_cast<float *>(s.box + G);
    s.label = reinterpret_cast<int *>(s.area + G);
    s.fidx = s.label + G;
    return s;
}
static size_t gt_smem_bytes(int G) {
    const int raw_floats = (G * 5 + 3) & ~3;
    return (size_t)raw_floats * 4 + (size_t)G * (16 + 4 + 4 + 4);
}

// Ordered compaction of annotation rows: keeps rows with class >= 0 (filtered index = rank
// among them, losses.py:338-339 / :667-668) and, if `cull`, only those whose box can overlap
// the CTA's region [rx1,ry1,rx2,ry2].  Returns {#kept, #valid}.  Order is preserved, which
// is what makes "first maximum / first minimum" tie rules exact.
__device__ __forceinline__ int2 compact_gt(const GtSmem &s, int G, bool cull, float rx1,
                                           float ry1, float rx2, float ry2, bool fcos_area,
                                           int *warp_cnt /* [2*8] smem */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = kAssignThreads / 32;
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
        if (cull && valid)
            keep = (fminf(rx2, x2) > fmaxf(rx1, x1)) && (fminf(ry2, y2) > fmaxf(ry1, y1));
        const unsigned bv = __ballot_sync(0xffffffffu, valid);
        const unsigned bk = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) {
            warp_cnt[warp] = __popc(bv);
            warp_cnt[nwarp + warp] = __popc(bk);
        }
        __syncthreads();
        int pv = base_valid, pk = base_keep, tv = 0, tk = 0;
#pragma unroll
        for (int w = 0; w < nwarp; ++w) {
            const int cv = warp_cnt[w], ck = warp_cnt[nwarp + w];
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
            s.fidx[pk] = pv;
        }
        base_valid += tv;
        base_keep += tk;
        __syncthreads();
    }
    return make_int2(base_keep, base_valid);
}

__device__ __forceinline__ void block_partial(int npos, float box, float ctr,
                                              AssignPartial *dst, float *red /* [3*8] smem */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = kAssignThreads / 32;
    const int wn = warp_sum_int(npos);
    const float wb = warp_sum(box), wc = warp_sum(ctr);
    if (lane == 0) {
        red[warp] = __int_as_float(wn);
        red[nwarp + warp] = wb;
        red[2 * nwarp + warp] = wc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        float b = 0.f, c = 0.f;
        for (int w = 0; w < nwarp; ++w) {
            n += __float_as_int(red[w]);
            b += red[nwarp + w];
            c += red[2 * nwarp + w];
        }
        AssignPartial p;
        p.npos = n;
        p.box = b;
        p.ctr = c;
        p.pad = 0.f;
        *dst = p;
    }
}

// ---------------------------------------------------------------------------------------
// Retina: anchor <-> GT IoU, max / arg-max, label, box loss
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAssignThreads)
    retina_assign_kernel(Geo g, BaseAnchors ba, const float *__restrict__ annots, int G,
                         PtrTab reg, int reg_dtype, int box_loss, float beta,
                         int *__restrict__ labels, int *__restrict__ matched, MutPtrTab reg_grad,
                         AssignPartial *__restrict__ partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ float red[4 * (kAssignThreads / 32)];
    __shared__ int warp_cnt[2 * (kAssignThreads / 32)];
    __shared__ float region[4];

    const int b = blockIdx.y;
    const int N = g.off[g.n_levels];
    const int row = blockIdx.x * kAssignThreads + threadIdx.x;
    const bool active = row < N;
    const GtSmem s = carve(smem_raw, G);

    const float *src = annots + (size_t)b * G * 5;
    const bool bulk = (((G * 5 * 4) & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    stage_rows_begin(s.raw, src, G * 5, &mbar, bulk);

    // this thread's anchor, generated in registers (models/anchor.py:59-86)
    int l = 0, local = 0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
        l = level_of_row(g, row);
        local = row - g.off[l];
        a = anchor_of(g, ba, l, local);
    }
    // CTA region = bounding box of its anchors (for GT culling)
    {
        const float big = 3.0e38f;
        float mnx = active ? a.x : big, mny = active ? a.y : big;
        float mxx = active ? a.z : -big, mxy = active ? a.w : -big;
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
            for (int w = 1; w < kAssignThreads / 32; ++w) {
                mnx = fminf(mnx, red[w * 4 + 0]);
                mny = fminf(mny, red[w * 4 + 1]);
                mxx = fmaxf(mxx, red[w * 4 + 2]);
                mxy = fmaxf(mxy, red[w * 4 + 3]);
            }
            region[0] = mnx;
            region[1] = mny;
            region[2] = mxx;
            region[3] = mxy;
        }
    }
    stage_rows_wait(&mbar, bulk);  // contains a __syncthreads(): region[] is visible too

    const int2 cnt = compact_gt(s, G, true, region[0], region[1], region[2], region[3], false,
                                warp_cnt);
    const int n_cand = cnt.x;
    const bool has_gt = cnt.y > 0;

    // IoU scan (losses.py:54-70, :357): strict '>' in GT order == first maximum.
    // Culled / non-overlapping GTs have IoU exactly 0 and can never beat best >= 0.
    const float aw = fmaxf(__fsub_rn(a.z, a.x), 0.f), ah = fmaxf(__fsub_rn(a.w, a.y), 0.f);
    const float area_a = __fmul_rn(aw, ah);
    float best = 0.f;
    int best_slot = -1;
    for (int k = 0; k < n_cand; ++k) {
        const float4 gt = s.box[k];
        const float mnx = fminf(a.z, gt.z), mxx = fmaxf(a.x, gt.x);
        const float mny = fminf(a.w, gt.w), mxy = fmaxf(a.y, gt.y);
        if (mnx > mxx && mny > mxy) {
            const float ov = __fmul_rn(__fsub_rn(mnx, mxx), __fsub_rn(mny, mxy));
            const float un = fmaxf(__fsub_rn(__fadd_rn(area_a, s.area[k]), ov), 1e-4f);
            const float iou = __fdiv_rn(ov, un);
            if (iou > best) {
                best = iou;
                best_slot = k;
            }
        }
    }

    int label = -1, match = -1;
    if (has_gt) {
        match = best_slot >= 0 ? s.fidx[best_slot] : 0;
        if (best < 0.4f) label = 0;
        if (best >= 0.5f) label = s.label[best_slot];
    }
    long long lm = 0;
    if (active) {
        lm = lm_index(g, b, l, local);
        labels[lm] = label;
        if (matched) matched[lm] = match;
    }

    // box loss for positives (losses.py:263-320)
    float box_term = 0.f;
    const bool pos = active && label > 0;
    if (box_loss != B200DET_BOX_NONE && active) {
        const long long rrow = (long long)b * g.rows[l] + local;
        float4 grad = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pos) {
            const float4 t = load_reg4(reg.p[l], reg_dtype, rrow);
            const float4 gt = s.box[best_slot];
            const float awx = __fsub_rn(a.z, a.x), awy = __fsub_rn(a.w, a.y);
            const float acx = __fadd_rn(a.x, __fmul_rn(0.5f, awx));
            const float acy = __fadd_rn(a.y, __fmul_rn(0.5f, awy));
            if (box_loss == B200DET_BOX_SMOOTHL1) {
                // targets (losses.py:390-409)
                const float gwx = fmaxf(__fsub_rn(gt.z, gt.x), 1e-4f);
                const float gwy = fmaxf(__fsub_rn(gt.w, gt.y), 1e-4f);
                const float gcx = __fadd_rn(gt.x, __fmul_rn(0.5f, gwx));
                const float gcy = __fadd_rn(gt.y, __fmul_rn(0.5f, gwy));
                const float tg[4] = {__fdiv_rn(__fsub_rn(gcx, acx), awx),
                                     __fdiv_rn(__fsub_rn(gcy, acy), awy),
                                     logf(__fdiv_rn(gwx, awx)), logf(__fdiv_rn(gwy, awy))};
                const float pr[4] = {t.x, t.y, t.z, t.w};
                float gr[4];
                const float half_beta = 0.5f * beta;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float d = __fsub_rn(pr[i], tg[i]);
                    const float x = fabsf(d);
                    if (x >= beta) {
                        box_term += __fsub_rn(x, half_beta);
                        gr[i] = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
                    } else {
                        box_term += __fdiv_rn(__fmul_rn(0.5f, __fmul_rn(x, x)), beta);
                        gr[i] = d / beta;
                    }
                }
                grad = make_float4(gr[0], gr[1], gr[2], gr[3]);
            } else {
                // decode (losses.py:411-429) then 1 - IoU-family (losses.py:286-293)
                const Dual tx = dvar(t.x, 0), ty = dvar(t.y, 1), tw = dvar(t.z, 2), th = dvar(t.w, 3);
                const Dual bw = dexp(tw) * awx, bh = dexp(th) * awy;
                const Dual cx = tx * awx + acx, cy = ty * awy + acy;
                const Dual hw = bw * 0.5f, hh = bh * 0.5f;
                const Dual p[4] = {cx - hw, cy - hh, cx + hw, cy + hh};
                const float gg[4] = {gt.x, gt.y, gt.z, gt.w};
                const Dual iou = iou_family(p, gg, box_loss);
                box_term = __fsub_rn(1.f, iou.v);
                grad = make_float4(-iou.d[0], -iou.d[1], -iou.d[2], -iou.d[3]);
            }
        }
        if (reg_grad.p[0] != nullptr)
            reinterpret_cast<float4 *>(reg_grad.p[l])[rrow] = grad;
    }
    block_partial(pos ? 1 : 0, box_term, 0.f,
                  partials + (size_t)blockIdx.y * gridDim.x + blockIdx.x, red);
}

// ---------------------------------------------------------------------------------------
// FCOS: point <-> GT with centre sampling, scale range, min-area; IoU + centre-ness losses
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAssignThreads)
    fcos_assign_kernel(Geo g, FcosTab ft, const float *__restrict__ annots, int G, PtrTab reg,
                       int reg_dtype, PtrTab ctr, int box_loss, int use_center_sample,
                       int *__restrict__ labels, int *__restrict__ matched,
                       float *__restrict__ targets, MutPtrTab reg_grad, MutPtrTab ctr_grad,
                       AssignPartial *__restrict__ partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ float red[4 * (kAssignThreads / 32)];
    __shared__ int warp_cnt[2 * (kAssignThreads / 32)];

    const int b = blockIdx.y;
    const int N = g.off[g.n_levels];
    const int row = blockIdx.x * kAssignThreads + threadIdx.x;
    const bool active = row < N;
    const GtSmem s = carve(smem_raw, G);

    const float *src = annots + (size_t)b * G * 5;
    const bool bulk = (((G * 5 * 4) & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    stage_rows_begin(s.raw, src, G * 5, &mbar, bulk);

    int l = 0, local = 0;
    float2 pt = make_float2(0.f, 0.f);
    if (active) {
        l = level_of_row(g, row);
        local = row - g.off[l];
        pt = point_of(g, l, local);
    }
    const float m0 = ft.mi_lo[l], m1 = ft.mi_hi[l], rad = ft.radius[l];
    stage_rows_wait(&mbar, bulk);
    const int2 cnt = compact_gt(s, G, false, 0.f, 0.f, 0.f, 0.f, true, warp_cnt);
    const int n_gt = cnt.x;

    // losses.py:688-735 (candidate tests) and :785-808 (smallest area, first minimum)
    float best_area = __int_as_float(0x7f800000);
    int best = -1;
    float bl = 0.f, bt = 0.f, br = 0.f, bb = 0.f;
    for (int k = 0; k < n_gt; ++k) {
        const float4 gt = s.box[k];
        const float cl = __fsub_rn(pt.x, gt.x), ct = __fsub_rn(pt.y, gt.y);
        const float cr = __fsub_rn(gt.z, pt.x), cb = __fsub_rn(gt.w, pt.y);
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
        if (ok && s.area[k] < best_area) {
            best_area = s.area[k];
            best = k;
            bl = cl;
            bt = ct;
            br = cr;
            bb = cb;
        }
    }
    const bool pos = active && best >= 0;
    int label = 0;
    float ctr_t = 0.f;
    if (pos) {
        label = s.label[best];
        // losses.py:822-824
        ctr_t = __fsqrt_rn(__fmul_rn(__fdiv_rn(fminf(bl, br), fmaxf(bl, br)),
                                     __fdiv_rn(fminf(bt, bb), fmaxf(bt, bb))));
    }
    if (active) {
        const long long lm = lm_index(g, b, l, local);
        labels[lm] = label;
        if (matched) matched[lm] = pos ? s.fidx[best] : -1;
        if (targets) {
            float *t = targets + lm * 6;
            t[0] = bl;
            t[1] = bt;
            t[2] = br;
            t[3] = bb;
            t[4] = (float)label;
            t[5] = ctr_t;
        }
    }

    float box_term = 0.f, ctr_term = 0.f;
    if (box_loss != B200DET_BOX_NONE && active) {
        const long long rrow = (long long)b * g.rows[l] + local;
        float4 grad = make_float4(0.f, 0.f, 0.f, 0.f);
        float cgrad = 0.f;
        if (pos) {
            // IoU loss, losses.py:550-586: boxes rebuilt around the point from l,t,r,b
            const float4 t = load_reg4(reg.p[l], reg_dtype, rrow);
            const Dual e0 = dexp(dvar(t.x, 0)), e1 = dexp(dvar(t.y, 1));
            const Dual e2 = dexp(dvar(t.z, 2)), e3 = dexp(dvar(t.w, 3));
            const Dual p[4] = {pt.x - e0, pt.y - e1, e2 + pt.x, e3 + pt.y};
            const float gg[4] = {__fsub_rn(pt.x, bl), __fsub_rn(pt.y, bt), __fadd_rn(pt.x, br),
                                 __fadd_rn(pt.y, bb)};
            const Dual iou = iou_family(p, gg, box_loss);
            box_term = __fmul_rn(__fsub_rn(1.f, iou.v), ctr_t);
            grad = make_float4(-iou.d[0] * ctr_t, -iou.d[1] * ctr_t, -iou.d[2] * ctr_t,
                               -iou.d[3] * ctr_t);
            // centre-ness BCE, losses.py:588-610 (prob clamped at losses.py:494)
            const float craw = __ldg(reinterpret_cast<const float *>(ctr.p[l]) + rrow);
            const float lo = 1e-4f, hi = 0.9999f;  // float32(1e-4), float32(1. - 1e-4)
            const float cp = fminf(fmaxf(craw, lo), hi);
            const float one_m = __fsub_rn(1.f, cp), one_t = __fsub_rn(1.f, ctr_t);
            ctr_term = -__fadd_rn(__fmul_rn(ctr_t, logf(cp)), __fmul_rn(one_t, logf(one_m)));
            if (craw >= lo && craw <= hi) cgrad = -(ctr_t / cp - one_t / one_m);
        }
        if (reg_grad.p[0] != nullptr)
            reinterpret_cast<float4 *>(reg_grad.p[l])[rrow] = grad;
        if (ctr_grad.p[0] != nullptr) reinterpret_cast<float *>(ctr_grad.p[l])[rrow] = cgrad;
    }
    block_partial(pos ? 1 : 0, box_term, ctr_term,
                  partials + (size_t)blockIdx.y * gridDim.x + blockIdx.x, red);
}

// ---------------------------------------------------------------------------------------
// utilities: materialise rows (tests), level-major -> image-major
// ---------------------------------------------------------------------------------------
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

}  // namespace b200det

using namespace b200det;

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int fill_ptrs(const void *const *src, int n, PtrTab *dst) {
    for (int i = 0; i < kMaxLevels; ++i) dst->p[i] = nullptr;
    if (!src) return 0;
    for (int i = 0; i < n; ++i) {
        if (!src[i]) return B200DET_EINVAL;
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
static int check_reg(const PtrTab &t, int n, int dtype) {
    if (dtype != B200DET_F32 && dtype != B200DET_F16 && dtype != B200DET_BF16)
        return B200DET_EINVAL;
    const uintptr_t mask = dtype == B200DET_F32 ? 15 : 7;
    for (int i = 0; i < n; ++i)
        if (reinterpret_cast<uintptr_t>(t.p[i]) & mask) return B200DET_EALIGN;
    return 0;
}

extern "C" int b200det_retina_assign(const b200det_geometry *geo, const float *annotations,
                                     int max_gt, const void *const *reg, int reg_dtype,
                                     int box_loss, float beta, int32_t *labels,
                                     int32_t *matched, void *const *reg_grad, void *workspace,
                                     size_t workspace_bytes, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!annotations || !labels || !workspace) return B200DET_EINVAL;
    if (max_gt < 1 || max_gt > B200DET_MAX_GT) return B200DET_ERANGE;
    if (box_loss < B200DET_BOX_NONE || box_loss > B200DET_BOX_EIOU) return B200DET_EINVAL;
    if (box_loss != B200DET_BOX_NONE && !reg) return B200DET_EINVAL;
    if (g.per_loc > kMaxPerLoc) return B200DET_ERANGE;
    PtrTab regt;
    MutPtrTab gradt;
    if ((rc = fill_ptrs(box_loss != B200DET_BOX_NONE ? reg : nullptr, g.n_levels, &regt))) return rc;
    if (box_loss != B200DET_BOX_NONE && (rc = check_reg(regt, g.n_levels, reg_dtype))) return rc;
    if ((rc = fill_mut_ptrs(box_loss != B200DET_BOX_NONE ? reg_grad : nullptr, g.n_levels, &gradt, 15)))
        return rc;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    BaseAnchors ba;
    for (int l = 0; l < kMaxLevels; ++l)
        for (int a = 0; a < kMaxPerLoc; ++a)
            for (int k = 0; k < 4; ++k) ba.v[l][a][k] = geo->base_anchors[l][a][k];

    const size_t smem = gt_smem_bytes(max_gt);
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        cudaError_t e = cudaFuncSetAttribute(retina_assign_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)gt_smem_bytes(B200DET_MAX_GT));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((unsigned)ws.assign_blocks_per_image, (unsigned)g.batch);
    retina_assign_kernel<<<grid, kAssignThreads, smem, (cudaStream_t)stream>>>(
        g, ba, annotations, max_gt, regt, reg_dtype, box_loss, beta, labels, matched, gradt,
        reinterpret_cast<AssignPartial *>(static_cast<char *>(workspace) + ws.off_assign));
    count_launch();
    return (int)cudaGetLastError();
}

extern "C" int b200det_fcos_assign(const b200det_geometry *geo, const float *annotations,
                                   int max_gt, const void *const *reg, int reg_dtype,
                                   const void *const *ctr, int box_loss, int use_center_sample,
                                   int32_t *labels, int32_t *matched, float *targets,
                                   void *const *reg_grad, void *const *ctr_grad, void *workspace,
                                   size_t workspace_bytes, void *stream) {
    Geo g;
    int rc = make_geo(geo, &g);
    if (rc) return rc;
    if (!annotations || !labels || !workspace) return B200DET_EINVAL;
    if (max_gt < 1 || max_gt > B200DET_MAX_GT) return B200DET_ERANGE;
    if (box_loss == B200DET_BOX_SMOOTHL1 || box_loss < B200DET_BOX_NONE || box_loss > B200DET_BOX_EIOU)
        return B200DET_EINVAL;
    if (g.per_loc != 1) return B200DET_EINVAL;
    const bool with_loss = box_loss != B200DET_BOX_NONE;
    if (with_loss && (!reg || !ctr)) return B200DET_EINVAL;
    PtrTab regt, ctrt;
    MutPtrTab rgrad, cgrad;
    if ((rc = fill_ptrs(with_loss ? reg : nullptr, g.n_levels, &regt))) return rc;
    if ((rc = fill_ptrs(with_loss ? ctr : nullptr, g.n_levels, &ctrt))) return rc;
    if (with_loss && (rc = check_reg(regt, g.n_levels, reg_dtype))) return rc;
    if ((rc = fill_mut_ptrs(with_loss ? reg_grad : nullptr, g.n_levels, &rgrad, 15))) return rc;
    if ((rc = fill_mut_ptrs(with_loss ? ctr_grad : nullptr, g.n_levels, &cgrad, 3))) return rc;
    const LossWs ws = loss_ws_layout(g);
    if (workspace_bytes < ws.total) return B200DET_EWORKSPACE;
    FcosTab ft;
    for (int l = 0; l < kMaxLevels; ++l) {
        ft.mi_lo[l] = geo->mi_lo[l];
        ft.mi_hi[l] = geo->mi_hi[l];
        ft.radius[l] = geo->radius[l];
    }
    const size_t smem = gt_smem_bytes(max_gt);
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        cudaError_t e = cudaFuncSetAttribute(fcos_assign_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)gt_smem_bytes(B200DET_MAX_GT));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((unsigned)ws.assign_blocks_per_image, (unsigned)g.batch);
    fcos_assign_kernel<<<grid, kAssignThreads, smem, (cudaStream_t)stream>>>(
        g, ft, annotations, max_gt, regt, reg_dtype, ctrt, box_loss, use_center_sample, labels,
        matched, targets, rgrad, cgrad,
        reinterpret_cast<AssignPartial *>(static_cast<char *>(workspace) + ws.off_assign));
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
    for (int l = 0; l < kMaxLevels; ++l)
        for (int a = 0; a < kMaxPerLoc; ++a)
            for (int k = 0; k < 4; ++k) ba.v[l][a][k] = geo->base_anchors[l][a][k];
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
