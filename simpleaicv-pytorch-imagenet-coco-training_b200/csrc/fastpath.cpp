// fastpath.cpp -- OPTIONAL host-side accelerator of the Python layer.  Not a compute path: it only
// does, in C++, the argument checking / pointer marshalling that losses.py and decode.py do in
// Python before they call the C ABI, for the COMMON case (contiguous, aligned CUDA tensors of the
// expected dtypes on one device), and then calls the very same C-ABI entry points
// (b200det_loss_forward_overlap, b200det_decode) through function pointers handed over by the
// ctypes binding.  Anything else returns None and the Python path (with its conversions and error
// messages) takes over.  Why it exists: the reference's eval loop is two synchronous calls per step
// (tools/scripts.py:733-740); the host time between the decoder's sync and the criterion's first
// launch is GPU idle time -- ~35 us per step in Python, 7 % of a 32-image shard's step.
#include <torch/extension.h>

#include <cstdint>
#include <vector>

#include "../../include/b200det.h"

namespace {

using LossFn = int (*)(const b200det_geometry *, const b200det_loss_params *, const float *, int,
                       const void *const *, const void *const *, const void *const *, int32_t *,
                       void *, size_t, const b200det_peer_exchange *, double *, float *, int32_t *,
                       void *, void *, void *, void *, int);
using DecodeFn = int (*)(const b200det_geometry *, const b200det_decode_params *,
                         const void *const *, const void *const *, const void *const *, uint32_t *,
                         int32_t *, float *, int32_t *, int32_t *, int32_t *, void *, size_t, void *);

// b200det_loss_forward_keys / b200det_decode_from_keys: the sweep hand-over (b200det/_handoff.py)
using LossKeysFn = int (*)(const b200det_geometry *, const b200det_loss_params *, const float *, int,
                           const void *const *, const void *const *, const void *const *, int32_t *,
                           void *, size_t, const b200det_peer_exchange *, double *, float *, int32_t *,
                           void *, void *, void *, void *, int, float, uint32_t *, int32_t *);
using FromKeysFn = int (*)(const b200det_geometry *, const b200det_decode_params *,
                           const void *const *, const void *const *, const void *const *,
                           const uint32_t *, const int32_t *, float *, int32_t *, int32_t *, int32_t *,
                           int32_t *, int, void *);

LossFn g_loss = nullptr;
DecodeFn g_decode = nullptr;
LossKeysFn g_loss_keys = nullptr;
FromKeysFn g_from_keys = nullptr;

void bind(uintptr_t loss_forward_overlap, uintptr_t decode, uintptr_t loss_forward_keys,
          uintptr_t decode_from_keys) {
    g_loss = reinterpret_cast<LossFn>(loss_forward_overlap);
    g_decode = reinterpret_cast<DecodeFn>(decode);
    g_loss_keys = reinterpret_cast<LossKeysFn>(loss_forward_keys);
    g_from_keys = reinterpret_cast<FromKeysFn>(decode_from_keys);
}

// per-level tensors -> device pointers; false when a tensor is not the common case
bool fill(const py::list &levels, int n_levels, c10::ScalarType dtype, int device, uintptr_t align_mask,
          const long long *numel, const void **out) {
    if ((int)levels.size() != n_levels) return false;
    int i = 0;
    for (py::handle h : levels) {
        if (!THPVariable_Check(h.ptr())) return false;
        const at::Tensor &t = THPVariable_Unpack(h.ptr());
        if (!t.is_cuda() || t.get_device() != device || t.scalar_type() != dtype ||
            !t.is_contiguous() || (numel && t.numel() != numel[i]))
            return false;
        const uintptr_t p = reinterpret_cast<uintptr_t>(t.data_ptr());
        if (p & align_mask) return false;
        out[i++] = reinterpret_cast<const void *>(p);
    }
    return true;
}

int reg_code(c10::ScalarType t) {
    if (t == at::kFloat) return B200DET_F32;
    if (t == at::kHalf) return B200DET_F16;
    if (t == at::kBFloat16) return B200DET_BF16;
    return -1;
}

struct Levels {
    const void *cls[B200DET_MAX_LEVELS], *reg[B200DET_MAX_LEVELS], *ctr[B200DET_MAX_LEVELS];
    int reg_dtype = -1;
    bool has_ctr = false;
};

// checks and marshals [cls, reg(, ctr)] against the plan's geometry
bool marshal(const b200det_geometry *geo, const py::list &cls, const py::list &reg,
             const py::object &ctr, int device, Levels *lv) {
    const int n = geo->n_levels;
    long long n_cls[B200DET_MAX_LEVELS], n_reg[B200DET_MAX_LEVELS], n_ctr[B200DET_MAX_LEVELS];
    for (int l = 0; l < n; ++l) {
        const long long rows = (long long)geo->batch * geo->height[l] * geo->width[l] * geo->per_loc;
        n_cls[l] = rows * geo->num_classes;
        n_reg[l] = rows * 4;
        n_ctr[l] = rows;
    }
    if (!fill(cls, n, at::kFloat, device, 15, n_cls, lv->cls)) return false;
    if (reg.size() == 0 || !THPVariable_Check(reg[0].ptr())) return false;
    const c10::ScalarType rt = THPVariable_Unpack(reg[0].ptr()).scalar_type();
    lv->reg_dtype = reg_code(rt);
    if (lv->reg_dtype < 0) return false;
    if (!fill(reg, n, rt, device, lv->reg_dtype == B200DET_F32 ? 15 : 7, n_reg, lv->reg)) return false;
    lv->has_ctr = !ctr.is_none();
    if (lv->has_ctr && !fill(py::cast<py::list>(ctr), n, at::kFloat, device, 3, n_ctr, lv->ctr))
        return false;
    return true;
}

// RetinaLoss / FCOSLoss no-grad forward.  p = (is_fcos, box_loss, use_center_sample, alpha, gamma,
// beta, w_cls, w_box, w_ctr, iou_neg, iou_pos).  sync: 0 = single process, 1 = the caller
// all-reduces the sums (no finish here), 2 = peer exchange (px).  keys != 0: the sweep also writes the
// decoder's keys / classes (thresholded with min_score).  Returns the 8-double result tensor
// (sums | losses | status word) or None.
py::object loss_eval(uintptr_t geo_addr, const py::list &cls, const py::list &reg, const py::object &ctr,
                     const py::object &ann_obj, const py::tuple &p, bool autocast,
                     uintptr_t scratch, size_t ws_bytes, int sync, uintptr_t px, uintptr_t side,
                     uintptr_t fork, uintptr_t join, uintptr_t stream, float min_score, uintptr_t keys,
                     uintptr_t classes) {
    if (!g_loss || !g_loss_keys) return py::none();
    const b200det_geometry *geo = reinterpret_cast<const b200det_geometry *>(geo_addr);
    if (cls.size() == 0 || !THPVariable_Check(cls[0].ptr()) || !THPVariable_Check(ann_obj.ptr()))
        return py::none();
    const at::Tensor &first = THPVariable_Unpack(cls[0].ptr());
    if (!first.is_cuda()) return py::none();
    const int device = first.get_device();
    Levels lv;
    if (!marshal(geo, cls, reg, ctr, device, &lv)) return py::none();
    const at::Tensor &ann = THPVariable_Unpack(ann_obj.ptr());
    if (!ann.is_cuda() || ann.get_device() != device || ann.scalar_type() != at::kFloat ||
        !ann.is_contiguous() || ann.dim() != 3 || ann.size(2) != 5 || ann.size(0) != geo->batch ||
        ann.size(1) < 1 || ann.size(1) > B200DET_MAX_GT)
        return py::none();
    b200det_loss_params lp;
    lp.is_fcos = p[0].cast<int>();
    lp.box_loss = p[1].cast<int>();
    lp.use_center_sample = p[2].cast<int>();
    lp.alpha = p[3].cast<float>();
    lp.gamma = p[4].cast<float>();
    lp.beta = p[5].cast<float>();
    lp.w_cls = p[6].cast<float>();
    lp.w_box = p[7].cast<float>();
    lp.w_ctr = p[8].cast<float>();
    lp.iou_neg = p[9].cast<float>();
    lp.iou_pos = p[10].cast<float>();
    if ((lp.is_fcos != 0) != lv.has_ctr) return py::none();
    // eager half arithmetic rounds exp() to half; under autocast exp runs in float32 (losses.py)
    lp.reg_dtype = lv.reg_dtype;
    if (lv.reg_dtype != B200DET_F32 && !autocast) lp.reg_dtype |= B200DET_REG_EXP_ROUNDED;
    at::Tensor out = at::empty({8}, first.options().dtype(at::kDouble));
    double *sums = out.data_ptr<double>();
    float *losses = sync == 1 ? nullptr : reinterpret_cast<float *>(sums + 4);
    int32_t *status = sync == 2 ? reinterpret_cast<int32_t *>(sums + 6) : nullptr;
    char *ws = reinterpret_cast<char *>(scratch);
    const b200det_peer_exchange *pxp =
        sync == 2 ? reinterpret_cast<const b200det_peer_exchange *>(px) : nullptr;
    int32_t *labels = reinterpret_cast<int32_t *>(ws + ws_bytes);
    const void *const *ctrp = lv.has_ctr ? lv.ctr : nullptr;
    const int rc =
        keys ? g_loss_keys(geo, &lp, ann.data_ptr<float>(), (int)ann.size(1), lv.cls, lv.reg, ctrp, labels,
                           ws, ws_bytes, pxp, sums, losses, status, reinterpret_cast<void *>(side),
                           reinterpret_cast<void *>(fork), reinterpret_cast<void *>(join),
                           reinterpret_cast<void *>(stream), 0, min_score,
                           reinterpret_cast<uint32_t *>(keys), reinterpret_cast<int32_t *>(classes))
             : g_loss(geo, &lp, ann.data_ptr<float>(), (int)ann.size(1), lv.cls, lv.reg, ctrp, labels, ws,
                      ws_bytes, pxp, sums, losses, status, reinterpret_cast<void *>(side),
                      reinterpret_cast<void *>(fork), reinterpret_cast<void *>(join),
                      reinterpret_cast<void *>(stream), 0);
    if (rc != 0) return py::int_(rc);
    return py::cast(out);
}

// RetinaDecoder / FCOSDecoder: arg-max sweep + select kernel writing into a fresh pinned host block.
// p = (is_fcos, topn, max_out, nms_type, min_score, nms_threshold).  from_keys: the scratch already
// holds this call's keys / classes (hand-over): selection only, with the device-side `stale` check.
// Returns the pinned float32 tensor [6 * B * max_out + 4] (scores | classes | boxes | stale word), an
// int error code, or None.
py::object decode_run(uintptr_t geo_addr, const py::list &cls, const py::list &reg, const py::object &ctr,
                      const py::tuple &p, uintptr_t scratch, size_t classes_off, size_t ws_off,
                      size_t ws_bytes, uintptr_t half_table_f16, uintptr_t stream, bool from_keys) {
    if (!g_decode || !g_from_keys) return py::none();
    const b200det_geometry *geo = reinterpret_cast<const b200det_geometry *>(geo_addr);
    if (cls.size() == 0 || !THPVariable_Check(cls[0].ptr())) return py::none();
    const at::Tensor &first = THPVariable_Unpack(cls[0].ptr());
    if (!first.is_cuda()) return py::none();
    Levels lv;
    if (!marshal(geo, cls, reg, ctr, first.get_device(), &lv)) return py::none();
    b200det_decode_params dp;
    dp.is_fcos = p[0].cast<int>();
    dp.topn = p[1].cast<int>();
    dp.max_out = p[2].cast<int>();
    dp.nms_type = p[3].cast<int>();
    dp.min_score = p[4].cast<float>();
    dp.nms_threshold = p[5].cast<double>();
    dp.scales = nullptr;
    dp.sizes = nullptr;
    dp.to_xywh = 0;
    if ((dp.is_fcos != 0) != lv.has_ctr) return py::none();
    // the decoders leave torch (NumPy on the host): exp on a half array rounds to half (decode.py)
    dp.reg_dtype = lv.reg_dtype == B200DET_F32 ? lv.reg_dtype : (lv.reg_dtype | B200DET_REG_EXP_ROUNDED);
    dp.half_exp_table = lv.reg_dtype == B200DET_F16 ? reinterpret_cast<const uint16_t *>(half_table_f16)
                                                    : nullptr;
    const long long n_out = (long long)6 * geo->batch * dp.max_out;
    at::Tensor out = at::empty({n_out + 4}, at::TensorOptions().dtype(at::kFloat).pinned_memory(true));
    char *base = reinterpret_cast<char *>(scratch);
    int rc;
    if (from_keys) {
        int32_t *stale = reinterpret_cast<int32_t *>(out.data_ptr<float>() + n_out);
        *stale = 0;
        rc = g_from_keys(geo, &dp, lv.cls, lv.has_ctr ? lv.ctr : nullptr, lv.reg,
                         reinterpret_cast<const uint32_t *>(base),
                         reinterpret_cast<const int32_t *>(base + classes_off), out.data_ptr<float>(),
                         nullptr, nullptr, nullptr, stale, 1, reinterpret_cast<void *>(stream));
    } else {
        rc = g_decode(geo, &dp, lv.cls, lv.has_ctr ? lv.ctr : nullptr, lv.reg,
                      reinterpret_cast<uint32_t *>(base), reinterpret_cast<int32_t *>(base + classes_off),
                      out.data_ptr<float>(), nullptr, nullptr, nullptr, base + ws_off, ws_bytes,
                      reinterpret_cast<void *>(stream));
    }
    if (rc != 0) return py::int_(rc);
    return py::cast(out);
}

}  // namespace

PYBIND11_MODULE(_fastpath, m) {
    m.doc() = "b200det host-side fast path (argument marshalling only; calls the C ABI)";
    m.def("bind", &bind);
    m.def("loss_eval", &loss_eval);
    m.def("decode_run", &decode_run);
}
