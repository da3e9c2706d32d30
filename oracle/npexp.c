/*
 * oracle/npexp.c -- TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
 *
 * CPU restatement of the float32 `np.exp` that the reference decoders call
 * (reference: simpleAICV/detection/decode.py:260 RetinaDecoder, :356 FCOSDecoder).
 *
 * The arithmetic lives in a third-party dependency that is NOT under /root/reference:
 * NumPy (unpinned in the reference's ReadMe; 2.3.5 in this image).  On x86 hosts with
 * AVX512F or AVX2+FMA3, NumPy's float32 exp is a SIMD kernel (published in
 * numpy/_core/src/umath/loops_exponent_log.dispatch.c.src, "simd_exp_FLOAT"):
 *   q   = rint(x * log2(e))                       (add/sub 1.5*2^23)
 *   r   = fma(q, -ln2_hi, x); r = fma(q, -ln2_lo, r)   (Cody-Waite)
 *   e^r = P5(r) / Q2(r)                           (Horner, every step one fma; IEEE divide)
 *   out = scalef(e^r, q)
 * with inputs >= 88.7228394 -> +inf, <= -103.972084 -> 0, NaN -> NaN.
 *
 * PINNED: this file was compared against np.exp (numpy 2.3.5, AVX512F host) on ALL 2^32
 * float32 bit patterns: 0 mismatches (script: tests/golden/check_npexp_exhaustive.py).
 * It is not correctly rounded (39 % of results differ from CR-exp by >= 1 ulp); the CUDA
 * decoders reproduce this exact operation sequence so that int32-truncated box
 * coordinates are bit-identical to the reference's.
 */
#include <math.h>
#include <stdint.h>

static inline float npexp_scalar(float x0)
{
    const float xmax = 88.72283935546875f;
    const float xmin = -103.97208404541015625f;
    const float ln2_hi = -6.93145752e-1f;
    const float ln2_lo = -1.42860677e-6f;
    const float p0 = 9.999999999980870924916e-01f;
    const float p1 = 7.257664613233124478488e-01f;
    const float p2 = 2.473615434895520810817e-01f;
    const float p3 = 5.114512081637298353406e-02f;
    const float p4 = 6.757896990527504603057e-03f;
    const float p5 = 5.082762527590693718096e-04f;
    const float q0 = 1.0f;
    const float q1 = -2.742335390411667452936e-01f;
    const float q2 = 2.159509375685829852307e-02f;
    const float magic = 0x1.8p+23f;
    const float log2e = 1.442695040888963407359924681001892137f;

    if (x0 != x0) return NAN;
    if (x0 >= xmax) return INFINITY;
    if (x0 <= xmin) return 0.0f;

    volatile float qv = x0 * log2e;   /* volatile: keep the add/sub rounding trick intact */
    qv = qv + magic;
    qv = qv - magic;
    const float q = qv;

    float r = fmaf(q, ln2_hi, x0);
    r = fmaf(q, ln2_lo, r);

    float num = fmaf(p5, r, p4);
    num = fmaf(num, r, p3);
    num = fmaf(num, r, p2);
    num = fmaf(num, r, p1);
    num = fmaf(num, r, p0);
    float den = fmaf(q2, r, q1);
    den = fmaf(den, r, q0);

    return ldexpf(num / den, (int)q);
}

void oracle_npexp_f32(const float *in, float *out, long n)
{
    for (long i = 0; i < n; ++i) out[i] = npexp_scalar(in[i]);
}
