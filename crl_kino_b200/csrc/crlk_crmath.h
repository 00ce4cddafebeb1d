// Correctly rounded atan2 and sincos for the float64 planner kernels.
//
// The reference scores every rollout with np.arctan2 on Python scalars (dwa.py:58) and integrates
// headings with math.cos / math.sin (differential_env.py:98-99): both end in the C library, whose
// atan2 / sin / cos return the correctly rounded result (IBM accurate mathematical library:
// double-double evaluation with a rounding test).  The DWA argmin is taken over ~4,000 costs that
// are normalised by their SUM, so a last-bit difference in any heading cost moves every cost; an
// implementation that is merely "1-2 ulp" picks a different rollout whenever two candidates are
// within a few ulp -- which the dense-map grid produces all the time (the two fastest velocity
// rows saturate to commands one ulp apart).  So the kernels evaluate these functions the same
// way: a double-double fast path with an a-priori error bound and Ziv's rounding test, and a slower,
// more precise path for the ~0.1 % of arguments whose fast result lies too close to a rounding
// boundary.  With the same inputs the results are then bit-identical to the reference's.
//
// Plain C++ (host) and CUDA (device) compile this header: tools/cr_check.cpp runs the host build
// against glibc on 10^8 arguments; tests/test_gpu_dwa.py runs the device build against math.atan2 / sin / cos.
#pragma once
#include <math.h>
#include <stdint.h>

#include "crlk_cr_tables.h"

#ifdef __CUDACC__
#define CRM_FN __device__ __forceinline__
#define CRM_SLOW static __device__ __noinline__
#else
#define CRM_FN static inline
#define CRM_SLOW static __attribute__((noinline))
#endif

// The slow path of atan2 reads two more fields than the fast path; an includer that stages only the
// fast fields (shared memory) points CRM_ATAN_FULL at the complete table.
#ifndef CRM_ATAN_FULL
#define CRM_ATAN_FULL(tab) (tab)
#endif

// atan table fields ([field][k], see tools/gen_cr_tables.py)
#define CRM_A1H 8
#define CRM_A1L 9
#define CRM_A2 10
#define CRM_A2L 19
#define CRM_A3L 20

struct crm_dd { double h, l; };

CRM_FN crm_dd crm_two_sum(double a, double b)
{
    crm_dd r;
    r.h = a + b;
    const double z = r.h - a;
    r.l = (a - (r.h - z)) + (b - z);
    return r;
}
CRM_FN crm_dd crm_fast_two_sum(double a, double b)      // |a| >= |b| (or a == 0)
{
    crm_dd r;
    r.h = a + b;
    r.l = b - (r.h - a);
    return r;
}
CRM_FN crm_dd crm_two_prod(double a, double b)
{
    crm_dd r;
    r.h = a * b;
    r.l = fma(a, b, -r.h);
    return r;
}
CRM_FN crm_dd crm_dd_add(crm_dd a, crm_dd b)
{
    crm_dd s = crm_two_sum(a.h, b.h);
    s.l += a.l + b.l;
    return crm_fast_two_sum(s.h, s.l);
}
CRM_FN crm_dd crm_dd_mul(crm_dd a, crm_dd b)
{
    crm_dd p = crm_two_prod(a.h, b.h);
    p.l += fma(a.h, b.l, a.l * b.h);
    return crm_fast_two_sum(p.h, p.l);
}

// RN(h + l) when the pair carries an absolute error of at most err: true if the rounding is certain.
CRM_FN bool crm_round_ok(double h, double l, double err)
{
    const double lo = l - err, hi = l + err;
    return (h + lo) == (h + hi);
}

// 1 / a for a normal, positive a (device: hardware seed + two Newton steps; host: the division)
CRM_FN double crm_rcp(double a)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    r = fma(fma(-a, r, 1.0), r, r);
    r = fma(fma(-a, r, 1.0), r, r);
    return r;
#else
    return 1.0 / a;
#endif
}

// ---------------------------------------------------------------------------------- atan2

// Slow path: the quadratic and cubic terms in double-double as well (error ~2^-85 of the result).
CRM_SLOW double crm_atan2_slow(const double *tab, int nk, int k, int cs, double d, double ql, bool neg)
{
    const crm_dd t = crm_fast_two_sum(d, ql);
    const crm_dd t2 = crm_dd_mul(t, t);
    const crm_dd t3 = crm_dd_mul(t2, t);
    const double dt = t.h;
    double w = tab[(CRM_A2 + 8) * nk + k];
    for (int n = 7; n >= 2; n--) w = fma(w, dt, tab[(CRM_A2 + n) * nk + k]);      // a_10 .. a_4
    const double tail = w * (t2.h * t2.h);
    crm_dd a1 = {tab[CRM_A1H * nk + k], tab[CRM_A1L * nk + k]};
    crm_dd a2 = {tab[CRM_A2 * nk + k], tab[CRM_A2L * nk + k]};
    crm_dd a3 = {tab[(CRM_A2 + 1) * nk + k], tab[CRM_A3L * nk + k]};
    crm_dd s = crm_dd_mul(a3, t3);
    s.l += tail;
    s = crm_dd_add(crm_dd_mul(a2, t2), s);
    s = crm_dd_add(crm_dd_mul(a1, t), s);
    if (neg) { s.h = -s.h; s.l = -s.l; }
    crm_dd c = {tab[(2 * cs) * nk + k], tab[(2 * cs + 1) * nk + k]};
    s = crm_dd_add(c, s);
    return s.h;
}

// atan2(y, x), correctly rounded, for 1e-280 < max(|y|, |x|) < 1e280 and min / max >= 2^-1000 or
// zero handled by the caller.  tab: CRLK_CR_ATAN_INIT as doubles ([field][k], nk = CRLK_CR_ATAN_NK).
CRM_FN double crm_atan2_core(double y, double x, const double *tab)
{
    const int nk = CRLK_CR_ATAN_NK;
    const double ay = fabs(y), ax = fabs(x);
    const bool steep = ay > ax;
    const double mx = steep ? ay : ax, mn = steep ? ax : ay;
    const double r = crm_rcp(mx);
    double qh = mn * r;
    qh = fma(fma(-mx, qh, mn), r, qh);
    const double ql = fma(-mx, qh, mn) * r;                  // q = qh + ql to ~2^-105
#ifdef __CUDA_ARCH__
    const int k = __double2int_rn(qh * 64.0);
#else
    const int k = (int)nearbyint(qh * 64.0);
#endif
    const double d = fma(-(double)k, 0.015625, qh);          // exact
    const double dt = d + ql;
    const bool xneg = signbit(x);                            // x < 0 or x == -0.0 (atan2(+-0, -0) = +-pi)
    const int cs = (steep ? 1 : 0) + (xneg ? 2 : 0);
    const bool neg = steep != xneg;                          // cases 1 and 2 subtract atan(q)
    // tail: sum_{n >= 2} a_n t^n in double
    double w = tab[(CRM_A2 + 8) * nk + k];
#pragma unroll
    for (int n = 7; n >= 0; n--) w = fma(w, dt, tab[(CRM_A2 + n) * nk + k]);
    const double tail = w * (dt * dt);
    // a_1 * t in double-double
    const double a1h = tab[CRM_A1H * nk + k], a1l = tab[CRM_A1L * nk + k];
    const double ph = a1h * d;
    double pl = fma(a1h, d, -ph);
    pl = fma(a1h, ql, pl);
    pl = fma(a1l, d, pl);
    const double ch = tab[(2 * cs) * nk + k], cl = tab[(2 * cs + 1) * nk + k];
    const double sph = neg ? -ph : ph;
    double low = pl + tail;
    low = neg ? -low : low;
    const crm_dd s = crm_fast_two_sum(ch, sph);              // |ch| >= |ph| (ch == 0 only for k == 0, case 0)
    low = (s.l + low) + cl;
    const double H = s.h + low;
    const double L = (s.h - H) + low;
    // error of the double part: |tail| <= 0.7 t^2 computed to 2^-51 relative, plus the product's tail
    const double err = fabs(H) * 1.1102230246251565e-19 + 1e-300;      // 2^-63 |H|
    double res = H;
    if (!crm_round_ok(H, L, err)) res = crm_atan2_slow(CRM_ATAN_FULL(tab), nk, k, cs, d, ql, neg);
    return copysign(res, y);
}

// ---------------------------------------------------------------------------------- sincos

CRM_SLOW bool crm_sincos_slow(const double *tab, int k, double t, bool sneg, double *sn, double *cs)
{
    const int nk = CRLK_CR_SC_NK;
    const crm_dd S = {tab[k], tab[nk + k]}, C = {tab[2 * nk + k], tab[3 * nk + k]};
    const crm_dd T = {t, 0.0};
    const crm_dd t2 = crm_two_prod(t, t);
    const crm_dd t3 = crm_dd_mul(t2, T);
    const crm_dd t4 = crm_dd_mul(t2, t2);
    const double z = t2.h;
    // sin t = t - t^3/6 + t^5/120 - ...;  cos t = 1 - t^2/2 + t^4/24 - ...
    const crm_dd c6 = {-0.16666666666666666, -9.25185853854297e-18};         // -1/6
    const crm_dd c24 = {0.041666666666666664, 2.3129646346357427e-18};       // 1/24
    double ps = fma(z, fma(z, 2.7557319223985893e-06, -1.9841269841269841e-04), 8.3333333333333332e-03);   // t^5 ..
    double pc = fma(z, fma(z, -2.7557319223985888e-07, 2.4801587301587302e-05), -1.3888888888888889e-03);  // t^6 ..
    crm_dd st = crm_dd_mul(c6, t3);                       // sin t - t (leading part)
    st.l += ps * (t3.h * z);
    crm_dd cm = {-0.5 * t2.h, -0.5 * t2.l};              // cos t - 1
    crm_dd c4 = crm_dd_mul(c24, t4);
    c4.l += pc * (t4.h * z);
    cm = crm_dd_add(cm, c4);
    const crm_dd sint = crm_dd_add(T, st);
    // sin(a + t) = S + (C sin t + S (cos t - 1));  cos(a + t) = C + (C (cos t - 1) - S sin t)
    crm_dd rs = crm_dd_add(crm_dd_mul(C, sint), crm_dd_mul(S, cm));
    rs = crm_dd_add(S, rs);
    crm_dd ssn = crm_dd_mul(S, sint);
    ssn.h = -ssn.h; ssn.l = -ssn.l;
    crm_dd rc = crm_dd_add(crm_dd_mul(C, cm), ssn);
    rc = crm_dd_add(C, rc);
    // a result many orders below its terms (theta within ~1e-9 of a zero of sin / cos other than 0)
    // has lost too many of the pair's bits: leave those arguments to the library
    if ((k != 0 && fabs(rs.h) < 1e-9) || fabs(rc.h) < 1e-9) return false;
    *sn = sneg ? -rs.h : rs.h;
    *cs = rc.h;
    return true;
}

// sin and cos of th, correctly rounded, for |th| <= 3.15 (every heading the planner produces is
// normalised into [-pi, pi)); returns false for other arguments (the caller uses the library).
CRM_FN bool crm_sincos_core(double th, const double *tab, double *sn, double *cs)
{
    const int nk = CRLK_CR_SC_NK;
    const double ath = fabs(th);
    if (!(ath <= 3.15)) return false;
#ifdef __CUDA_ARCH__
    const int k = __double2int_rn(ath * 64.0);
#else
    const int k = (int)nearbyint(ath * 64.0);
#endif
    const double t = fma(-(double)k, 0.015625, ath);          // exact, |t| <= 1/128
    const double Sh = tab[k], Sl = tab[nk + k], Ch = tab[2 * nk + k], Cl = tab[3 * nk + k];
    const double z = t * t;
    // sin t - t = t^3 ps(z), cos t - 1 = z pc(z)   (|t| <= 2^-7: t^9 / 9! and t^10 / 10! are below 2^-80)
    const double ps = fma(z, fma(z, fma(z, 2.7557319223985893e-06, -1.9841269841269841e-04), 8.3333333333333332e-03),
                          -1.6666666666666666e-01);
    const double pc = fma(z, fma(z, fma(z, 2.4801587301587302e-05, -1.3888888888888889e-03), 4.1666666666666664e-02), -0.5);
    const double sm = (t * z) * ps;                           // sin t - t
    const double cm = z * pc;                                 // cos t - 1
    const bool sneg = th < 0.0;
    // ---- sin: S + C t + (C (sin t - t) + S (cos t - 1))
    {
        const double ph = Ch * t;
        double pl = fma(Ch, t, -ph);
        pl = fma(Cl, t, pl);
        const double rest = fma(Ch, sm, Sh * cm);
        const crm_dd s = crm_two_sum(Sh, ph);
        const double low = (s.l + (pl + rest)) + Sl;
        const double H = s.h + low, L = (s.h - H) + low;
        const double err = (fabs(Ch * sm) + fabs(Sh * cm)) * 8.9e-16 + 1e-300;       // 2^-50 of the double part
        if (!crm_round_ok(H, L, err)) return crm_sincos_slow(tab, k, t, sneg, sn, cs);
        *sn = sneg ? -H : H;
    }
    // ---- cos: C - S t + (C (cos t - 1) - S (sin t - t))
    {
        const double ph = -(Sh * t);
        double pl = -fma(Sh, t, ph);
        pl = fma(-Sl, t, pl);
        const double rest = fma(Ch, cm, -(Sh * sm));
        const crm_dd s = crm_two_sum(Ch, ph);
        const double low = (s.l + (pl + rest)) + Cl;
        const double H = s.h + low, L = (s.h - H) + low;
        const double err = (fabs(Sh * sm) + fabs(Ch * cm)) * 8.9e-16 + 1e-300;
        if (!crm_round_ok(H, L, err)) return crm_sincos_slow(tab, k, t, sneg, sn, cs);
        *cs = H;
    }
    return true;
}
