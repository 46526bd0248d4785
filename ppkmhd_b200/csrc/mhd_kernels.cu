// Hand-written fp64 CUDA kernels (sm_100a) for ppkMHD's 3-D MUSCL-Hancock + CT MHD step, variant
// "implementationVersion 0" (src/muscl/SolverMHDMuscl.cpp:494-517).
//
// This translation unit is compiled TWICE:
//   -DPPK_EXACT=1 --fmad=false : every expression keeps the reference's operation order and is
//                                evaluated without FMA contraction => results are bit-identical to
//                                the reference's OpenMP build (IEEE fp64 +,*,/,sqrt on both sides);
//   -DPPK_EXACT=0              : same source, FMA contraction allowed, the EMF upwind blend is a
//                                branch instead of a 0/1-weighted sum of all branches.
// Not a port of the Kokkos functors: the reference stores 18 reconstructed states per cell
// (1152 B) between its trace and Riemann kernels; here the trace kernel stores a 32-number
// "basis" per cell and the face-flux / edge-EMF kernels rebuild the states they need in registers
// (each state component is an exact 1- or 2-addition combination of basis numbers).
#include "mhd_common.h"

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#ifndef PPK_EXACT
#  error "compile with -DPPK_EXACT=0 or 1"
#endif
#if PPK_EXACT
#  define PPK_NS ppk_exact
#else
#  define PPK_NS ppk_fast
#endif

namespace ppk {
namespace PPK_NS {

#define DEV __device__ __forceinline__
// write-once streams (basis, fluxes, EMFs, the new state): optionally stored with the evict-first hint (st.global.cs)
#ifdef PPK_STREAM_STORES
#  define ST(lhs, val) __stcs(&(lhs), (val))
#else
#  define ST(lhs, val) ((lhs) = (val))
#endif

DEV long long cidx(const GridParams &g, int i, int j, int k) {
  return (long long)i + (long long)g.isize * ((long long)j + (long long)g.jsize * (long long)k);
}

// ---------------------------------------------------------------------------------------------
// device math
// ---------------------------------------------------------------------------------------------

// min/max of the floors and limiters: the exact build keeps fmin/fmax (the reference's calls); the fast build uses
// a comparison + select (3 instructions instead of the 6-7 of fmin/fmax with their NaN-quieting path on sm_100)
#if PPK_EXACT
DEV double vmax(double a, double b) { return fmax(a, b); }
DEV double vmin(double a, double b) { return fmin(a, b); }
#else
DEV double vmax(double a, double b) { return a > b ? a : b; }
DEV double vmin(double a, double b) { return a < b ? a : b; }
#endif

// TVD limited slope, MHDBaseFunctor3D.h:280-288 (hydro) and :605-665 (face B)
#if PPK_EXACT
DEV double limited_slope(double st, double q, double qplus, double qminus) {
  const double dlft = st * (q - qminus);
  const double drgt = st * (qplus - q);
  const double dcen = 0.5 * (qplus - qminus);
  const double dsgn = (dcen >= 0.0) ? 1.0 : -1.0;
  const double slop = vmin(fabs(dlft), fabs(drgt));
  double dlim = slop;
  if ((dlft * drgt) <= 0.0) dlim = 0.0;
  return dsgn * vmin(dlim, fabs(dcen));
}
#else
// The same value with 16 instructions instead of 21 (27 slopes per cell are a third of the producer's instructions):
// min(|st a|, |st b|) = st min(|a|, |b|) (st > 0, rounding is monotonic); when a and b have the same sign, qplus - qminus
// has it too, so the sign of the result is the sign of a and "a b <= 0" is a test on the two sign bits (a zero
// difference makes the minimum zero by itself). Only the sign of a zero result can differ from the expression above.
DEV double limited_slope(double st, double q, double qplus, double qminus) {
  const double a = q - qminus, b = qplus - q, cen = qplus - qminus;
  const double fa = fabs(a), fb = fabs(b);
  const double m = st * (fa < fb ? fa : fb);
  const double hc = 0.5 * fabs(cen);
  const double r = m < hc ? m : hc;
  const int ha = __double2hiint(a), hb = __double2hiint(b);
  const bool opposite = (ha ^ hb) < 0;
  const int hi = opposite ? 0 : (__double2hiint(r) | (ha & 0x80000000));
  const int lo = opposite ? 0 : __double2loint(r);
  return __hiloint2double(hi, lo);
}
#endif

// find_speed_fast<dir>, mhd_utils.h:89-117; `n` is the field component normal to the direction
DEV void fast_speed_common(double gamma0, double d, double p, double a, double b, double c, double &c2, double &d2) {
  const double b2 = a * a + b * b + c * c;
  c2 = gamma0 * p / d;
  d2 = 0.5 * (b2 / d + c2);
}
DEV double fast_speed_dir(double c2, double d2, double d, double n) {
  return sqrt(d2 + sqrt(d2 * d2 - c2 * n * n / d));
}

#if !PPK_EXACT
// ---- fast-arithmetic primitives ---------------------------------------------------------------------
// nvcc expands an fp64 '/' or sqrt() into MUFU seed + Newton steps + a guarded slow path (~10-14 FP64-pipe
// instructions, a branch and a CALL). The physical quantities here are normal, finite and non-zero, so the
// fast build uses the 20-bit MUFU seed followed by ONE cubically convergent step (error e -> e^3 = 2^-60,
// i.e. the result is within ~1 ulp): 1 MUFU + 3 (rcp) or 5 (rsqrt) FP64 instructions, no branch.
DEV double frcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}
DEV double frsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y;
  const double e = fma(-t, y, 1.0);           // 1 - x y^2
  const double p = fma(e, 0.375, 0.5) * e;    // e/2 + 3 e^2/8
  return fma(y, p, y);
}
// sqrt for x >= 0 (x == 0 must give 0: the clamped discriminant of the fast speed can vanish)
DEV double fsqrt(double x) { return x * frsqrt(fmax(x, 1e-300)); }
// Comparison-select min/max (DSETP + 2 SEL): fmin/fmax() cost 6-7 instructions each on sm_100 because of their
// NaN-quieting path, and the fast build has ~60 of them per edge. Arguments here are never NaN.
DEV double dmax(double a, double b) { return a > b ? a : b; }
DEV double dmin(double a, double b) { return a < b ? a : b; }
// sqrt of a quantity that is mathematically >= 0 but may come out <= 0 by round-off (the discriminant of the fast
// speed): everything below the smallest normal, including negatives, is replaced by 2^-1000 (sqrt = 1e-150 ~ 0)
// with 3 integer-pipe instructions on the high word, no FP64-pipe slot.
DEV double fsqrt_clamped(double x) {
  const int hi = __double2hiint(x);
  const bool tiny = hi < 0x00100000;  // negative doubles have a negative high word
  const double xc = __hiloint2double(tiny ? 0x01700000 : hi, tiny ? 0 : __double2loint(x));
  return xc * frsqrt(xc);
}
// max of POSITIVE doubles on the integer pipe (positive doubles order like their bit patterns): keeps the
// comparison chains of the squared wave speeds off the FP64 pipe, which bounds the EMF kernels
DEV double pmax(double a, double b) { return __double_as_longlong(a) > __double_as_longlong(b) ? a : b; }
DEV double pmax4(double a0, double a1, double a2, double a3) { return pmax(pmax(a0, a1), pmax(a2, a3)); }
DEV double pmax5(double a0, double a1, double a2, double a3, double a4) { return pmax(pmax4(a0, a1, a2, a3), a4); }
// sqrt of a strictly positive normal number
DEV double fsqrt_pos(double x) { return x * frsqrt(x); }
// max(x, 0) / min(x, 0) on the sign bit
DEV double clamp_lo0(double x) {
  const int hi = __double2hiint(x);
  const int m = ~(hi >> 31);
  return __hiloint2double(hi & m, __double2loint(x) & m);
}
DEV double clamp_hi0(double x) {
  const int hi = __double2hiint(x);
  const int m = hi >> 31;
  return __hiloint2double(hi & m, __double2loint(x) & m);
}

// Fast-arithmetic variant of riemann_hlld (same algebra as RiemannSolvers_MHD.h:133-367, evaluated with
// shared reciprocals: fp64 '/' and sqrt cost ~10 FP64-pipe instructions each and were 3/4 of the kernel).
// 34 divisions+square roots become 15. Differences to the reference are a few ulp per operation, far
// inside the 1e-12 per-cell tolerance (tests/test_gpu_parity.py::test_fast_mode_*).
DEV void riemann_hlld_fast(double gamma0, double rl, double pl, double ul, double vl, double wl, double al, double bl,
                           double cl, double rr, double pr, double ur, double vr, double wr, double ar, double br,
                           double cr, double &f_d, double &f_p, double &f_u, double &f_v, double &f_w,
                           double *f_bz = nullptr) {
  const double entho = frcp(gamma0 - 1.0);
  const double a = 0.5 * (al + ar);
  const double a2 = a * a;
  const double sgnm = (a >= 0) ? 1.0 : -1.0;

  const double ecinl = 0.5 * (ul * ul + vl * vl + wl * wl) * rl;
  const double emagl = 0.5 * (a2 + bl * bl + cl * cl);
  const double etotl = pl * entho + ecinl + emagl;
  const double ptotl = pl + emagl;
  const double vdotbl = ul * a + vl * bl + wl * cl;
  const double ecinr = 0.5 * (ur * ur + vr * vr + wr * wr) * rr;
  const double emagr = 0.5 * (a2 + br * br + cr * cr);
  const double etotr = pr * entho + ecinr + emagr;
  const double ptotr = pr + emagr;
  const double vdotbr = ur * a + vr * br + wr * cr;

  // fast magnetosonic speeds: max(fsqrt(x),fsqrt(y)) = fsqrt(max(x,y))
  const double irl = frcp(rl), irr = frcp(rr);
  const double c2l = gamma0 * pl * irl, d2l = 0.5 * (2.0 * emagl * irl + c2l);
  const double c2r = gamma0 * pr * irr, d2r = 0.5 * (2.0 * emagr * irr + c2r);
  const double cf2l = d2l + fsqrt_clamped(d2l * d2l - c2l * a2 * irl);
  const double cf2r = d2r + fsqrt_clamped(d2r * d2r - c2r * a2 * irr);
  const double cfmax = fsqrt_pos(pmax(cf2l, cf2r));
  const double sl = dmin(ul, ur) - cfmax;
  const double sr = dmax(ul, ur) + cfmax;

  const double rcl = rl * (ul - sl);
  const double rcr = rr * (sr - ur);
  const double irc = frcp(rcr + rcl);
  const double ustar = (rcr * ur + rcl * ul + (ptotl - ptotr)) * irc;
  const double ptotstar = (rcr * ptotl + rcl * ptotr + rcl * rcr * (ul - ur)) * irc;

  // left star region
  const double isl = frcp(sl - ustar);
  const double rstarl = rl * (sl - ul) * isl;
  double estar = rl * (sl - ul) * (sl - ustar) - a2;
  const double el = rl * (sl - ul) * (sl - ul) - a2;
  double vstarl, wstarl, bstarl, cstarl;
  if (a2 > 0 && fabs(estar - a2) <= 1e-8 * a2) {  // |estar/a^2 - 1| <= 1e-8
    vstarl = vl; bstarl = bl; wstarl = wl; cstarl = cl;
  } else {
    const double ie = frcp(estar);
    const double k1 = a * (ustar - ul) * ie, k2 = el * ie;
    vstarl = vl - bl * k1; bstarl = bl * k2;
    wstarl = wl - cl * k1; cstarl = cl * k2;
  }
  const double vdotbstarl = ustar * a + vstarl * bstarl + wstarl * cstarl;
  const double etotstarl = ((sl - ul) * etotl - ptotl * ul + ptotstar * ustar + a * (vdotbl - vdotbstarl)) * isl;
  const double rsql = frsqrt(rstarl);
  const double sqrrstarl = rstarl * rsql;
  const double sal = ustar - fabs(a) * rsql;

  // right star region
  const double isr = frcp(sr - ustar);
  const double rstarr = rr * (sr - ur) * isr;
  estar = rr * (sr - ur) * (sr - ustar) - a2;
  const double er = rr * (sr - ur) * (sr - ur) - a2;
  double vstarr, wstarr, bstarr, cstarr;
  if (a2 > 0 && fabs(estar - a2) <= 1e-8 * a2) {
    vstarr = vr; bstarr = br; wstarr = wr; cstarr = cr;
  } else {
    const double ie = frcp(estar);
    const double k1 = a * (ustar - ur) * ie, k2 = er * ie;
    vstarr = vr - br * k1; bstarr = br * k2;
    wstarr = wr - cr * k1; cstarr = cr * k2;
  }
  const double vdotbstarr = ustar * a + vstarr * bstarr + wstarr * cstarr;
  const double etotstarr = ((sr - ur) * etotr - ptotr * ur + ptotstar * ustar + a * (vdotbr - vdotbstarr)) * isr;
  const double rsqr = frsqrt(rstarr);
  const double sqrrstarr = rstarr * rsqr;
  const double sar = ustar + fabs(a) * rsqr;

  // sample the fan at x/t = 0 (the double-star state is only built when it is selected)
  double ro, uo, vo, wo, bo, co, ptoto, etoto, vdotbo;
  if (sl > 0) {
    ro = rl; uo = ul; vo = vl; wo = wl; bo = bl; co = cl; ptoto = ptotl; etoto = etotl; vdotbo = vdotbl;
  } else if (sal > 0) {
    ro = rstarl; uo = ustar; vo = vstarl; wo = wstarl; bo = bstarl; co = cstarl; ptoto = ptotstar; etoto = etotstarl; vdotbo = vdotbstarl;
  } else if (sar > 0) {
    const double iss = frcp(sqrrstarl + sqrrstarr);
    const double ss = sgnm * sqrrstarl * sqrrstarr;
    vo = (sqrrstarl * vstarl + sqrrstarr * vstarr + sgnm * (bstarr - bstarl)) * iss;
    wo = (sqrrstarl * wstarl + sqrrstarr * wstarr + sgnm * (cstarr - cstarl)) * iss;
    bo = (sqrrstarl * bstarr + sqrrstarr * bstarl + ss * (vstarr - vstarl)) * iss;
    co = (sqrrstarl * cstarr + sqrrstarr * cstarl + ss * (wstarr - wstarl)) * iss;
    vdotbo = ustar * a + vo * bo + wo * co;
    uo = ustar; ptoto = ptotstar;
    if (ustar > 0) { ro = rstarl; etoto = etotstarl - sgnm * sqrrstarl * (vdotbstarl - vdotbo); }
    else { ro = rstarr; etoto = etotstarr + sgnm * sqrrstarr * (vdotbstarr - vdotbo); }
  } else if (sr > 0) {
    ro = rstarr; uo = ustar; vo = vstarr; wo = wstarr; bo = bstarr; co = cstarr; ptoto = ptotstar; etoto = etotstarr; vdotbo = vdotbstarr;
  } else {
    ro = rr; uo = ur; vo = vr; wo = wr; bo = br; co = cr; ptoto = ptotr; etoto = etotr; vdotbo = vdotbr;
  }
  f_d = ro * uo;
  f_p = (etoto + ptoto) * uo - a * vdotbo;
  f_u = ro * uo * uo - a2 + ptoto;
  f_v = ro * uo * vo - a * bo;
  f_w = ro * uo * wo - a * co;
  if (f_bz) *f_bz = co * uo - a * wo;  // induction flux of the out-of-plane field: the 2-D path only (:366)
}
#endif

// riemann_hlld, RiemannSolvers_MHD.h:133-367. Inputs are in the frame of the face normal
// (un,bn normal; t1,t2 transverse). Only the 5 hydro fluxes are produced: the induction part of
// the reference's flux vector is never used (the field is advanced by the edge EMFs).
DEV void riemann_hlld(double gamma0, double rl, double pl, double ul, double vl, double wl, double al, double bl,
                      double cl, double rr, double pr, double ur, double vr, double wr, double ar, double br,
                      double cr, double &f_d, double &f_p, double &f_u, double &f_v, double &f_w, double *f_bz = nullptr) {
  const double entho = 1.0 / (gamma0 - 1.0);
  const double a = 0.5 * (al + ar);
  const double sgnm = (a >= 0) ? 1.0 : -1.0;

  const double ecinl = 0.5 * (ul * ul + vl * vl + wl * wl) * rl;
  const double emagl = 0.5 * (a * a + bl * bl + cl * cl);
  const double etotl = pl * entho + ecinl + emagl;
  const double ptotl = pl + emagl;
  const double vdotbl = ul * a + vl * bl + wl * cl;

  const double ecinr = 0.5 * (ur * ur + vr * vr + wr * wr) * rr;
  const double emagr = 0.5 * (a * a + br * br + cr * cr);
  const double etotr = pr * entho + ecinr + emagr;
  const double ptotr = pr + emagr;
  const double vdotbr = ur * a + vr * br + wr * cr;

  double c2, d2;
  fast_speed_common(gamma0, rl, pl, a, bl, cl, c2, d2);
  const double cfastl = fast_speed_dir(c2, d2, rl, a);
  fast_speed_common(gamma0, rr, pr, a, br, cr, c2, d2);
  const double cfastr = fast_speed_dir(c2, d2, rr, a);

  const double cfmax = fmax(cfastl, cfastr);
  const double sl = fmin(ul, ur) - cfmax;
  const double sr = fmax(ul, ur) + cfmax;

  const double rcl = rl * (ul - sl);
  const double rcr = rr * (sr - ur);

  const double ustar = (rcr * ur + rcl * ul + (ptotl - ptotr)) / (rcr + rcl);
  const double ptotstar = (rcr * ptotl + rcl * ptotr + rcl * rcr * (ul - ur)) / (rcr + rcl);

  // left star region
  const double rstarl = rl * (sl - ul) / (sl - ustar);
  double estar = rl * (sl - ul) * (sl - ustar) - a * a;
  const double el = rl * (sl - ul) * (sl - ul) - a * a;
  double vstarl, wstarl, bstarl, cstarl;
  if (a * a > 0 && fabs(estar / (a * a) - 1.0) <= 1e-8) {
    vstarl = vl; bstarl = bl; wstarl = wl; cstarl = cl;
  } else {
    vstarl = vl - a * bl * (ustar - ul) / estar;
    bstarl = bl * el / estar;
    wstarl = wl - a * cl * (ustar - ul) / estar;
    cstarl = cl * el / estar;
  }
  const double vdotbstarl = ustar * a + vstarl * bstarl + wstarl * cstarl;
  const double etotstarl = ((sl - ul) * etotl - ptotl * ul + ptotstar * ustar + a * (vdotbl - vdotbstarl)) / (sl - ustar);
  const double sqrrstarl = sqrt(rstarl);
  const double calfvenl = fabs(a) / sqrrstarl;
  const double sal = ustar - calfvenl;

  // right star region
  const double rstarr = rr * (sr - ur) / (sr - ustar);
  estar = rr * (sr - ur) * (sr - ustar) - a * a;
  const double er = rr * (sr - ur) * (sr - ur) - a * a;
  double vstarr, wstarr, bstarr, cstarr;
  if (a * a > 0 && fabs(estar / (a * a) - 1.0) <= 1e-8) {
    vstarr = vr; bstarr = br; wstarr = wr; cstarr = cr;
  } else {
    vstarr = vr - a * br * (ustar - ur) / estar;
    bstarr = br * er / estar;
    wstarr = wr - a * cr * (ustar - ur) / estar;
    cstarr = cr * er / estar;
  }
  const double vdotbstarr = ustar * a + vstarr * bstarr + wstarr * cstarr;
  const double etotstarr = ((sr - ur) * etotr - ptotr * ur + ptotstar * ustar + a * (vdotbr - vdotbstarr)) / (sr - ustar);
  const double sqrrstarr = sqrt(rstarr);
  const double calfvenr = fabs(a) / sqrrstarr;
  const double sar = ustar + calfvenr;

  // double star region
  const double vstarstar = (sqrrstarl * vstarl + sqrrstarr * vstarr + sgnm * (bstarr - bstarl)) / (sqrrstarl + sqrrstarr);
  const double wstarstar = (sqrrstarl * wstarl + sqrrstarr * wstarr + sgnm * (cstarr - cstarl)) / (sqrrstarl + sqrrstarr);
  const double bstarstar =
    (sqrrstarl * bstarr + sqrrstarr * bstarl + sgnm * sqrrstarl * sqrrstarr * (vstarr - vstarl)) / (sqrrstarl + sqrrstarr);
  const double cstarstar =
    (sqrrstarl * cstarr + sqrrstarr * cstarl + sgnm * sqrrstarl * sqrrstarr * (wstarr - wstarl)) / (sqrrstarl + sqrrstarr);
  const double vdotbstarstar = ustar * a + vstarstar * bstarstar + wstarstar * cstarstar;
  const double etotstarstarl = etotstarl - sgnm * sqrrstarl * (vdotbstarl - vdotbstarstar);
  const double etotstarstarr = etotstarr + sgnm * sqrrstarr * (vdotbstarr - vdotbstarstar);

  // sample the fan at x/t = 0
  double ro, uo, vo, wo, bo, co, ptoto, etoto, vdotbo;
  if (sl > 0) {
    ro = rl; uo = ul; vo = vl; wo = wl; bo = bl; co = cl; ptoto = ptotl; etoto = etotl; vdotbo = vdotbl;
  } else if (sal > 0) {
    ro = rstarl; uo = ustar; vo = vstarl; wo = wstarl; bo = bstarl; co = cstarl; ptoto = ptotstar; etoto = etotstarl; vdotbo = vdotbstarl;
  } else if (ustar > 0) {
    ro = rstarl; uo = ustar; vo = vstarstar; wo = wstarstar; bo = bstarstar; co = cstarstar; ptoto = ptotstar; etoto = etotstarstarl; vdotbo = vdotbstarstar;
  } else if (sar > 0) {
    ro = rstarr; uo = ustar; vo = vstarstar; wo = wstarstar; bo = bstarstar; co = cstarstar; ptoto = ptotstar; etoto = etotstarstarr; vdotbo = vdotbstarstar;
  } else if (sr > 0) {
    ro = rstarr; uo = ustar; vo = vstarr; wo = wstarr; bo = bstarr; co = cstarr; ptoto = ptotstar; etoto = etotstarr; vdotbo = vdotbstarr;
  } else {
    ro = rr; uo = ur; vo = vr; wo = wr; bo = br; co = cr; ptoto = ptotr; etoto = etotr; vdotbo = vdotbr;
  }

  f_d = ro * uo;
  f_p = (etoto + ptoto) * uo - a * vdotbo;
  f_u = ro * uo * uo - a * a + ptoto;
  f_v = ro * uo * vo - a * bo;
  f_w = ro * uo * wo - a * co;
  if (f_bz) *f_bz = co * uo - a * wo;  // flux[IBZ], RiemannSolvers_MHD.h:366 (2-D path only)
}

// find_mhd_flux (mhd_utils.h:175-231, cIso == 0), hydro part: conservative variables and fluxes of (rho, E, mn, mt1, mt2)
DEV void mhd_flux5(double gamma0, double d, double p, double u, double v, double w, double a, double b, double c,
                   double cv[5], double ff[5]) {
  const double entho = 1.0 / (gamma0 - 1.0);
  const double ecin = 0.5 * (u * u + v * v + w * w) * d;
  const double emag = 0.5 * (a * a + b * b + c * c);
  const double etot = p * entho + ecin + emag;
  const double ptot = p + emag;
  cv[0] = d; cv[1] = etot; cv[2] = d * u; cv[3] = d * v; cv[4] = d * w;
  ff[0] = d * u;
  ff[1] = (etot + ptot) * u - a * (a * u + b * v + c * w);
  ff[2] = d * u * u - a * a + ptot;
  ff[3] = d * u * v - a * b;
  ff[4] = d * u * w - a * c;
}
// riemann_hll (RiemannSolvers_MHD.h:27-69) and riemann_llf (:83-111), hydro fluxes only (the induction components of
// the flux vector are never read: the field is advanced by the edge EMFs). Same frame as riemann_hlld.
DEV void riemann_hll(double gamma0, double rl, double pl, double ul, double vl, double wl, double al, double bl, double cl,
                     double rr, double pr, double ur, double vr, double wr, double ar, double br, double cr, double f[5],
                     double *f_bz = nullptr) {
  const double a = 0.5 * (al + ar);
  double cl5[5], fl5[5], cr5[5], fr5[5];
  mhd_flux5(gamma0, rl, pl, ul, vl, wl, a, bl, cl, cl5, fl5);
  mhd_flux5(gamma0, rr, pr, ur, vr, wr, a, br, cr, cr5, fr5);
  double c2, d2;
  fast_speed_common(gamma0, rl, pl, a, bl, cl, c2, d2);
  const double cfl_ = fast_speed_dir(c2, d2, rl, a);
  fast_speed_common(gamma0, rr, pr, a, br, cr, c2, d2);
  const double cfr = fast_speed_dir(c2, d2, rr, a);
  const double sl = fmin(fmin(ul, ur) - fmax(cfl_, cfr), 0.0);
  const double sr = fmax(fmax(ul, ur) + fmax(cfl_, cfr), 0.0);
#pragma unroll
  for (int v = 0; v < 5; ++v) f[v] = (sr * fl5[v] - sl * fr5[v] + sr * sl * (cr5[v] - cl5[v])) / (sr - sl);
  // out-of-plane field (2-D path only): conservative variable c, flux c*u - a*w (mhd_utils.h:219, 229)
  if (f_bz) *f_bz = (sr * (cl * ul - a * wl) - sl * (cr * ur - a * wr) + sr * sl * (cr - cl)) / (sr - sl);
}
DEV void riemann_llf(double gamma0, double rl, double pl, double ul, double vl, double wl, double al, double bl, double cl,
                     double rr, double pr, double ur, double vr, double wr, double ar, double br, double cr, double f[5],
                     double *f_bz = nullptr) {
  const double a = 0.5 * (al + ar);
  double cl5[5], fl5[5], cr5[5], fr5[5];
  mhd_flux5(gamma0, rl, pl, ul, vl, wl, a, bl, cl, cl5, fl5);
  mhd_flux5(gamma0, rr, pr, ur, vr, wr, a, br, cr, cr5, fr5);
  double c2, d2;
  fast_speed_common(gamma0, rl, pl, a, bl, cl, c2, d2);
  const double cleft = fast_speed_dir(c2, d2, rl, a) + fabs(ul);  // find_speed_info, mhd_utils.h:377-402
  fast_speed_common(gamma0, rr, pr, a, br, cr, c2, d2);
  const double cright = fast_speed_dir(c2, d2, rr, a) + fabs(ur);
  const double vel_info = fmax(cleft, cright);
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    f[v] = (fl5[v] + fr5[v]) / 2;
    f[v] -= vel_info * (cr5[v] - cl5[v]) / 2;
  }
  if (f_bz) {
    double fb = ((cl * ul - a * wl) + (cr * ur - a * wr)) / 2;
    fb -= vel_info * (cr - cl) / 2;
    *f_bz = fb;
  }
}
// riemann_mhd (RiemannSolvers_MHD.h:372-392): the solver [hydro] riemann= selects (uniform over the grid).
// RS >= 0 fixes the solver at compile time (the hot TMA kernels are instantiated for HLLD alone), RS < 0 reads g.riemann.
template <int RS = -1>
DEV void riemann_face(const GridParams &g, double rl, double pl, double ul, double vl, double wl, double al, double bl, double cl,
                      double rr, double pr, double ur, double vr, double wr, double ar, double br, double cr, double &f_d,
                      double &f_p, double &f_u, double &f_v, double &f_w, double *f_bz = nullptr) {
  const int rs = RS >= 0 ? RS : g.riemann;
  if (rs == RIEMANN_HLLD) {
#if PPK_EXACT
    riemann_hlld(g.gamma0, rl, pl, ul, vl, wl, al, bl, cl, rr, pr, ur, vr, wr, ar, br, cr, f_d, f_p, f_u, f_v, f_w, f_bz);
#else
    riemann_hlld_fast(g.gamma0, rl, pl, ul, vl, wl, al, bl, cl, rr, pr, ur, vr, wr, ar, br, cr, f_d, f_p, f_u, f_v, f_w, f_bz);
#endif
  } else {
    double f[5];
    if (rs == RIEMANN_HLL) riemann_hll(g.gamma0, rl, pl, ul, vl, wl, al, bl, cl, rr, pr, ur, vr, wr, ar, br, cr, f, f_bz);
    else riemann_llf(g.gamma0, rl, pl, ul, vl, wl, al, bl, cl, rr, pr, ur, vr, wr, ar, br, cr, f, f_bz);
    f_d = f[0]; f_p = f[1]; f_u = f[2]; f_v = f[3]; f_w = f[4];
  }
}

// comparison chains of mhd_utils.h:36-77 (not fmax/fmin)
DEV double max4(double a0, double a1, double a2, double a3) {
  double r = a0; r = (a1 > r) ? a1 : r; r = (a2 > r) ? a2 : r; r = (a3 > r) ? a3 : r; return r;
}
DEV double min4(double a0, double a1, double a2, double a3) {
  double r = a0; r = (a1 < r) ? a1 : r; r = (a2 < r) ? a2 : r; r = (a3 < r) ? a3 : r; return r;
}
DEV double max5(double a0, double a1, double a2, double a3, double a4) {
  double r = a0; r = (a1 > r) ? a1 : r; r = (a2 > r) ? a2 : r; r = (a3 > r) ? a3 : r; r = (a4 > r) ? a4 : r; return r;
}

// One corner state of the 2-D magnetic Riemann problem: r,p, the two in-plane velocities and the
// three field components in the (d1,d2,e) frame. (The out-of-plane velocity is never read.)
struct Corner { double r, p, u, v, a, b, c; };

#if !PPK_EXACT
// Fast-arithmetic variant of mag_riemann2d_hlld (RiemannSolvers_MHD.h:398-630): identical algebra with
//  * one reciprocal per distinct denominator (1/rho, 1/(S-ustar), 1/(S-vstar), 1/(SAR-SAL), 1/(SAT-SAB)),
//  * max_i sqrt(x_i) evaluated as sqrt(max_i x_i) for the fast and Alfven speeds,
//  * |Astar|/sqrt(rstar) = (|a|/sqrt(rstar_x)) * sqrt(gy) folded into the squared comparison,
//  * the upwind selection as a branch.
// 98 divisions + square roots become 34.
struct CornerAux { double gx, gy, igx, igy; };
DEV double mag_riemann2d_hlld_fast(double gamma0, double smallc, const Corner &LL, const Corner &RL, const Corner &LR,
                                   const Corner &RR, double ELL, double ERL, double ELR, double ERR) {
  const double iLL = frcp(LL.r), iLR = frcp(LR.r), iRL = frcp(RL.r), iRR = frcp(RR.r);
  auto mag2 = [](const Corner &q) { return q.a * q.a + q.b * q.b + q.c * q.c; };
  const double m2LL = mag2(LL), m2LR = mag2(LR), m2RL = mag2(RL), m2RR = mag2(RR);
  // squared fast speeds along x (normal field a) and y (normal field b)
  auto cf2 = [&](const Corner &q, double ir, double m2, double &cx2, double &cy2) {
    const double c2 = gamma0 * q.p * ir;
    const double d2 = 0.5 * (m2 * ir + c2);
    const double dd = d2 * d2, k = c2 * ir;
    cx2 = d2 + fsqrt_clamped(dd - k * q.a * q.a);
    cy2 = d2 + fsqrt_clamped(dd - k * q.b * q.b);
  };
  double xLL, yLL, xLR, yLR, xRL, yRL, xRR, yRR;
  cf2(LL, iLL, m2LL, xLL, yLL); cf2(LR, iLR, m2LR, xLR, yLR); cf2(RL, iRL, m2RL, xRL, yRL); cf2(RR, iRR, m2RR, xRR, yRR);
  const double cxmax = fsqrt_pos(pmax4(xLL, xLR, xRL, xRR));
  const double cymax = fsqrt_pos(pmax4(yLL, yLR, yRL, yRR));
  const double SL = min4(LL.u, LR.u, RL.u, RR.u) - cxmax;
  const double SR = max4(LL.u, LR.u, RL.u, RR.u) + cxmax;
  const double SB = min4(LL.v, LR.v, RL.v, RR.v) - cymax;
  const double ST = max4(LL.v, LR.v, RL.v, RR.v) + cymax;

  const double PtotLL = LL.p + 0.5 * m2LL, PtotLR = LR.p + 0.5 * m2LR, PtotRL = RL.p + 0.5 * m2RL, PtotRR = RR.p + 0.5 * m2RR;
  const double rcLLx = LL.r * (LL.u - SL), rcRLx = RL.r * (SR - RL.u), rcLRx = LR.r * (LR.u - SL), rcRRx = RR.r * (SR - RR.u);
  const double rcLLy = LL.r * (LL.v - SB), rcLRy = LR.r * (ST - LR.v), rcRLy = RL.r * (RL.v - SB), rcRRy = RR.r * (ST - RR.v);
  const double ustar = (rcLLx * LL.u + rcLRx * LR.u + rcRLx * RL.u + rcRRx * RR.u + (PtotLL - PtotRL + PtotLR - PtotRR)) *
                       frcp(rcLLx + rcLRx + rcRLx + rcRRx);
  const double vstar = (rcLLy * LL.v + rcLRy * LR.v + rcRLy * RL.v + rcRRy * RR.v + (PtotLL - PtotLR + PtotRL - PtotRR)) *
                       frcp(rcLLy + rcLRy + rcRLy + rcRRy);

  // compression factors g = (S - u)/(S - u*) of every corner in x and y, and their inverses
  const double iSL = frcp(SL - ustar), iSR = frcp(SR - ustar), iSB = frcp(SB - vstar), iST = frcp(ST - vstar);
  const double gxLL = (SL - LL.u) * iSL, gxLR = (SL - LR.u) * iSL, gxRL = (SR - RL.u) * iSR, gxRR = (SR - RR.u) * iSR;
  const double gyLL = (SB - LL.v) * iSB, gyRL = (SB - RL.v) * iSB, gyLR = (ST - LR.v) * iST, gyRR = (ST - RR.v) * iST;
  const double BstarLL = LL.b * gxLL, BstarLR = LR.b * gxLR, BstarRL = RL.b * gxRL, BstarRR = RR.b * gxRR;
  const double AstarLL = LL.a * gyLL, AstarLR = LR.a * gyLR, AstarRL = RL.a * gyRL, AstarRR = RR.a * gyRR;

  // squared Alfven speeds: a^2/rstar_x = a^2 /(r gx), Astar^2/rstar = (a^2/(r gx)) gy ; same for b with x<->y
  const double axLL = LL.a * LL.a * iLL * frcp(gxLL), axLR = LR.a * LR.a * iLR * frcp(gxLR);
  const double axRL = RL.a * RL.a * iRL * frcp(gxRL), axRR = RR.a * RR.a * iRR * frcp(gxRR);
  const double byLL = LL.b * LL.b * iLL * frcp(gyLL), byLR = LR.b * LR.b * iLR * frcp(gyLR);
  const double byRL = RL.b * RL.b * iRL * frcp(gyRL), byRR = RR.b * RR.b * iRR * frcp(gyRR);
  const double sc2 = smallc * smallc;
  const double calfvenL = fsqrt_pos(pmax5(axLR, axLR * gyLR, axLL, axLL * gyLL, sc2));
  const double calfvenR = fsqrt_pos(pmax5(axRR, axRR * gyRR, axRL, axRL * gyRL, sc2));
  const double calfvenB = fsqrt_pos(pmax5(byLL, byLL * gxLL, byRL, byRL * gxRL, sc2));
  const double calfvenT = fsqrt_pos(pmax5(byLR, byLR * gxLR, byRR, byRR * gxRR, sc2));

  const double SAL = clamp_hi0(ustar - calfvenL);
  const double SAR = clamp_lo0(ustar + calfvenR);
  const double SAB = clamp_hi0(vstar - calfvenB);
  const double SAT = clamp_lo0(vstar + calfvenT);

  const bool SB_pos = !signbit(SB), ST_pos = !signbit(ST), SL_pos = !signbit(SL), SR_pos = !signbit(SR);
  if (SB_pos) {
    if (SL_pos) return ELL;
    if (!SR_pos) return ERL;
    return (SAR * (ustar * BstarLL - LL.v * LL.a) - SAL * (ustar * BstarRL - RL.v * RL.a) + SAR * SAL * (RL.b - LL.b)) * frcp(SAR - SAL);
  }
  if (!ST_pos) {
    if (SL_pos) return ELR;
    if (!SR_pos) return ERR;
    return (SAR * (ustar * BstarLR - LR.v * LR.a) - SAL * (ustar * BstarRR - RR.v * RR.a) + SAR * SAL * (RR.b - LR.b)) * frcp(SAR - SAL);
  }
  if (SL_pos) return (SAT * (LL.u * LL.b - vstar * AstarLL) - SAB * (LR.u * LR.b - vstar * AstarLR) - SAT * SAB * (LR.a - LL.a)) * frcp(SAT - SAB);
  if (!SR_pos) return (SAT * (RL.u * RL.b - vstar * AstarRL) - SAB * (RR.u * RR.b - vstar * AstarRR) - SAT * SAB * (RR.a - RL.a)) * frcp(SAT - SAB);
  const double iA = frcp(SAR - SAL), iB = frcp(SAT - SAB);
  const double AstarT = (SAR * AstarRR - SAL * AstarLR) * iA;
  const double AstarB = (SAR * AstarRL - SAL * AstarLL) * iA;
  const double BstarR = (SAT * BstarRR - SAB * BstarRL) * iB;
  const double BstarL = (SAT * BstarLR - SAB * BstarLL) * iB;
  const double EstarLL = ustar * BstarLL - vstar * AstarLL, EstarLR = ustar * BstarLR - vstar * AstarLR;
  const double EstarRL = ustar * BstarRL - vstar * AstarRL, EstarRR = ustar * BstarRR - vstar * AstarRR;
  return (SAL * SAB * EstarRR - SAL * SAT * EstarRL - SAR * SAB * EstarLR + SAR * SAT * EstarLL) * iA * iB -
         SAT * SAB * iB * (AstarT - AstarB) + SAR * SAL * iA * (BstarR - BstarL);
}
#endif

// mag_riemann2d_hlld, RiemannSolvers_MHD.h:398-630
DEV double mag_riemann2d_hlld(double gamma0, double smallc, const Corner &LL, const Corner &RL, const Corner &LR,
                              const Corner &RR, double ELL, double ERL, double ELR, double ERR) {
  double c2, d2;
  fast_speed_common(gamma0, LL.r, LL.p, LL.a, LL.b, LL.c, c2, d2);
  const double cFastLLx = fast_speed_dir(c2, d2, LL.r, LL.a), cFastLLy = fast_speed_dir(c2, d2, LL.r, LL.b);
  fast_speed_common(gamma0, LR.r, LR.p, LR.a, LR.b, LR.c, c2, d2);
  const double cFastLRx = fast_speed_dir(c2, d2, LR.r, LR.a), cFastLRy = fast_speed_dir(c2, d2, LR.r, LR.b);
  fast_speed_common(gamma0, RL.r, RL.p, RL.a, RL.b, RL.c, c2, d2);
  const double cFastRLx = fast_speed_dir(c2, d2, RL.r, RL.a), cFastRLy = fast_speed_dir(c2, d2, RL.r, RL.b);
  fast_speed_common(gamma0, RR.r, RR.p, RR.a, RR.b, RR.c, c2, d2);
  const double cFastRRx = fast_speed_dir(c2, d2, RR.r, RR.a), cFastRRy = fast_speed_dir(c2, d2, RR.r, RR.b);

  const double cxmax = max4(cFastLLx, cFastLRx, cFastRLx, cFastRRx);
  const double cymax = max4(cFastLLy, cFastLRy, cFastRLy, cFastRRy);
  const double SL = min4(LL.u, LR.u, RL.u, RR.u) - cxmax;
  const double SR = max4(LL.u, LR.u, RL.u, RR.u) + cxmax;
  const double SB = min4(LL.v, LR.v, RL.v, RR.v) - cymax;
  const double ST = max4(LL.v, LR.v, RL.v, RR.v) + cymax;

  const double PtotLL = LL.p + 0.5 * (LL.a * LL.a + LL.b * LL.b + LL.c * LL.c);
  const double PtotLR = LR.p + 0.5 * (LR.a * LR.a + LR.b * LR.b + LR.c * LR.c);
  const double PtotRL = RL.p + 0.5 * (RL.a * RL.a + RL.b * RL.b + RL.c * RL.c);
  const double PtotRR = RR.p + 0.5 * (RR.a * RR.a + RR.b * RR.b + RR.c * RR.c);

  const double rcLLx = LL.r * (LL.u - SL), rcRLx = RL.r * (SR - RL.u), rcLRx = LR.r * (LR.u - SL), rcRRx = RR.r * (SR - RR.u);
  const double rcLLy = LL.r * (LL.v - SB), rcLRy = LR.r * (ST - LR.v), rcRLy = RL.r * (RL.v - SB), rcRRy = RR.r * (ST - RR.v);

  const double ustar = (rcLLx * LL.u + rcLRx * LR.u + rcRLx * RL.u + rcRRx * RR.u + (PtotLL - PtotRL + PtotLR - PtotRR)) /
                       (rcLLx + rcLRx + rcRLx + rcRRx);
  const double vstar = (rcLLy * LL.v + rcLRy * LR.v + rcRLy * RL.v + rcRRy * RR.v + (PtotLL - PtotLR + PtotRL - PtotRR)) /
                       (rcLLy + rcLRy + rcRLy + rcRRy);

  const double rstarLLx = LL.r * (SL - LL.u) / (SL - ustar);
  const double BstarLL = LL.b * (SL - LL.u) / (SL - ustar);
  const double rstarLLy = LL.r * (SB - LL.v) / (SB - vstar);
  const double AstarLL = LL.a * (SB - LL.v) / (SB - vstar);
  const double rstarLL = LL.r * (SL - LL.u) / (SL - ustar) * (SB - LL.v) / (SB - vstar);
  const double EstarLLx = ustar * BstarLL - LL.v * LL.a;
  const double EstarLLy = LL.u * LL.b - vstar * AstarLL;
  const double EstarLL = ustar * BstarLL - vstar * AstarLL;

  const double rstarLRx = LR.r * (SL - LR.u) / (SL - ustar);
  const double BstarLR = LR.b * (SL - LR.u) / (SL - ustar);
  const double rstarLRy = LR.r * (ST - LR.v) / (ST - vstar);
  const double AstarLR = LR.a * (ST - LR.v) / (ST - vstar);
  const double rstarLR = LR.r * (SL - LR.u) / (SL - ustar) * (ST - LR.v) / (ST - vstar);
  const double EstarLRx = ustar * BstarLR - LR.v * LR.a;
  const double EstarLRy = LR.u * LR.b - vstar * AstarLR;
  const double EstarLR = ustar * BstarLR - vstar * AstarLR;

  const double rstarRLx = RL.r * (SR - RL.u) / (SR - ustar);
  const double BstarRL = RL.b * (SR - RL.u) / (SR - ustar);
  const double rstarRLy = RL.r * (SB - RL.v) / (SB - vstar);
  const double AstarRL = RL.a * (SB - RL.v) / (SB - vstar);
  const double rstarRL = RL.r * (SR - RL.u) / (SR - ustar) * (SB - RL.v) / (SB - vstar);
  const double EstarRLx = ustar * BstarRL - RL.v * RL.a;
  const double EstarRLy = RL.u * RL.b - vstar * AstarRL;
  const double EstarRL = ustar * BstarRL - vstar * AstarRL;

  const double rstarRRx = RR.r * (SR - RR.u) / (SR - ustar);
  const double BstarRR = RR.b * (SR - RR.u) / (SR - ustar);
  const double rstarRRy = RR.r * (ST - RR.v) / (ST - vstar);
  const double AstarRR = RR.a * (ST - RR.v) / (ST - vstar);
  const double rstarRR = RR.r * (SR - RR.u) / (SR - ustar) * (ST - RR.v) / (ST - vstar);
  const double EstarRRx = ustar * BstarRR - RR.v * RR.a;
  const double EstarRRy = RR.u * RR.b - vstar * AstarRR;
  const double EstarRR = ustar * BstarRR - vstar * AstarRR;

  const double calfvenL = max5(fabs(LR.a) / sqrt(rstarLRx), fabs(AstarLR) / sqrt(rstarLR), fabs(LL.a) / sqrt(rstarLLx),
                               fabs(AstarLL) / sqrt(rstarLL), smallc);
  const double calfvenR = max5(fabs(RR.a) / sqrt(rstarRRx), fabs(AstarRR) / sqrt(rstarRR), fabs(RL.a) / sqrt(rstarRLx),
                               fabs(AstarRL) / sqrt(rstarRL), smallc);
  const double calfvenB = max5(fabs(LL.b) / sqrt(rstarLLy), fabs(BstarLL) / sqrt(rstarLL), fabs(RL.b) / sqrt(rstarRLy),
                               fabs(BstarRL) / sqrt(rstarRL), smallc);
  const double calfvenT = max5(fabs(LR.b) / sqrt(rstarLRy), fabs(BstarLR) / sqrt(rstarLR), fabs(RR.b) / sqrt(rstarRRy),
                               fabs(BstarRR) / sqrt(rstarRR), smallc);

  const double SAL = fmin(ustar - calfvenL, 0.0);
  const double SAR = fmax(ustar + calfvenR, 0.0);
  const double SAB = fmin(vstar - calfvenB, 0.0);
  const double SAT = fmax(vstar + calfvenT, 0.0);

#if PPK_EXACT
  // RiemannSolvers_MHD.h:555-596 evaluated as written: every branch, blended with 0/1 weights
  const double AstarT = (SAR * AstarRR - SAL * AstarLR) / (SAR - SAL);
  const double AstarB = (SAR * AstarRL - SAL * AstarLL) / (SAR - SAL);
  const double BstarR = (SAT * BstarRR - SAB * BstarRL) / (SAT - SAB);
  const double BstarL = (SAT * BstarLR - SAB * BstarLL) / (SAT - SAB);

  double E = 0, tmpE = 0;
  const int SB_pos = signbit(SB) ? 0 : 1, SB_neg = 1 - SB_pos;
  const int ST_pos = signbit(ST) ? 0 : 1, ST_neg = 1 - ST_pos;
  const int SL_pos = signbit(SL) ? 0 : 1, SL_neg = 1 - SL_pos;
  const int SR_pos = signbit(SR) ? 0 : 1, SR_neg = 1 - SR_pos;

  tmpE = (SAL * SAB * EstarRR - SAL * SAT * EstarRL - SAR * SAB * EstarLR + SAR * SAT * EstarLL) / (SAR - SAL) / (SAT - SAB) -
         SAT * SAB / (SAT - SAB) * (AstarT - AstarB) + SAR * SAL / (SAR - SAL) * (BstarR - BstarL);
  E += (double)(SB_neg * ST_pos * SL_neg * SR_pos) * tmpE;

  tmpE = (SAR * EstarLLx - SAL * EstarRLx + SAR * SAL * (RL.b - LL.b)) / (SAR - SAL);
  tmpE = (double)SL_pos * ELL + (double)(SL_neg * SR_neg) * ERL + (double)(SL_neg * SR_pos) * tmpE;
  E += (double)SB_pos * tmpE;

  tmpE = (SAR * EstarLRx - SAL * EstarRRx + SAR * SAL * (RR.b - LR.b)) / (SAR - SAL);
  tmpE = (double)SL_pos * ELR + (double)(SL_neg * SR_neg) * ERR + (double)(SL_neg * SR_pos) * tmpE;
  E += (double)(SB_neg * ST_neg) * tmpE;

  tmpE = (SAT * EstarLLy - SAB * EstarLRy - SAT * SAB * (LR.a - LL.a)) / (SAT - SAB);
  E += (double)(SB_neg * ST_pos * SL_pos) * tmpE;

  tmpE = (SAT * EstarRLy - SAB * EstarRRy - SAT * SAB * (RR.a - RL.a)) / (SAT - SAB);
  E += (double)(SB_neg * ST_pos * SL_neg * SR_neg) * tmpE;
  return E;
#else
  // same selection (sign bits, RiemannSolvers_MHD.h:569-572), only the chosen branch is evaluated
  const bool SB_pos = !signbit(SB), ST_pos = !signbit(ST), SL_pos = !signbit(SL), SR_pos = !signbit(SR);
  if (SB_pos) {
    if (SL_pos) return ELL;
    if (!SR_pos) return ERL;
    return (SAR * EstarLLx - SAL * EstarRLx + SAR * SAL * (RL.b - LL.b)) / (SAR - SAL);
  }
  if (!ST_pos) {
    if (SL_pos) return ELR;
    if (!SR_pos) return ERR;
    return (SAR * EstarLRx - SAL * EstarRRx + SAR * SAL * (RR.b - LR.b)) / (SAR - SAL);
  }
  if (SL_pos) return (SAT * EstarLLy - SAB * EstarLRy - SAT * SAB * (LR.a - LL.a)) / (SAT - SAB);
  if (!SR_pos) return (SAT * EstarRLy - SAB * EstarRRy - SAT * SAB * (RR.a - RL.a)) / (SAT - SAB);
  const double AstarT = (SAR * AstarRR - SAL * AstarLR) / (SAR - SAL);
  const double AstarB = (SAR * AstarRL - SAL * AstarLL) / (SAR - SAL);
  const double BstarR = (SAT * BstarRR - SAB * BstarRL) / (SAT - SAB);
  const double BstarL = (SAT * BstarLR - SAB * BstarLL) / (SAT - SAB);
  return (SAL * SAB * EstarRR - SAL * SAT * EstarRL - SAR * SAB * EstarLR + SAR * SAT * EstarLL) / (SAR - SAL) / (SAT - SAB) -
         SAT * SAB / (SAT - SAB) * (AstarT - AstarB) + SAR * SAL / (SAR - SAL) * (BstarR - BstarL);
#endif
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------

// Ghost fill of one direction (both faces), MakeBoundariesFunctor3D_MHD<face>
// (BoundariesFunctors.h:749-1053). Faces whose BC is BC_COPY belong to the halo exchange.
// For x and y only the planes k in [kb0, kb0+nkb) are filled (the whole array, or the planes around an
// in-flight z exchange).
template <int DIR>
__global__ void k_boundary(const GridParams g, double *__restrict__ U, const int kb0, const int nkb) {
  const int gw = g.gw;
  const int e0 = DIR == 0 ? g.jsize : g.isize;
  const int e1 = DIR == 2 ? g.jsize : nkb;
  const int n = DIR == 0 ? g.nx : (DIR == 1 ? g.ny : g.nz);
  const long long total = 2LL * gw * e0 * e1;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  // fastest index: for DIR==0 the ghost layer (only 3 wide), otherwise i
  int a, b, gl, hi;
  long long r = t;
  if (DIR == 0) { gl = (int)(r % gw); r /= gw; a = (int)(r % e0); r /= e0; b = (int)(r % e1); hi = (int)(r / e1); }
  else { a = (int)(r % e0); r /= e0; gl = (int)(r % gw); r /= gw; b = (int)(r % e1); hi = (int)(r / e1); }
  const int bc = g.bc[2 * DIR + hi];
  if (bc == BC_COPY) return;
  const int c = hi ? gl + n + gw : gl;
  int c0;
  if (bc == BC_DIRICHLET) c0 = hi ? 2 * n + 2 * gw - 1 - c : 2 * gw - 1 - c;
  else if (bc == BC_NEUMANN) c0 = hi ? n + gw - 1 : gw;
  else c0 = hi ? c - n : n + c;
  long long dst, src;
  if (DIR == 0) { dst = cidx(g, c, a, b + kb0); src = cidx(g, c0, a, b + kb0); }
  else if (DIR == 1) { dst = cidx(g, a, c, b + kb0); src = cidx(g, a, c0, b + kb0); }
  else { dst = cidx(g, a, b, c); src = cidx(g, a, b, c0); }
  const int vflip = IU + DIR, bflip = IA + DIR;
#pragma unroll
  for (int v = 0; v < NBVAR; ++v) {
    double val = U[src + v * g.ncell];
    if (bc == BC_DIRICHLET && (v == vflip || v == bflip)) val = val * -1.0;
    U[dst + v * g.ncell] = val;
  }
}

// x / y faces of a block-decomposed run: the gw layers next to a face are strided in memory, so they travel through a
// packed buffer with the shape of the reference's border buffers (CopyDataArray_To_BorderBuf / CopyBorderBuf_To_DataArray,
// mpiBorderUtils.h:184-330; borderBufSend_xmin_3d(gw, jsize, ksize, nbvar), SolverBase.cpp:77-95), ghosts of the other two
// directions included so that edges and corners propagate through the X -> Y -> Z order of make_boundaries.
//   DIR 0: buf[g + gw*(j + jsize*(k + ksize*v))] <-> U[c0+g, j, k, v]      DIR 1: buf[i + isize*(g + gw*(k + ksize*v))] <-> U[i, c0+g, k, v]
template <int DIR, bool PACK>
__global__ void __launch_bounds__(256) k_face_copy(const GridParams g, double *__restrict__ U, double *__restrict__ buf, const int c0) {
  const int gw = g.gw;
  const long long per_var = (long long)gw * (DIR == 0 ? g.jsize : g.isize) * g.ksize;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= per_var) return;
  int i, j, k;
  long long r = t;
  if (DIR == 0) { i = c0 + (int)(r % gw); r /= gw; j = (int)(r % g.jsize); k = (int)(r / g.jsize); }
  else { i = (int)(r % g.isize); r /= g.isize; j = c0 + (int)(r % gw); k = (int)(r / gw); }
  const long long c = cidx(g, i, j, k);
#pragma unroll
  for (int v = 0; v < NBVAR; ++v) {
    if (PACK) buf[t + v * per_var] = U[c + v * g.ncell];
    else U[c + v * g.ncell] = buf[t + v * per_var];
  }
}

DEV double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ConvertToPrimitivesFunctor3D_MHD (MHDRunFunctors3D.h:88-163, constoprim_mhd MHDBaseFunctor3D.h:192-242)
// fused with ComputeDtFunctor3D_MHD (MHDRunFunctors3D.h:16-83, find_speed_info<3> mhd_utils.h:319-366):
// Q is written on [0,size-1)^3 and the CFL max is reduced over interior cells with warp shuffles,
// one shared-memory stage and one 64-bit atomicMax per block (positive doubles order like their bits).
template <bool WRITEQ>
__global__ void __launch_bounds__(256) k_prim_dt(const GridParams g, const double *__restrict__ U, double *__restrict__ Q,
                                                 StepState *st, int k0) {
  const int k = k0 + blockIdx.y;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned j = t / (unsigned)g.isize;
  const unsigned i = t - j * (unsigned)g.isize;
  double inv = 0.0;
  if (j < (unsigned)(g.jsize - 1) && i < (unsigned)(g.isize - 1) && k < g.ksize - 1) {
    const long long c = cidx(g, i, j, k);
    const long long N = g.ncell;
    const double ur = U[c + ID * N], ue = U[c + IP * N];
    const double mu = U[c + IU * N], mv = U[c + IV * N], mw = U[c + IW * N];
    const double fa = U[c + IA * N], fb = U[c + IB * N], fc = U[c + IC * N];
    const double fa1 = U[c + 1 + IA * N];
    const double fb1 = U[c + g.isize + IB * N];
    const double fc1 = U[c + (long long)g.isize * g.jsize + IC * N];
    const double r = vmax(ur, g.smallr);
#if PPK_EXACT
    const double u = mu / r, v = mv / r, w = mw / r;
#else
    const double ir = frcp(r);
    const double u = mu * ir, v = mv * ir, w = mw * ir;
#endif
    const double A = 0.5 * (fa + fa1), B = 0.5 * (fb + fb1), C = 0.5 * (fc + fc1);
    const double eken = 0.5 * (u * u + v * v + w * w);
    const double emag = 0.5 * (A * A + B * B + C * C);
#if PPK_EXACT
    const double eint = (ue - emag) / r - eken;
#else
    const double eint = (ue - emag) * ir - eken;
#endif
    const double p = vmax((g.gamma0 - 1.0) * r * eint, r * g.smallp);
    if (WRITEQ) {
      Q[c + ID * N] = r; Q[c + IP * N] = p; Q[c + IU * N] = u; Q[c + IV * N] = v; Q[c + IW * N] = w;
      Q[c + IA * N] = A; Q[c + IB * N] = B; Q[c + IC * N] = C;
    }
    const int gw = g.gw;
    if ((int)i >= gw && (int)i < g.isize - gw && (int)j >= gw && (int)j < g.jsize - gw && k >= gw && k < g.ksize - gw) {
#if PPK_EXACT
      double c2, d2;
      fast_speed_common(g.gamma0, r, p, A, B, C, c2, d2);
      const double vx = fast_speed_dir(c2, d2, r, A) + fabs(u);
      const double vy = fast_speed_dir(c2, d2, r, B) + fabs(v);
      const double vz = fast_speed_dir(c2, d2, r, C) + fabs(w);
      inv = vx / g.dx + vy / g.dy + vz / g.dz;
#else
      const double c2 = g.gamma0 * p * ir, d2 = 0.5 * (2.0 * emag * ir + c2), dd = d2 * d2, kk = c2 * ir;
      const double vx = fsqrt_pos(d2 + fsqrt_clamped(dd - kk * A * A)) + fabs(u);
      const double vy = fsqrt_pos(d2 + fsqrt_clamped(dd - kk * B * B)) + fabs(v);
      const double vz = fsqrt_pos(d2 + fsqrt_clamped(dd - kk * C * C)) + fabs(w);
      inv = vx * g.idx + vy * g.idy + vz * g.idz;
#endif
    }
  }
  // NaN-safe like fmax(invDt, x) in the reference: fmax drops NaNs
  inv = warp_max(inv);
  __shared__ double smax[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smax[wid] = inv;
  __syncthreads();
  if (wid == 0) {
    inv = lane < (int)(blockDim.x >> 5) ? smax[lane] : 0.0;
    inv = warp_max(inv);
    if (lane == 0 && inv > 0.0) atomicMax(&st->inv_dt_bits, (unsigned long long)__double_as_longlong(inv));
  }
}

// compute_dt_local (SolverMHDMuscl.h:724-741) + the clamp of SolverBase::compute_dt (SolverBase.cpp:174-177)
__global__ void k_finalize_dt(const GridParams g, StepState *st) {
  const double inv = __longlong_as_double((long long)st->inv_dt_bits);
  double dt = g.cfl / inv;
  if (st->t + dt > st->t_end) dt = st->t_end - st->t;
  st->dt = dt;
  st->dtdx = dt / g.dx; st->dtdy = dt / g.dy; st->dtdz = dt / g.dz;
  st->inv_dt_bits = 0ull;
}
// ++m_iteration; m_t += m_dt (SolverBase.cpp:216-218)
__global__ void k_advance_time(StepState *st) {
  st->iteration += 1;
  st->t += st->dt;
}

// ComputeElecFieldFunctor3D (MHDRunFunctors3D.h:278-362) + ComputeMagSlopesFunctor3D (:441-538,
// slope_unsplit_mhd_3d MHDBaseFunctor3D.h:561-668) on [1,size-1)^3.
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_elec_dbf(const GridParams g, const double *__restrict__ U,
                                                  const double *__restrict__ Q, double *__restrict__ E,
                                                  double *__restrict__ DBF, const int jslab) {
  const int k = 1 + blockIdx.y;
  const unsigned ni = g.isize - 2;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned jj = t / ni;
  const int i = 1 + (int)(t - jj * ni), j = 1 + (int)(blockIdx.z * jslab + jj);  // y-slabs: see slab_rows()
  if (jj >= (unsigned)jslab || j >= g.jsize - 1) return;
  const long long N = g.ncell, sj = g.isize, sk = (long long)g.isize * g.jsize;
  const long long c = cidx(g, i, j, k);
  const double *Qu = Q + IU * N, *Qv = Q + IV * N, *Qw = Q + IW * N;
  const double *Ua = U + IA * N, *Ub = U + IB * N, *Uc = U + IC * N;
  const double a0 = Ua[c], b0 = Ub[c], c0 = Uc[c];
  double u, v, w, A, B, C;
  // Ex: average over (j-1,k-1),(j-1,k),(j,k-1),(j,k) in this order
  v = 0.25 * (Qv[c - sj - sk] + Qv[c - sj] + Qv[c - sk] + Qv[c]);
  w = 0.25 * (Qw[c - sj - sk] + Qw[c - sj] + Qw[c - sk] + Qw[c]);
  B = 0.5 * (Ub[c - sk] + b0);
  C = 0.5 * (Uc[c - sj] + c0);
  E[c + 0 * N] = v * C - w * B;
  // Ey: (i-1,k-1),(i-1,k),(i,k-1),(i,k)
  u = 0.25 * (Qu[c - 1 - sk] + Qu[c - 1] + Qu[c - sk] + Qu[c]);
  w = 0.25 * (Qw[c - 1 - sk] + Qw[c - 1] + Qw[c - sk] + Qw[c]);
  A = 0.5 * (Ua[c - sk] + a0);
  C = 0.5 * (Uc[c - 1] + c0);
  E[c + 1 * N] = w * A - u * C;
  // Ez: (i-1,j-1),(i-1,j),(i,j-1),(i,j)
  u = 0.25 * (Qu[c - 1 - sj] + Qu[c - 1] + Qu[c - sj] + Qu[c]);
  v = 0.25 * (Qv[c - 1 - sj] + Qv[c - 1] + Qv[c - sj] + Qv[c]);
  A = 0.5 * (Ua[c - sj] + a0);
  B = 0.5 * (Ub[c - 1] + b0);
  E[c + 2 * N] = u * B - v * A;

  const double st = fmin(g.slope_type, 2.0);
  DBF[c + 0 * N] = limited_slope(st, a0, Ua[c + sj], Ua[c - sj]);  // dA/dy
  DBF[c + 1 * N] = limited_slope(st, a0, Ua[c + sk], Ua[c - sk]);  // dA/dz
  DBF[c + 2 * N] = limited_slope(st, b0, Ub[c + 1], Ub[c - 1]);    // dB/dx
  DBF[c + 3 * N] = limited_slope(st, b0, Ub[c + sk], Ub[c - sk]);  // dB/dz
  DBF[c + 4 * N] = limited_slope(st, c0, Uc[c + 1], Uc[c - 1]);    // dC/dx
  DBF[c + 5 * N] = limited_slope(st, c0, Uc[c + sj], Uc[c - sj]);  // dC/dy
}

// ComputeTraceFunctor3D_MHD (MHDRunFunctors3D.h:543-856): hydro slopes (slope_unsplit_hydro_3d,
// MHDBaseFunctor3D.h:362-495) + Hancock half step (trace_unsplit_mhd_3d_simpler, :688-896).
// Writes the 32-number basis on [2,size-2)^3 (the cells whose states a face or an edge consumes).
// (Measured alternative: accumulating the source terms direction by direction -- same summation order, 80-96 registers,
// 5-6 CTAs/SM, no spills -- is SLOWER, 1.59-1.76 ms against 1.42 ms: three dependent load phases with the slope stores
// between them lose more than the occupancy gains. Hoisting the face-field block (15 loads, 3 stores) to the top of the
// kernel, to remove the last dependent load phase the profile shows, is slower too (1.77 ms). One load burst at 128
// registers in this source order stays.)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_trace(const GridParams g, const StepState *__restrict__ stp,
                                                     const double *__restrict__ U, const double *__restrict__ Q,
                                                     const double *__restrict__ E, double *__restrict__ BASIS,
                                                     const int jslab) {
  const int k = 2 + blockIdx.y;
  const unsigned ni = g.isize - 4;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned jj = t / ni;
  const int i = 2 + (int)(t - jj * ni), j = 2 + (int)(blockIdx.z * jslab + jj);  // y-slabs: see slab_rows()
  if (jj >= (unsigned)jslab || j >= g.jsize - 2) return;
  const long long N = g.ncell, sj = g.isize, sk = (long long)g.isize * g.jsize;
  const long long c = cidx(g, i, j, k);
  const double dtdx = stp->dtdx, dtdy = stp->dtdy, dtdz = stp->dtdz;
  const double st = g.slope_type;
  const bool lim = (st == 1.0 || st == 2.0);

  double q[NBVAR], sx[NBVAR], sy[NBVAR], sz[NBVAR];
#pragma unroll
  for (int v = 0; v < NBVAR; ++v) {
    const double *Qv = Q + v * N;
    const double qc = Qv[c];
    q[v] = qc;
    // slopes are halved in place first (MHDBaseFunctor3D.h:767-813)
    sx[v] = lim ? 0.5 * limited_slope(st, qc, Qv[c + 1], Qv[c - 1]) : 0.0;
    sy[v] = lim ? 0.5 * limited_slope(st, qc, Qv[c + sj], Qv[c - sj]) : 0.0;
    sz[v] = lim ? 0.5 * limited_slope(st, qc, Qv[c + sk], Qv[c - sk]) : 0.0;
  }
  double r = q[ID], p = q[IP], u = q[IU], v = q[IV], w = q[IW], A = q[IA], B = q[IB], C = q[IC];
  const double drx = sx[ID], dpx = sx[IP], dux = sx[IU], dvx = sx[IV], dwx = sx[IW], dBx = sx[IB], dCx = sx[IC];
  const double dry = sy[ID], dpy = sy[IP], duy = sy[IU], dvy = sy[IV], dwy = sy[IW], dAy = sy[IA], dCy = sy[IC];
  const double drz = sz[ID], dpz = sz[IP], duz = sz[IU], dvz = sz[IV], dwz = sz[IW], dAz = sz[IA], dBz = sz[IB];

  double AL = U[c + IA * N], BL = U[c + IB * N], CL = U[c + IC * N];
  const double AR = U[c + 1 + IA * N], BR = U[c + sj + IB * N], CR = U[c + sk + IC * N];
  const double dAx = 0.5 * (AR - AL), dBy = 0.5 * (BR - BL), dCz = 0.5 * (CR - CL);

  const double gamma = g.gamma0;
  // source terms, MHDBaseFunctor3D.h:843-857
#if PPK_EXACT
#  define OVER_R(x) ((x) / r)
#else
  const double ir = frcp(r);
#  define OVER_R(x) ((x) * ir)
#endif
  const double sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy + (-w * drz - dwz * r) * dtdz;
  const double su0 = (-u * dux - OVER_R(dpx + B * dBx + C * dCx)) * dtdx + (-v * duy + OVER_R(B * dAy)) * dtdy +
                     (-w * duz + OVER_R(C * dAz)) * dtdz;
  const double sv0 = (-u * dvx + OVER_R(A * dBx)) * dtdx + (-v * dvy - OVER_R(dpy + A * dAy + C * dCy)) * dtdy +
                     (-w * dvz + OVER_R(C * dBz)) * dtdz;
  const double sw0 = (-u * dwx + OVER_R(A * dCx)) * dtdx + (-v * dwy + OVER_R(B * dCy)) * dtdy +
                     (-w * dwz - OVER_R(dpz + A * dAz + B * dBz)) * dtdz;
#undef OVER_R
  const double sp0 = (-u * dpx - dux * gamma * p) * dtdx + (-v * dpy - dvy * gamma * p) * dtdy +
                     (-w * dpz - dwz * gamma * p) * dtdz;
  const double sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy + (u * dCz + C * duz - w * dAz - A * dwz) * dtdz;
  const double sB0 = (v * dAx + A * dvx - u * dBx - B * dux) * dtdx + (v * dCz + C * dvz - w * dBz - B * dwz) * dtdz;
  const double sC0 = (w * dAx + A * dwx - u * dCx - C * dux) * dtdx + (w * dBy + B * dwy - v * dCy - C * dvy) * dtdy;

  // face-centred field from the edge electric field, :872-877
  const double *Ex = E, *Ey = E + N, *Ez = E + 2 * N;
  // Only the three LOWER face values are produced: the upper-face value of a cell (AR = U_A(c+1) + sAR0 etc.)
  // is the same expression on the same operands as the lower-face value of its +1 neighbour, hence the same bits,
  // and the face / edge kernels read it there.
  const double ELL = Ex[c], ELR = Ex[c + sk], ERL = Ex[c + sj];
  const double FLL = Ey[c], FLR = Ey[c + sk], FRL = Ey[c + 1];
  const double GLL = Ez[c], GLR = Ez[c + sj], GRL = Ez[c + 1];
  const double sAL0 = +(GLR - GLL) * dtdy * 0.5 - (FLR - FLL) * dtdz * 0.5;
  const double sBL0 = -(GRL - GLL) * dtdx * 0.5 + (ELR - ELL) * dtdz * 0.5;
  const double sCL0 = +(FRL - FLL) * dtdx * 0.5 - (ERL - ELL) * dtdy * 0.5;

  r = r + sr0; u = u + su0; v = v + sv0; w = w + sw0; p = p + sp0; A = A + sA0; B = B + sB0; C = C + sC0;
  AL = AL + sAL0; BL = BL + sBL0; CL = CL + sCL0;

  double *Bs = BASIS + c;
  ST(Bs[(BQ + ID) * N], r); ST(Bs[(BQ + IP) * N], p); ST(Bs[(BQ + IU) * N], u); ST(Bs[(BQ + IV) * N], v); ST(Bs[(BQ + IW) * N], w);
  ST(Bs[(BQ + IA) * N], A); ST(Bs[(BQ + IB) * N], B); ST(Bs[(BQ + IC) * N], C);
  ST(Bs[(BSX + 0) * N], drx); ST(Bs[(BSX + 1) * N], dpx); ST(Bs[(BSX + 2) * N], dux); ST(Bs[(BSX + 3) * N], dvx); ST(Bs[(BSX + 4) * N], dwx);
  ST(Bs[(BSX + 5) * N], dBx); ST(Bs[(BSX + 6) * N], dCx);
  ST(Bs[(BSY + 0) * N], dry); ST(Bs[(BSY + 1) * N], dpy); ST(Bs[(BSY + 2) * N], duy); ST(Bs[(BSY + 3) * N], dvy); ST(Bs[(BSY + 4) * N], dwy);
  ST(Bs[(BSY + 5) * N], dAy); ST(Bs[(BSY + 6) * N], dCy);
  ST(Bs[(BSZ + 0) * N], drz); ST(Bs[(BSZ + 1) * N], dpz); ST(Bs[(BSZ + 2) * N], duz); ST(Bs[(BSZ + 3) * N], dvz); ST(Bs[(BSZ + 4) * N], dwz);
  ST(Bs[(BSZ + 5) * N], dAz); ST(Bs[(BSZ + 6) * N], dBz);
  ST(Bs[(BFACE + 0) * N], AL); ST(Bs[(BFACE + 1) * N], BL); ST(Bs[(BFACE + 2) * N], CL);
}

// index, inside the 7 slopes of direction D, of field component m (m != D): r,p,u,v,w then the two
// transverse field components in increasing order
DEV constexpr int slope_b(int D, int m) { return 5 + (m > D ? m - 1 : m); }
DEV constexpr int slope_base(int D) { return D == 0 ? BSX : (D == 1 ? BSY : BSZ); }

// ComputeFluxesAndStoreFunctor3D_MHD (MHDRunFunctors3D.h:1783-1909), one face.
// Face cR = lower D-face of cell cR: left state = qm_D of cell cR - e_D, right state = qp_D of
// the cell (MHDBaseFunctor3D.h:898-968), rotated into the face frame exactly like the swapValues calls
// of the reference (y: u<->v, A<->B ; z: u<->w, A<->C). Produces (rho, E, normal, t1, t2) fluxes.
template <int D>
DEV void flux_face(const GridParams &g, const double *__restrict__ BASIS, long long cR, double &fd, double &fp,
                   double &fu, double &fv, double &fw) {
  const long long N = g.ncell;
  const long long sD = D == 0 ? 1 : (D == 1 ? (long long)g.isize : (long long)g.isize * g.jsize);
  const long long cL = cR - sD;
  constexpr int T1 = D == 0 ? 1 : (D == 1 ? 0 : 1);  // frame after the reference's swaps
  constexpr int T2 = D == 0 ? 2 : (D == 1 ? 2 : 0);
  constexpr int SB = slope_base(D);
  const double *BL_ = BASIS + cL, *BR_ = BASIS + cR;

  // left state: q + slope (qm), normal field = upper-face value of the left cell
  const double rl = vmax(g.smallr, BL_[(BQ + ID) * N] + BL_[(SB + 0) * N]);
  const double pl = vmax(g.smallp, BL_[(BQ + IP) * N] + BL_[(SB + 1) * N]);
  const double unl = BL_[(BQ + IU + D) * N] + BL_[(SB + 2 + D) * N];
  const double t1l = BL_[(BQ + IU + T1) * N] + BL_[(SB + 2 + T1) * N];
  const double t2l = BL_[(BQ + IU + T2) * N] + BL_[(SB + 2 + T2) * N];
  const double bnl = BR_[(BFACE + D) * N];  // upper face of the left cell == lower face of the right cell
  const double b1l = BL_[(BQ + IA + T1) * N] + BL_[(SB + slope_b(D, T1)) * N];
  const double b2l = BL_[(BQ + IA + T2) * N] + BL_[(SB + slope_b(D, T2)) * N];
  // right state: q - slope (qp), normal field = lower-face value of the right cell
  const double rr = vmax(g.smallr, BR_[(BQ + ID) * N] - BR_[(SB + 0) * N]);
  const double pr = vmax(g.smallp, BR_[(BQ + IP) * N] - BR_[(SB + 1) * N]);
  const double unr = BR_[(BQ + IU + D) * N] - BR_[(SB + 2 + D) * N];
  const double t1r = BR_[(BQ + IU + T1) * N] - BR_[(SB + 2 + T1) * N];
  const double t2r = BR_[(BQ + IU + T2) * N] - BR_[(SB + 2 + T2) * N];
  const double bnr = BR_[(BFACE + D) * N];
  const double b1r = BR_[(BQ + IA + T1) * N] - BR_[(SB + slope_b(D, T1)) * N];
  const double b2r = BR_[(BQ + IA + T2) * N] - BR_[(SB + slope_b(D, T2)) * N];

  riemann_face(g, rl, pl, unl, t1l, t2l, bnl, b1l, b2l, rr, pr, unr, t1r, t2r, bnr, b1r, b2r, fd, fp, fu, fv, fw);
}

// One direction per launch (the unfused pipeline). Only faces the update reads are computed:
// normal index in [gw, n+gw], transverse indices interior.
template <int D, int MINB>
__global__ void __launch_bounds__(128, MINB) k_flux(const GridParams g, const double *__restrict__ BASIS, double *__restrict__ F,
                                                    int ib, unsigned ni) {
  // faces i in [gw+ib, gw+ib+ni): the whole range, or the columns left over by the TMA-tiled kernel
  const int gw = g.gw;
  const int k = gw + blockIdx.y;
  const unsigned nj = g.ny + (D == 1 ? 1 : 0);
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned jj = t / ni;
  if (jj >= nj) return;
  const int i = gw + ib + (int)(t - jj * ni), j = gw + (int)jj;
  const long long N = g.ncell;
  const long long cR = cidx(g, i, j, k);
  double fd, fp, fu, fv, fw;
  flux_face<D>(g, BASIS, cR, fd, fp, fu, fv, fw);
  double *Fo = F + cR;
  Fo[0 * N] = fd; Fo[1 * N] = fp; Fo[2 * N] = fu; Fo[3 * N] = fv; Fo[4 * N] = fw;
}

// index of the limited slope of face component m along direction a (a != m) in DBF
DEV constexpr int dbf_idx(int m, int a) { return 2 * m + (a > m ? a - 1 : a); }

// One of the 12 edge states of trace_unsplit_mhd_3d_simpler (MHDBaseFunctor3D.h:970-1112) in the
// (d1,d2,e) frame of edge direction E, with signs (s1,s2) = position of the edge relative to the cell
// centre along d1,d2 (RT:++, RB:+-, LT:-+, LB:--).
template <int E>
DEV Corner edge_state(const GridParams &g, const double *__restrict__ BASIS, const double *__restrict__ DBF,
                      long long c, const bool s1p, const bool s2p) {
  constexpr int D1 = (E + 1) % 3, D2 = (E + 2) % 3;
  constexpr int S1 = slope_base(D1), S2 = slope_base(D2);
  const long long N = g.ncell;
  const long long st1 = D1 == 0 ? 1 : (D1 == 1 ? (long long)g.isize : (long long)g.isize * g.jsize);
  const long long st2 = D2 == 0 ? 1 : (D2 == 1 ? (long long)g.isize : (long long)g.isize * g.jsize);
  const double *Bc = BASIS + c;
  auto comb = [&](int qi, int i1, int i2) {
    const double a = Bc[(S1 + i1) * N], b = Bc[(S2 + i2) * N];
    return Bc[qi * N] + ((s1p ? a : -a) + (s2p ? b : -b));
  };
  Corner o;
  o.r = vmax(g.smallr, comb(BQ + ID, 0, 0));
  o.p = vmax(g.smallp, comb(BQ + IP, 1, 1));
  o.u = comb(BQ + IU + D1, 2 + D1, 2 + D1);
  o.v = comb(BQ + IU + D2, 2 + D2, 2 + D2);
  // field component normal to d1: face value on side s1, plus/minus half its limited slope along d2
  {
    const double face = Bc[(s1p ? st1 : 0) + (BFACE + D1) * N];  // upper face = lower face of the +d1 neighbour
    const double h = 0.5 * DBF[(s1p ? c + st1 : c) + dbf_idx(D1, D2) * N];
    o.a = face + (s2p ? h : -h);
  }
  {
    const double face = Bc[(s2p ? st2 : 0) + (BFACE + D2) * N];
    const double h = 0.5 * DBF[(s2p ? c + st2 : c) + dbf_idx(D2, D1) * N];
    o.b = face + (s1p ? h : -h);
  }
  o.c = comb(BQ + IA + E, slope_b(D1, E), slope_b(D2, E));
  return o;
}

// compute_emf<dir> (RiemannSolvers_MHD.h:651-874) from the four edge states around an edge:
// LL<-RT, RL<-LT, LR<-RB, RR<-LB ; in-plane field averaged across the edge
DEV double emf_from_corners(const GridParams &g, const Corner &RT, const Corner &RB, const Corner &LT, const Corner &LB) {
  Corner LL = RT, RL = LT, LR = RB, RR = LB;
  const double a_top = 0.5 * (RT.a + LT.a), a_bot = 0.5 * (RB.a + LB.a);
  const double b_rgt = 0.5 * (RT.b + RB.b), b_lft = 0.5 * (LT.b + LB.b);
  LL.a = a_top; RL.a = a_top; LR.a = a_bot; RR.a = a_bot;
  LL.b = b_rgt; LR.b = b_rgt; RL.b = b_lft; RR.b = b_lft;
  const double ELL = LL.u * LL.b - LL.v * LL.a;
  const double ERL = RL.u * RL.b - RL.v * RL.a;
  const double ELR = LR.u * LR.b - LR.v * LR.a;
  const double ERR = RR.u * RR.b - RR.v * RR.a;
#if PPK_EXACT
  return mag_riemann2d_hlld(g.gamma0, g.smallc, LL, RL, LR, RR, ELL, ERL, ELR, ERR);
#else
  return mag_riemann2d_hlld_fast(g.gamma0, g.smallc, LL, RL, LR, RR, ELL, ERL, ELR, ERR);
#endif
}


// ComputeEmfAndStoreFunctor3D (MHDRunFunctors3D.h:2100-2238) + compute_emf<dir>
// (RiemannSolvers_MHD.h:651-874), one edge. With the cyclic frame
// (d1,d2) = Z:(x,y) X:(y,z) Y:(z,x) the reference's three cases (including its RB/LT swap for EMF_y)
// are one pattern: RT from c-e1-e2, RB from c-e1, LT from c-e2, LB from c.
template <int E>
DEV double emf_edge(const GridParams &g, const double *__restrict__ BASIS, const double *__restrict__ DBF, long long c) {
  constexpr int D1 = (E + 1) % 3, D2 = (E + 2) % 3;
  const long long st1 = D1 == 0 ? 1 : (D1 == 1 ? (long long)g.isize : (long long)g.isize * g.jsize);
  const long long st2 = D2 == 0 ? 1 : (D2 == 1 ? (long long)g.isize : (long long)g.isize * g.jsize);

  const Corner RT = edge_state<E>(g, BASIS, DBF, c - st1 - st2, true, true);
  const Corner RB = edge_state<E>(g, BASIS, DBF, c - st1, true, false);
  const Corner LT = edge_state<E>(g, BASIS, DBF, c - st2, false, true);
  const Corner LB = edge_state<E>(g, BASIS, DBF, c, false, false);
  return emf_from_corners(g, RT, RB, LT, LB);
}

// One edge direction per launch (the unfused pipeline).
template <int E, int MINB>
__global__ void __launch_bounds__(128, MINB) k_emf(const GridParams g, const double *__restrict__ BASIS,
                                                   const double *__restrict__ DBF, double *__restrict__ EMF, int ib,
                                                   unsigned ni) {
  const int gw = g.gw;
  const int k = gw + blockIdx.y;
  const unsigned nj = g.ny + (E == 1 ? 0 : 1);
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned jj = t / ni;
  if (jj >= nj) return;
  const int i = gw + ib + (int)(t - jj * ni), j = gw + (int)jj;
  const long long c = cidx(g, i, j, k);
  EMF[c + (2 - E) * g.ncell] = emf_edge<E>(g, BASIS, DBF, c);
}

// ---------------------------------------------------------------------------------------------
// TMA-staged variants of the flux / EMF kernels.
//
// The basis is a 4-D tensor (x, y, z, component) in the reference's SoA layout, so the cells one CTA needs
// of one component are a 3-D box: one `cp.async.bulk.tensor.4d` (SASS UTMALDG) per component, issued by a
// single thread, lands the box in shared memory and signals an mbarrier. The Riemann problems then read
// their 2 (face) or 4 (edge) cells from shared memory at compile-time offsets: no per-load 64-bit address
// arithmetic, no long-scoreboard stall in the middle of the FP64 chains, and the loads of the CTA that is
// waiting overlap the arithmetic of the other CTAs resident on the SM.
// Needs an even isize (16-byte global strides); otherwise, and for the x-columns beyond the last full
// 32-wide tile, the plain kernels above are used.
// ---------------------------------------------------------------------------------------------
DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
DEV void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DEV void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DEV void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
    "{\n"
    ".reg .pred P1;\n"
    "LAB_WAIT:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
    "@P1 bra DONE;\n"
    "bra LAB_WAIT;\n"
    "DONE:\n"
    "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
DEV void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DEV void tma_load_box(double *dst, const CUtensorMap *map, int x, int y, int z, int comp, unsigned long long *bar) {
  asm volatile(
    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
      smem_u32(dst)),
    "l"(map), "r"(x), "r"(y), "r"(z), "r"(comp), "r"(smem_u32(bar))
    : "memory");
}

// tile geometry of one TMA-staged kernel: TX x TY x TZ faces/edges per CTA, HX/HY/HZ lower halo cells.
// A box must start on a 16-byte boundary of the innermost dimension, i.e. at an EVEN x index for fp64: tiles
// start at i0 = gw + 32*bx (odd, gw = 3), so every configuration takes HX = 1 (box from i0-1, 34 wide) whether
// or not its stencil reaches into x.
template <int TY_, int TZ_, int HX_, int HY_, int HZ_, int NSLOT_, int MINB_>
struct TileCfg {
  static constexpr int TX = 32, TY = TY_, TZ = TZ_, HX = HX_, HY = HY_, HZ = HZ_, NSLOT = NSLOT_;
  static constexpr int MINB = MINB_;  // CTAs per SM the register allocation aims at (measured best per kernel)
  static constexpr int XB = TX + 2 * HX;  // x extent padded to a 16-byte multiple (one unused column when HX = 1)
  static constexpr int YB = TY + HY, ZB = TZ + HZ;
  static constexpr int BOX = XB * YB * ZB;                          // doubles per component box
  static constexpr int SLOT = (BOX * 8 + 127) / 128 * 16;           // slot stride in doubles (128-byte aligned)
  static constexpr int SMEM_BYTES = NSLOT * SLOT * 8;
  static constexpr int THREADS = TX * TY * TZ;
};
// Measured on B200 at 256^3: the staged EMF kernels gain from more, smaller resident CTAs (z-edges: 128-thread CTAs x5
// at 96 registers 0.73 ms vs 0.92 ms for 256 x2). The tiles with a z halo take ONE plane of edges per CTA (box of 2 planes):
// 32x4x2 tiles needed 62-78 KB and fitted 2-3 times per SM; 32x4x1 fits 4-5 times although it re-reads the lower plane
// (x-edges 1.00 -> 0.80 ms, y-edges 0.77 -> 0.69 ms, A/B on one box).
template <int E> struct EmfCfg;  // edge direction E: halo of one cell along d1 and d2
template <> struct EmfCfg<2> : TileCfg<4, 1, 1, 1, 0, 19, 5> {};
#ifndef PPK_EMFX_TY  // tile-shape experiments: -DPPK_EMFX_TY=.. -DPPK_EMFX_TZ=.. -DPPK_EMFX_MINB=.. (same for EMFY)
#  define PPK_EMFX_TY 4
#  define PPK_EMFX_TZ 1
#  define PPK_EMFX_MINB 4
#endif
#ifndef PPK_EMFY_TY
#  define PPK_EMFY_TY 4
#  define PPK_EMFY_TZ 1
#  define PPK_EMFY_MINB 5
#endif
template <> struct EmfCfg<0> : TileCfg<PPK_EMFX_TY, PPK_EMFX_TZ, 1, 1, 1, 19, PPK_EMFX_MINB> {};
template <> struct EmfCfg<1> : TileCfg<PPK_EMFY_TY, PPK_EMFY_TZ, 1, 0, 1, 19, PPK_EMFY_MINB> {};
template <int D> struct FluxCfg;  // face direction D: halo of one cell along D
#ifndef PPK_FLUXXY_TY
#  define PPK_FLUXXY_TY 4  // 32x4 tiles, 5 CTAs/SM (<= 102 registers, no spills): 0.61 -> 0.55 ms against 32x8 tiles x3 (80 registers)
#  define PPK_FLUXXY_MINB 5
#endif
template <> struct FluxCfg<0> : TileCfg<PPK_FLUXXY_TY, 1, 1, 0, 0, 15, PPK_FLUXXY_MINB> {};
template <> struct FluxCfg<1> : TileCfg<PPK_FLUXXY_TY, 1, 1, 1, 0, 15, PPK_FLUXXY_MINB> {};
#ifndef PPK_FLUXZ_TY
#  define PPK_FLUXZ_TY 4
#  define PPK_FLUXZ_TZ 1
#  define PPK_FLUXZ_MINB 5
#endif
template <> struct FluxCfg<2> : TileCfg<PPK_FLUXZ_TY, PPK_FLUXZ_TZ, 1, 0, 1, 15, PPK_FLUXZ_MINB> {};

// component (row of the 4-D tensor) staged in slot s of the EMF kernel: 0-4 q, 5-9 slopes along d1, 10-14 slopes
// along d2, 15/16 lower-face field normal to d1/d2; slots 17,18 come from DBF
template <int E>
DEV constexpr int emf_slot_comp(int s) {
  constexpr int D1 = (E + 1) % 3, D2 = (E + 2) % 3;
  constexpr int S1 = slope_base(D1), S2 = slope_base(D2);
  const int q[5] = {BQ + ID, BQ + IP, BQ + IU + D1, BQ + IU + D2, BQ + IA + E};
  const int i1[5] = {0, 1, 2 + D1, 2 + D2, slope_b(D1, E)};
  const int i2[5] = {0, 1, 2 + D1, 2 + D2, slope_b(D2, E)};
  if (s < 5) return q[s];
  if (s < 10) return S1 + i1[s - 5];
  if (s < 15) return S2 + i2[s - 10];
  if (s < 17) return BFACE + (s == 15 ? D1 : D2);
  return s == 17 ? dbf_idx(D1, D2) : dbf_idx(D2, D1);
}

// edge_state<E> reading the staged tile; `o` = offset of the cell inside a component box
template <int E, class Cfg>
DEV Corner edge_state_smem(const GridParams &g, const double *__restrict__ sm, int o, const bool s1p, const bool s2p) {
  constexpr int D1 = (E + 1) % 3, D2 = (E + 2) % 3;
  constexpr int st[3] = {1, Cfg::XB, Cfg::XB * Cfg::YB};
  constexpr int S = Cfg::SLOT;
  auto comb = [&](int sq) {
    const double a = sm[(sq + 5) * S + o], b = sm[(sq + 10) * S + o];
    return sm[sq * S + o] + ((s1p ? a : -a) + (s2p ? b : -b));
  };
  Corner c;
  c.r = vmax(g.smallr, comb(0));
  c.p = vmax(g.smallp, comb(1));
  c.u = comb(2);
  c.v = comb(3);
  {
    const double face = sm[15 * S + (s1p ? o + st[D1] : o)];
    const double h = 0.5 * sm[17 * S + (s1p ? o + st[D1] : o)];
    c.a = face + (s2p ? h : -h);
  }
  {
    const double face = sm[16 * S + (s2p ? o + st[D2] : o)];
    const double h = 0.5 * sm[18 * S + (s2p ? o + st[D2] : o)];
    c.b = face + (s1p ? h : -h);
  }
  c.c = comb(4);
  return c;
}

// One tile of edges of direction E: stage the 19 component boxes, wait, solve, store. `bar` is a CTA-shared mbarrier.
template <int E, class Cfg = EmfCfg<E>>
DEV void emf_tile_body(const GridParams &g, const CUtensorMap *mapB, const CUtensorMap *mapD, double *__restrict__ EMF,
                       double *sm, unsigned long long *bar, const int i0, const int j0, const int k0) {
  constexpr int D1 = (E + 1) % 3, D2 = (E + 2) % 3;
  const int tid = threadIdx.x;
  const int tx = tid % Cfg::TX, ty = (tid / Cfg::TX) % Cfg::TY, tz = tid / (Cfg::TX * Cfg::TY);
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  if (tid >= Cfg::THREADS) return;  // (a CTA wider than the tile: k_riemann_all's x-edge task)
  if (tid == 0) {
    mbar_expect_tx(bar, Cfg::NSLOT * Cfg::BOX * 8);
#pragma unroll
    for (int s = 0; s < Cfg::NSLOT; ++s)
      tma_load_box(sm + s * Cfg::SLOT, s < 17 ? mapB : mapD, i0 - Cfg::HX, j0 - Cfg::HY, k0 - Cfg::HZ, emf_slot_comp<E>(s), bar);
  }
  mbar_wait(bar, 0);
  const int i = i0 + tx, j = j0 + ty, k = k0 + tz;
  // edges the CT update reads: index <= n+gw along d1 and d2, interior along the edge direction
  const int nj = g.ny + (E == 1 ? 0 : 1), nk = g.nz + (E == 2 ? 0 : 1);
  if (j >= g.gw + nj || k >= g.gw + nk) return;  // (the x range is tiled exactly by the launcher)
  constexpr int st[3] = {1, Cfg::XB, Cfg::XB * Cfg::YB};
  const int o = (tx + Cfg::HX) + Cfg::XB * ((ty + Cfg::HY) + Cfg::YB * (tz + Cfg::HZ));
  const Corner RT = edge_state_smem<E, Cfg>(g, sm, o - st[D1] - st[D2], true, true);
  const Corner RB = edge_state_smem<E, Cfg>(g, sm, o - st[D1], true, false);
  const Corner LT = edge_state_smem<E, Cfg>(g, sm, o - st[D2], false, true);
  const Corner LB = edge_state_smem<E, Cfg>(g, sm, o, false, false);
  ST(EMF[cidx(g, i, j, k) + (2 - E) * g.ncell], emf_from_corners(g, RT, RB, LT, LB));
}

template <int E, bool SLAB>
__global__ void __launch_bounds__(EmfCfg<E>::THREADS, EmfCfg<E>::MINB)
  k_emf_tma(const GridParams g, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapD,
            double *__restrict__ EMF, const int yslab, const unsigned ymagic) {
  using Cfg = EmfCfg<E>;
  extern __shared__ __align__(128) double sm[];
  __shared__ unsigned long long bar;
  // grid (x-tiles, y-tiles, z-tiles) when one y-slab covers the plane (!SLAB); otherwise x = x-tiles,
  // y = (z-tile, y-tile inside a slab of `yslab` tiles), z = slab -- see slab_rows()
  // (blockIdx.y / yslab by multiply-high: exact for 16-bit operands, keeps the prologue off the slow integer division)
  const unsigned ydiv = SLAB ? __umulhi(blockIdx.y, ymagic) : 0u;
  const int by = SLAB ? blockIdx.z * yslab + (int)(blockIdx.y - ydiv * yslab) : blockIdx.y;
  const int bz = SLAB ? (int)ydiv : blockIdx.z;
  const int i0 = g.gw + blockIdx.x * Cfg::TX, j0 = g.gw + by * Cfg::TY, k0 = g.gw + bz * Cfg::TZ;
  if (SLAB && j0 >= g.gw + g.ny + (E == 1 ? 0 : 1)) return;  // the last slab may be short
  emf_tile_body<E>(g, &mapB, &mapD, EMF, sm, &bar, i0, j0, k0);
}

// component staged in slot s of the flux kernel: 0-6 q (r,p,un,t1,t2,b1,b2), 7-13 their slopes along D,
// 14 lower-face normal field of the right cell (= the upper-face value of the left cell)
template <int D>
DEV constexpr int flux_slot_comp(int s) {
  constexpr int T1 = D == 0 ? 1 : (D == 1 ? 0 : 1), T2 = D == 0 ? 2 : (D == 1 ? 2 : 0);
  constexpr int SB = slope_base(D);
  const int q[7] = {BQ + ID, BQ + IP, BQ + IU + D, BQ + IU + T1, BQ + IU + T2, BQ + IA + T1, BQ + IA + T2};
  const int sl[7] = {0, 1, 2 + D, 2 + T1, 2 + T2, slope_b(D, T1), slope_b(D, T2)};
  if (s < 7) return q[s];
  if (s < 14) return SB + sl[s - 7];
  return BFACE + D;
}

// One tile of faces of direction D: stage the 15 component boxes, wait, solve, store.
template <int D, int RS>
DEV void flux_tile_body(const GridParams &g, const CUtensorMap *mapB, double *__restrict__ F, double *sm,
                        unsigned long long *bar, const int i0, const int j0, const int k0) {
  using Cfg = FluxCfg<D>;
  const int tid = threadIdx.x;
  const int tx = tid % Cfg::TX, ty = (tid / Cfg::TX) % Cfg::TY, tz = tid / (Cfg::TX * Cfg::TY);
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, Cfg::NSLOT * Cfg::BOX * 8);
#pragma unroll
    for (int s = 0; s < Cfg::NSLOT; ++s)
      tma_load_box(sm + s * Cfg::SLOT, mapB, i0 - Cfg::HX, j0 - Cfg::HY, k0 - Cfg::HZ, flux_slot_comp<D>(s), bar);
  }
  mbar_wait(bar, 0);
  const int i = i0 + tx, j = j0 + ty, k = k0 + tz;
  const int nj = g.ny + (D == 1 ? 1 : 0), nk = g.nz + (D == 2 ? 1 : 0);
  if (j >= g.gw + nj || k >= g.gw + nk) return;
  constexpr int st[3] = {1, Cfg::XB, Cfg::XB * Cfg::YB};
  constexpr int S = Cfg::SLOT;
  const int oR = (tx + Cfg::HX) + Cfg::XB * ((ty + Cfg::HY) + Cfg::YB * (tz + Cfg::HZ)), oL = oR - st[D];
  // left state: q + slope (qm) of the cell below the face; right state: q - slope (qp) of the cell above
  const double rl = vmax(g.smallr, sm[0 * S + oL] + sm[7 * S + oL]);
  const double pl = vmax(g.smallp, sm[1 * S + oL] + sm[8 * S + oL]);
  const double unl = sm[2 * S + oL] + sm[9 * S + oL];
  const double t1l = sm[3 * S + oL] + sm[10 * S + oL];
  const double t2l = sm[4 * S + oL] + sm[11 * S + oL];
  const double bnl = sm[14 * S + oR];
  const double b1l = sm[5 * S + oL] + sm[12 * S + oL];
  const double b2l = sm[6 * S + oL] + sm[13 * S + oL];
  const double rr = vmax(g.smallr, sm[0 * S + oR] - sm[7 * S + oR]);
  const double pr = vmax(g.smallp, sm[1 * S + oR] - sm[8 * S + oR]);
  const double unr = sm[2 * S + oR] - sm[9 * S + oR];
  const double t1r = sm[3 * S + oR] - sm[10 * S + oR];
  const double t2r = sm[4 * S + oR] - sm[11 * S + oR];
  const double bnr = sm[14 * S + oR];
  const double b1r = sm[5 * S + oR] - sm[12 * S + oR];
  const double b2r = sm[6 * S + oR] - sm[13 * S + oR];
  double fd, fp, fu, fv, fw;
  riemann_face<RS>(g, rl, pl, unl, t1l, t2l, bnl, b1l, b2l, rr, pr, unr, t1r, t2r, bnr, b1r, b2r, fd, fp, fu, fv, fw);
  const long long N = g.ncell;
  double *Fo = F + cidx(g, i, j, k);
  ST(Fo[0 * N], fd); ST(Fo[1 * N], fp); ST(Fo[2 * N], fu); ST(Fo[3 * N], fv); ST(Fo[4 * N], fw);
}

template <int D, bool SLAB, int RS>
__global__ void __launch_bounds__(FluxCfg<D>::THREADS, FluxCfg<D>::MINB)
  k_flux_tma(const GridParams g, const __grid_constant__ CUtensorMap mapB, double *__restrict__ F, const int yslab,
             const unsigned ymagic) {
  using Cfg = FluxCfg<D>;
  extern __shared__ __align__(128) double sm[];
  __shared__ unsigned long long bar;
  // grid (x-tiles, y-tiles, z-tiles) when one y-slab covers the plane (!SLAB); otherwise x = x-tiles,
  // y = (z-tile, y-tile inside a slab of `yslab` tiles), z = slab -- see slab_rows()
  // (blockIdx.y / yslab by multiply-high: exact for 16-bit operands, keeps the prologue off the slow integer division)
  const unsigned ydiv = SLAB ? __umulhi(blockIdx.y, ymagic) : 0u;
  const int by = SLAB ? blockIdx.z * yslab + (int)(blockIdx.y - ydiv * yslab) : blockIdx.y;
  const int bz = SLAB ? (int)ydiv : blockIdx.z;
  const int i0 = g.gw + blockIdx.x * Cfg::TX, j0 = g.gw + by * Cfg::TY, k0 = g.gw + bz * Cfg::TZ;
  if (SLAB && j0 >= g.gw + g.ny + (D == 1 ? 1 : 0)) return;  // the last slab may be short
  flux_tile_body<D, RS>(g, &mapB, F, sm, &bar, i0, j0, k0);
}

#include "mhd_pgroup.inc"
#include "mhd_xzgroup.inc"

// ---------------------------------------------------------------------------------------------
// All six Riemann tasks of the step in ONE launch, ordered for the L2: the grid is one-dimensional and CTAs are
// dispatched in the order (y-slab, z-plane, task, tile), so the 38 basis / slope numbers of a plane are fetched from
// HBM by the first task that touches them and come out of the 126 MB L2 for the five others (and for the z-halo of
// the next plane). Launched one kernel per task, the same boxes came from HBM 3.3 times (every launch sweeps an array
// far larger than the L2). CTAs that are resident together mostly run the same task (a task-plane is about one wave),
// which keeps the instruction cache warm. Every task uses 32 x 4 x 1 tiles: the bodies are those of k_flux_tma /
// k_emf_tma.
// ---------------------------------------------------------------------------------------------
struct RiemannPlan {
  int ntx, nrows, rows;  // x tiles; rows of faces / edges (ny+1); rows per y-slab (a multiple of 12)
  unsigned per_task4, per_task3, per_plane, per_slab;
  int merged;  // 1: x-faces + y-faces + z-edges of a tile as ONE work item on shared boxes (mhd_pgroup.inc): 4 items per plane, not 6
};
struct RiemannMaps {
  CUtensorMap fluxB[3], emfB[3], emfD[3];
};
// x-edges inside k_riemann_all: 32 x 3 tiles (their boxes span two planes and one extra row: 34 x 4 x 2 cells x 19
// components = 41 KB like the y-edge boxes), so that every task fits five CTAs per SM
struct EmfXCfg3 : TileCfg<3, 1, 1, 1, 1, 19, 5> {};
constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int RALL_SMEM = cmax(cmax(cmax(EmfXCfg3::SMEM_BYTES, EmfCfg<1>::SMEM_BYTES), cmax(EmfCfg<2>::SMEM_BYTES, FluxCfg<2>::SMEM_BYTES)), PGroup::SMEM_BYTES);
static_assert(EmfCfg<1>::THREADS == 128 && EmfCfg<2>::THREADS == 128 && FluxCfg<0>::THREADS == 128 && FluxCfg<1>::THREADS == 128 &&
                FluxCfg<2>::THREADS == 128 && EmfCfg<1>::TY == 4 && EmfCfg<2>::TY == 4 && FluxCfg<0>::TY == 4 && FluxCfg<1>::TY == 4 &&
                FluxCfg<2>::TY == 4 && EmfXCfg3::THREADS == 96,
              "k_riemann_all assumes 32x4x1 tiles (32x3x1 for the x-edges)");
static_assert(5 * (RALL_SMEM + 1024) <= 227 * 1024, "five CTAs of k_riemann_all per SM");

template <int RS>
__global__ void __launch_bounds__(128, 5)
  k_riemann_all(const GridParams g, const RiemannPlan pl, const __grid_constant__ RiemannMaps maps, double *__restrict__ F0,
                double *__restrict__ F1, double *__restrict__ F2, double *__restrict__ EMF) {
  extern __shared__ __align__(128) double sm[];
  __shared__ unsigned long long bar;
  unsigned r = blockIdx.x;
  const unsigned slab = r / pl.per_slab; r -= slab * pl.per_slab;
  const unsigned plane = r / pl.per_plane; r -= plane * pl.per_plane;
  unsigned task, th;  // th = tile height
  if (pl.merged) {  // items of a plane: plane group (task 6), z-faces (3), y-edges (4), x-edges (5)
    if (r < 3u * pl.per_task4) { const unsigned t = r / pl.per_task4; r -= t * pl.per_task4; task = t == 0 ? 6u : 2u + t; th = 4; }
    else { task = 5; r -= 3u * pl.per_task4; th = 3; }
  } else if (r < 5u * pl.per_task4) { task = r / pl.per_task4; r -= task * pl.per_task4; th = 4; }
  else { task = 5; r -= 5u * pl.per_task4; th = 3; }
  const unsigned ty = r / (unsigned)pl.ntx, bx = r - ty * (unsigned)pl.ntx;
  const int jrow = (int)(slab * pl.rows + ty * th);  // first row of the tile
  if (jrow >= pl.nrows || jrow >= (int)(slab + 1) * pl.rows) return;
  const int i0 = g.gw + (int)bx * 32, j0 = g.gw + jrow, k0 = g.gw + (int)plane;
  const bool top = (int)plane == g.nz;  // only the z-faces and the x- / y-edges exist on the plane above the last cells
  const bool last_row = jrow >= g.ny;   // only y-faces and z- / x-edges exist on the row above the last cells
  switch (task) {
    case 0: if (!top && !last_row) flux_tile_body<0, RS>(g, &maps.fluxB[0], F0, sm, &bar, i0, j0, k0); break;
    case 1: if (!top) flux_tile_body<1, RS>(g, &maps.fluxB[1], F1, sm, &bar, i0, j0, k0); break;
    case 2: if (!top) emf_tile_body<2>(g, &maps.emfB[2], &maps.emfD[2], EMF, sm, &bar, i0, j0, k0); break;
    case 3: if (!last_row) flux_tile_body<2, RS>(g, &maps.fluxB[2], F2, sm, &bar, i0, j0, k0); break;
    case 4: if (!last_row) emf_tile_body<1>(g, &maps.emfB[1], &maps.emfD[1], EMF, sm, &bar, i0, j0, k0); break;
    case 6: if (!top) pgroup_tile_body<RS>(g, &maps.fluxB[1], &maps.emfD[2], F0, F1, EMF, sm, &bar, i0, j0, k0, last_row ? 2 : 3); break;
    default: emf_tile_body<0, EmfXCfg3>(g, &maps.emfB[0], &maps.emfD[0], EMF, sm, &bar, i0, j0, k0); break;
  }
}

#include "mhd_rpers.inc"
#include "mhd_trace_tma.inc"

// Kokkos::deep_copy(data_out, data_in) (SolverMHDMuscl.cpp:477) + UpdateFunctor3D_MHD
// (MHDRunFunctors3D.h:2430-2544) + UpdateEmfFunctor3D (:2549-2628) in one pass over the array:
// ghost cells are copied, interior cells receive the 6 face fluxes (fixed order) and the CT update.
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_update(const GridParams g, const StepState *__restrict__ stp,
                                                const double *__restrict__ Uin, double *__restrict__ Uout,
                                                const double *__restrict__ Fx, const double *__restrict__ Fy,
                                                const double *__restrict__ Fz, const double *__restrict__ EMF, const int kb0,
                                                const int jslab) {
  const int k = kb0 + blockIdx.y;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned jj = t / (unsigned)g.isize;
  const int i = (int)(t - jj * (unsigned)g.isize), j = (int)(blockIdx.z * jslab + jj);  // y-slabs: see slab_rows()
  if (jj >= (unsigned)jslab || j >= g.jsize) return;
  const long long N = g.ncell, sj = g.isize, sk = (long long)g.isize * g.jsize;
  const long long c = cidx(g, i, j, k);
  double u[NBVAR];
#pragma unroll
  for (int v = 0; v < NBVAR; ++v) u[v] = Uin[c + v * N];
  const int gw = g.gw;
  if (i >= gw && i < g.isize - gw && j >= gw && j < g.jsize - gw && k >= gw && k < g.ksize - gw) {
    const double dtdx = stp->dtdx, dtdy = stp->dtdy, dtdz = stp->dtdz;
    // +x neighbour of the x-faces and of the y- / z-edges. With x periodic the values at i = nx+gw are bit-identical to
    // those at i = gw and the flux / EMF launchers skip that column (see launch_flux): the one thread per row that
    // needs it re-reads the wrapped column (a branch, so that every other thread keeps its immediate-offset loads).
    const double *Ez = EMF, *Ey = EMF + N, *Ex = EMF + 2 * N;
    double fxh[NFLUX], ezx = Ez[c + 1], eyx = Ey[c + 1];
#pragma unroll
    for (int v = 0; v < NFLUX; ++v) fxh[v] = Fx[c + 1 + v * N];
    if (g.wrap_x && i + 1 == gw + g.nx) {
      const long long cw = c + 1 - g.nx;
#pragma unroll
      for (int v = 0; v < NFLUX; ++v) fxh[v] = Fx[cw + v * N];
      ezx = Ez[cw]; eyx = Ey[cw];
    }
    // x faces: (rho,E,mx,my,mz) <- (0,1,2,3,4)
    u[ID] += Fx[c + 0 * N] * dtdx; u[IP] += Fx[c + 1 * N] * dtdx; u[IU] += Fx[c + 2 * N] * dtdx;
    u[IV] += Fx[c + 3 * N] * dtdx; u[IW] += Fx[c + 4 * N] * dtdx;
    u[ID] -= fxh[0] * dtdx; u[IP] -= fxh[1] * dtdx; u[IU] -= fxh[2] * dtdx;
    u[IV] -= fxh[3] * dtdx; u[IW] -= fxh[4] * dtdx;
    // y faces, stored in the rotated frame: normal=my (2), t1=mx (3), t2=mz (4)
    u[ID] += Fy[c + 0 * N] * dtdy; u[IP] += Fy[c + 1 * N] * dtdy; u[IU] += Fy[c + 3 * N] * dtdy;
    u[IV] += Fy[c + 2 * N] * dtdy; u[IW] += Fy[c + 4 * N] * dtdy;
    u[ID] -= Fy[c + sj + 0 * N] * dtdy; u[IP] -= Fy[c + sj + 1 * N] * dtdy; u[IU] -= Fy[c + sj + 3 * N] * dtdy;
    u[IV] -= Fy[c + sj + 2 * N] * dtdy; u[IW] -= Fy[c + sj + 4 * N] * dtdy;
    // z faces: normal=mz (2), t1=my (3), t2=mx (4)
    u[ID] += Fz[c + 0 * N] * dtdz; u[IP] += Fz[c + 1 * N] * dtdz; u[IU] += Fz[c + 4 * N] * dtdz;
    u[IV] += Fz[c + 3 * N] * dtdz; u[IW] += Fz[c + 2 * N] * dtdz;
    u[ID] -= Fz[c + sk + 0 * N] * dtdz; u[IP] -= Fz[c + sk + 1 * N] * dtdz; u[IU] -= Fz[c + sk + 4 * N] * dtdz;
    u[IV] -= Fz[c + sk + 3 * N] * dtdz; u[IW] -= Fz[c + sk + 2 * N] * dtdz;
    // constrained transport, exact expression order of MHDRunFunctors3D.h:2602-2616
    const double ez = Ez[c], ey = Ey[c], ex = Ex[c];
    u[IA] += (Ez[c + sj] - ez) * dtdy;
    u[IB] -= (ezx - ez) * dtdx;
    u[IA] -= (Ey[c + sk] - ey) * dtdz;
    u[IB] += (Ex[c + sk] - ex) * dtdz;
    u[IC] += (eyx - ey) * dtdx;
    u[IC] -= (Ex[c + sj] - ex) * dtdy;
  }
#pragma unroll
  for (int v = 0; v < NBVAR; ++v) ST(Uout[c + v * N], u[v]);
}

// ---------------------------------------------------------------------------------------------
// Fused consumer: ComputeFluxesAndStore (x,y,z) + ComputeEmfAndStore (z,y,x) + UpdateFunctor3D_MHD +
// UpdateEmfFunctor3D (MHDRunFunctors3D.h:1783-1909, 2100-2238, 2430-2544, 2549-2628) in ONE kernel that
// never writes a flux or an EMF to HBM.
//
// A CTA owns a (CBX-1) x (CBY-1) tile of (i,j) columns and marches upwards in k. Thread (tx,ty) solves the
// Riemann problems on the three LOWER faces and the three LOWER edges of cell (i,j,k); the upper-face /
// upper-edge values a cell needs come from its +x / +y neighbours in the tile through shared memory (the
// last row and column of threads exist only to provide them: tiles overlap by one), and from the same
// thread one plane later in z. The reference's update order per cell
//     x-lo, x-hi, y-lo, y-hi, z-lo | z-hi          and for CT   Ez-terms, in-plane C terms | Ey/Ex(k+1) terms
// is exactly "everything of plane k, then the z-hi terms when plane k+1 has been solved", so the partially
// updated cell waits in registers for one iteration and the result is bit-identical to the unfused order.
// Ghost cells of Uout are not written: every step begins by refilling all ghost layers of its input array.
// ---------------------------------------------------------------------------------------------
constexpr int CBX = 32, CBY = 8;

template <bool HYDRO, bool CT>
__global__ void __launch_bounds__(CBX *CBY, 2)
  k_consume(const GridParams g, const StepState *__restrict__ stp, const double *__restrict__ BASIS,
            const double *__restrict__ DBF, const double *__restrict__ Uin, double *__restrict__ Uout, int zchunk) {
  __shared__ double sF[2][NFLUX][CBY][CBX];  // x- and y-face fluxes of the current plane
  __shared__ double sE[3][CBY][CBX];         // Ez, Ey, Ex of the current plane
  const int gw = g.gw;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = gw + blockIdx.x * (CBX - 1) + tx;
  const int j = gw + blockIdx.y * (CBY - 1) + ty;
  const bool in_i = i <= gw + g.nx, in_j = j <= gw + g.ny;  // faces and edges exist up to index n+gw inclusive
  const bool cell_i = i < gw + g.nx, cell_j = j < gw + g.ny;
  const bool own = tx < CBX - 1 && ty < CBY - 1 && cell_i && cell_j;
  // the solves an owner of this tile consumes from this thread
  const bool do_fx = HYDRO && in_i && cell_j && ty < CBY - 1;
  const bool do_fy = HYDRO && cell_i && in_j && tx < CBX - 1;
  const bool do_fz = HYDRO && own;
  const bool do_ez = CT && in_i && in_j && !(tx == CBX - 1 && ty == CBY - 1);
  const bool do_ey = CT && in_i && cell_j && ty < CBY - 1;
  const bool do_ex = CT && cell_i && in_j && tx < CBX - 1;

  const long long N = g.ncell, sk = (long long)g.isize * g.jsize;
  const double dtdx = stp->dtdx, dtdy = stp->dtdy, dtdz = stp->dtdz;
  // this CTA marches over cells k0 .. k1-1 (blockIdx.z selects the z-chunk; chunks only exist to give the grid
  // enough CTAs) and additionally solves the z-coupled problems of plane k1 to close its last cells
  const int k0 = gw + blockIdx.z * zchunk;
  const int k1 = min(k0 + zchunk, gw + g.nz);
  long long c = cidx(g, i, j, k0);
  double u[NBVAR];
  double ey_prev = 0.0, ex_prev = 0.0;
#pragma unroll
  for (int v = 0; v < NBVAR; ++v) u[v] = 0.0;

  for (int k = k0; k <= k1; ++k, c += sk) {
    double fz[NFLUX] = {0, 0, 0, 0, 0}, ey = 0.0, ex = 0.0;
    if (do_fz) flux_face<2>(g, BASIS, c, fz[0], fz[1], fz[2], fz[3], fz[4]);
    if (do_ey) ey = emf_edge<1>(g, BASIS, DBF, c);
    if (do_ex) ex = emf_edge<0>(g, BASIS, DBF, c);
    if (own && k > k0) {  // cell k-1 receives its z-hi terms and is complete
      const long long cp = c - sk;
      if (HYDRO) {
        u[ID] -= fz[0] * dtdz; u[IP] -= fz[1] * dtdz; u[IU] -= fz[4] * dtdz; u[IV] -= fz[3] * dtdz; u[IW] -= fz[2] * dtdz;
        Uout[cp + ID * N] = u[ID]; Uout[cp + IP * N] = u[IP]; Uout[cp + IU * N] = u[IU];
        Uout[cp + IV * N] = u[IV]; Uout[cp + IW * N] = u[IW];
      }
      if (CT) {
        u[IA] -= (ey - ey_prev) * dtdz;
        u[IB] += (ex - ex_prev) * dtdz;
        Uout[cp + IA * N] = u[IA]; Uout[cp + IB * N] = u[IB]; Uout[cp + IC * N] = u[IC];
      }
    }
    if (k == k1) break;  // uniform: the top plane only closes the last cells

    double fx[NFLUX] = {0, 0, 0, 0, 0}, fy[NFLUX] = {0, 0, 0, 0, 0}, ez = 0.0;
    if (do_fx) flux_face<0>(g, BASIS, c, fx[0], fx[1], fx[2], fx[3], fx[4]);
    if (do_fy) flux_face<1>(g, BASIS, c, fy[0], fy[1], fy[2], fy[3], fy[4]);
    if (do_ez) ez = emf_edge<2>(g, BASIS, DBF, c);
    if (HYDRO) {
#pragma unroll
      for (int v = 0; v < NFLUX; ++v) { sF[0][v][ty][tx] = fx[v]; sF[1][v][ty][tx] = fy[v]; }
    }
    if (CT) { sE[0][ty][tx] = ez; sE[1][ty][tx] = ey; sE[2][ty][tx] = ex; }
    __syncthreads();
    if (own) {
      if (HYDRO) {
        u[ID] = Uin[c + ID * N]; u[IP] = Uin[c + IP * N]; u[IU] = Uin[c + IU * N]; u[IV] = Uin[c + IV * N]; u[IW] = Uin[c + IW * N];
        // x faces: (rho,E,mx,my,mz) <- (0,1,2,3,4)
        u[ID] += fx[0] * dtdx; u[IP] += fx[1] * dtdx; u[IU] += fx[2] * dtdx; u[IV] += fx[3] * dtdx; u[IW] += fx[4] * dtdx;
        u[ID] -= sF[0][0][ty][tx + 1] * dtdx; u[IP] -= sF[0][1][ty][tx + 1] * dtdx; u[IU] -= sF[0][2][ty][tx + 1] * dtdx;
        u[IV] -= sF[0][3][ty][tx + 1] * dtdx; u[IW] -= sF[0][4][ty][tx + 1] * dtdx;
        // y faces, rotated frame: normal=my (2), t1=mx (3), t2=mz (4)
        u[ID] += fy[0] * dtdy; u[IP] += fy[1] * dtdy; u[IU] += fy[3] * dtdy; u[IV] += fy[2] * dtdy; u[IW] += fy[4] * dtdy;
        u[ID] -= sF[1][0][ty + 1][tx] * dtdy; u[IP] -= sF[1][1][ty + 1][tx] * dtdy; u[IU] -= sF[1][3][ty + 1][tx] * dtdy;
        u[IV] -= sF[1][2][ty + 1][tx] * dtdy; u[IW] -= sF[1][4][ty + 1][tx] * dtdy;
        // z-lo face: normal=mz (2), t1=my (3), t2=mx (4)
        u[ID] += fz[0] * dtdz; u[IP] += fz[1] * dtdz; u[IU] += fz[4] * dtdz; u[IV] += fz[3] * dtdz; u[IW] += fz[2] * dtdz;
      }
      if (CT) {
        u[IA] = Uin[c + IA * N]; u[IB] = Uin[c + IB * N]; u[IC] = Uin[c + IC * N];
        // expression order of MHDRunFunctors3D.h:2602-2616; the two Ey/Ex(k+1) terms follow one plane later
        u[IA] += (sE[0][ty + 1][tx] - ez) * dtdy;
        u[IB] -= (sE[0][ty][tx + 1] - ez) * dtdx;
        u[IC] += (sE[1][ty][tx + 1] - ey) * dtdx;
        u[IC] -= (sE[2][ty + 1][tx] - ex) * dtdy;
        ey_prev = ey; ex_prev = ex;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Streamed hydro consumer (PPK_PIPELINE_STREAMED): ComputeFluxesAndStore x,y,z (MHDRunFunctors3D.h:1783-1909) +
// UpdateFunctor3D_MHD (:2430-2544) in one z-marching kernel whose threads exchange one-sided STATES, not basis
// cells.
//
// A CTA owns 32 x SHY columns of cells and marches upwards in k. Thread (lane, w) reads the basis of ITS cell only
// (coalesced, every basis number is read once), forms the six one-sided states q -/+ slope of that cell, and hands
//   qm_x to lane+1 (warp shuffle), qm_y to row w+1 (shared memory), qm_z to itself one plane later (shared memory
// slot of its own), then solves the Riemann problems of its three LOWER faces. The upper-face fluxes come back the
// same way (shuffle / shared memory / next plane). One extra warp solves the faces on the far side of the tile
// (row j0+SHY: 32 y-faces; column i0+32: SHY x-faces in 8 lanes) through the SAME call sites as the tile's warps,
// so a face solved by two neighbouring CTAs gets the same bits from both. The left neighbours' states (column i0-1,
// row j0-1) are formed by lane 0 / warp 0 from 14 extra loads. No flux reaches HBM, and the update keeps the
// reference's order x-lo, x-hi, y-lo, y-hi, z-lo | z-hi (the last term one plane later): bit-identical.
// Only the 5 hydro variables of interior cells are written (the field is advanced by k_update_ct; ghost cells of
// Uout are refilled by the next step's boundary kernels before anything reads them).
// ---------------------------------------------------------------------------------------------
constexpr int SHY = 8;
constexpr int SH_THREADS = (SHY + 1) * 32;

// one-sided states of a cell along D in the face frame (r, p, un, t1, t2, b1, b2): qm = q + slope is the LEFT state
// of the cell's upper D-face, qp = q - slope the RIGHT state of its lower D-face (MHDBaseFunctor3D.h:898-968; the
// frame is the one the reference's swapValues calls produce, see flux_face)
template <int D>
DEV void one_sided_states(const GridParams &g, const double *__restrict__ Bc, const long long N, double qm[7], double qp[7]) {
  constexpr int T1 = D == 0 ? 1 : (D == 1 ? 0 : 1), T2 = D == 0 ? 2 : (D == 1 ? 2 : 0);
  constexpr int SB = slope_base(D);
  constexpr int qi[7] = {BQ + ID, BQ + IP, BQ + IU + D, BQ + IU + T1, BQ + IU + T2, BQ + IA + T1, BQ + IA + T2};
  constexpr int si[7] = {0, 1, 2 + D, 2 + T1, 2 + T2, slope_b(D, T1), slope_b(D, T2)};
#pragma unroll
  for (int v = 0; v < 7; ++v) {
    const double q = Bc[qi[v] * N], s = Bc[(SB + si[v]) * N];
    qm[v] = q + s;
    qp[v] = q - s;
  }
  qm[0] = vmax(g.smallr, qm[0]); qm[1] = vmax(g.smallp, qm[1]);
  qp[0] = vmax(g.smallr, qp[0]); qp[1] = vmax(g.smallp, qp[1]);
}
// the same states from a register copy b[NBASIS] of the cell's basis (k_hydro loads a whole cell in one burst so that a
// plane costs one exposed memory round trip instead of one per phase)
template <int D>
DEV void one_sided_states_reg(const GridParams &g, const double b[NBASIS], double qm[7], double qp[7]) {
  constexpr int T1 = D == 0 ? 1 : (D == 1 ? 0 : 1), T2 = D == 0 ? 2 : (D == 1 ? 2 : 0);
  constexpr int SB = slope_base(D);
  constexpr int qi[7] = {BQ + ID, BQ + IP, BQ + IU + D, BQ + IU + T1, BQ + IU + T2, BQ + IA + T1, BQ + IA + T2};
  constexpr int si[7] = {0, 1, 2 + D, 2 + T1, 2 + T2, slope_b(D, T1), slope_b(D, T2)};
#pragma unroll
  for (int v = 0; v < 7; ++v) {
    qm[v] = b[qi[v]] + b[SB + si[v]];
    qp[v] = b[qi[v]] - b[SB + si[v]];
  }
  qm[0] = vmax(g.smallr, qm[0]); qm[1] = vmax(g.smallp, qm[1]);
  qp[0] = vmax(g.smallr, qp[0]); qp[1] = vmax(g.smallp, qp[1]);
}
// loads of the basis numbers direction D needs (q, slopes along D, lower D-face field)
template <int D>
DEV void load_basis_dir(const double *__restrict__ Bc, const long long N, double b[NBASIS]) {
  constexpr int T1 = D == 0 ? 1 : (D == 1 ? 0 : 1), T2 = D == 0 ? 2 : (D == 1 ? 2 : 0);
  constexpr int SB = slope_base(D);
  constexpr int qi[7] = {BQ + ID, BQ + IP, BQ + IU + D, BQ + IU + T1, BQ + IU + T2, BQ + IA + T1, BQ + IA + T2};
#pragma unroll
  for (int v = 0; v < 7; ++v) { b[qi[v]] = Bc[qi[v] * N]; b[SB + v] = Bc[(SB + v) * N]; }
  b[BFACE + D] = Bc[(BFACE + D) * N];
}
DEV void benign_basis(double b[NBASIS]) {
#pragma unroll
  for (int v = 0; v < NBASIS; ++v) b[v] = 0.0;
  b[BQ + ID] = 1.0; b[BQ + IP] = 1.0;
}
DEV void benign_state(double q[7]) {  // operands of a solve whose result nobody reads
  q[0] = 1.0; q[1] = 1.0; q[2] = 0.0; q[3] = 0.0; q[4] = 0.0; q[5] = 0.0; q[6] = 0.0;
}
DEV void hlld7(const GridParams &g, const double L[7], const double R[7], const double bn, double f[5]) {
  riemann_face(g, L[0], L[1], L[2], L[3], L[4], bn, L[5], L[6], R[0], R[1], R[2], R[3], R[4], bn, R[5], R[6], f[0], f[1], f[2], f[3], f[4]);
}

// shared-memory layout of k_hydro (doubles): per-thread slots are [component][thread] so that a warp's access is
// conflict-free
struct HydroSmem {
  double xL[7][SHY * 32], xR[7][SHY * 32], xbn[SHY * 32];  // operands of the thread's x-face (formed one phase earlier)
  double yR[7][SHY * 32], ybn[SHY * 32];                   // right state of its y-face
  double lz[7][SHY * 32];                                  // qm_z of its cell, one plane below
  double yrow[SHY + 1][7][32];                             // qm_y of rows j0-1 .. j0+SHY-1 (index = tile row + 1)
  double fy[SHY + 1][5][32];                               // y-face fluxes of rows j0 .. j0+SHY
  double xcol[SHY][7];                                     // qm_x of the last cell of each row (left state, far x-face)
  double fxx[SHY][5];                                      // far x-face fluxes
  unsigned long long fxx_bar;                              // mbarrier: one phase per plane, completed by the extra warp
};

template <int MAXREG>
__global__ void __launch_bounds__(SH_THREADS) __maxnreg__(MAXREG)
  k_hydro(const GridParams g, const StepState *__restrict__ stp, const double *__restrict__ BASIS,
          const double *__restrict__ Uin, double *__restrict__ Uout, const int zchunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HydroSmem &S = *reinterpret_cast<HydroSmem *>(smem_raw);

  const int gw = g.gw;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool main_warp = w < SHY;
  const int i0 = gw + blockIdx.x * 32, j0 = gw + blockIdx.y * SHY;
  const int ie = gw + g.nx, je = gw + g.ny;  // one past the last interior cell
  const long long N = g.ncell, sj = g.isize, sk = (long long)g.isize * g.jsize;
  const double dtdx = stp->dtdx, dtdy = stp->dtdy, dtdz = stp->dtdz;
  const int k0 = gw + blockIdx.z * zchunk;
  const int k1 = min(k0 + zchunk, gw + g.nz);

  // tile warps: cell (i, j); the extra warp: cell (i0+lane, j0+SHY) for the far y-faces, cell (i0+32, j0+lane) for
  // the far x-faces
  const int i = i0 + lane, j = main_warp ? j0 + w : j0 + SHY;
  const bool cell_ok = main_warp && i < ie && j < je;
  const bool ld_ok = main_warp && i <= ie && j <= je;  // cells whose lower x- or y-face a cell of the tile consumes
  const bool ey_ok = !main_warp && i < ie && j <= je;
  const bool ex_ok = !main_warp && lane < SHY && i0 + 32 <= ie && j0 + lane < je;
  long long c = cidx(g, i, j, k0);                               // tile warps: own cell; extra warp: far-row cell
  long long cx = cidx(g, i0 + 32, j0 + (lane & (SHY - 1)), k0);  // extra warp: far-column cell
  if (tid == 0) mbar_init(&S.fxx_bar, 1);  // (visible to the CTA after barrier (1) of the first plane)

  if (main_warp) {  // prologue: left state of the lowest z-face
    double qm[7], qp[7];
    if (cell_ok) one_sided_states<2>(g, BASIS + c - sk, N, qm, qp);
    else benign_state(qm);
#pragma unroll
    for (int v = 0; v < 7; ++v) S.lz[v][tid] = qm[v];
  }
  double u[5] = {0.0, 0.0, 0.0, 0.0, 0.0};

  for (int k = k0; k <= k1; ++k, c += sk, cx += sk) {
    const bool last = k == k1;  // the top plane only closes the cells below it
    double fz[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    double Rx[7], Ry[7], bnx = 0.0, bny = 0.0;  // extra warp: right states of its far faces (registers)
    // ---- phase A: every load of this plane in one burst, all one-sided states, operands parked in shared memory ----
    if (main_warp) {
      double Lz[7], Rz[7], bnz;
      {
        double b[NBASIS];
        if (last) {
          if (cell_ok) load_basis_dir<2>(BASIS + c, N, b); else benign_basis(b);
        } else if (ld_ok) {
#pragma unroll
          for (int v = 0; v < NBASIS; ++v) b[v] = BASIS[c + v * N];
        } else benign_basis(b);
        double qm[7], qp[7];
        if (!last) {
          // y: left state of the row above goes through shared memory, the own right state is parked
          one_sided_states_reg<1>(g, b, qm, qp);
#pragma unroll
          for (int v = 0; v < 7; ++v) { S.yrow[w + 1][v][lane] = qm[v]; S.yR[v][tid] = qp[v]; }
          S.ybn[tid] = b[BFACE + 1];
          // x: left state from lane-1 (lane 0: column i0-1, formed by the extra warp), own right state
          one_sided_states_reg<0>(g, b, qm, qp);
#pragma unroll
          for (int v = 0; v < 7; ++v) {
            const double up = __shfl_up_sync(0xffffffffu, qm[v], 1);
            if (lane != 0) S.xL[v][tid] = up;
            S.xR[v][tid] = qp[v];
            if (lane == 31) S.xcol[w][v] = qm[v];
          }
          S.xbn[tid] = b[BFACE + 0];
        }
        one_sided_states_reg<2>(g, b, qm, Rz);
        bnz = b[BFACE + 2];
#pragma unroll
        for (int v = 0; v < 7; ++v) { Lz[v] = S.lz[v][tid]; S.lz[v][tid] = qm[v]; }
      }
      // ---- z face (needs no neighbour: runs before the barrier) ----
      hlld7(g, Lz, Rz, bnz, fz);
      if (cell_ok && k > k0) {  // cell k-1 receives its z-hi term and is complete
        const long long cp = c - sk;
        u[0] -= fz[0] * dtdz; u[1] -= fz[1] * dtdz; u[2] -= fz[4] * dtdz; u[3] -= fz[3] * dtdz; u[4] -= fz[2] * dtdz;
        Uout[cp + ID * N] = u[0]; Uout[cp + IP * N] = u[1]; Uout[cp + IU * N] = u[2];
        Uout[cp + IV * N] = u[3]; Uout[cp + IW * N] = u[4];
      }
      if (cell_ok && !last) {
        u[0] = Uin[c + ID * N]; u[1] = Uin[c + IP * N]; u[2] = Uin[c + IU * N]; u[3] = Uin[c + IV * N]; u[4] = Uin[c + IW * N];
      }
    } else if (!last) {
      // the extra warp forms every state that comes from outside the tile: right states of the far row / far column
      // (kept in registers for its own solves), left states of row j0-1 and of column i0-1 (published)
      double b[NBASIS], qm[7], qp[7];
      if (ey_ok) load_basis_dir<1>(BASIS + c, N, b); else benign_basis(b);
      one_sided_states_reg<1>(g, b, qm, Ry);
      bny = b[BFACE + 1];
      if (i < ie) load_basis_dir<1>(BASIS + c - (SHY + 1) * sj, N, b); else benign_basis(b);  // row j0-1
      one_sided_states_reg<1>(g, b, qm, qp);
#pragma unroll
      for (int v = 0; v < 7; ++v) S.yrow[0][v][lane] = qm[v];
      // lanes 0..SHY-1: far column i0+32 (right states); lanes SHY..2SHY-1: column i0-1 (left states)
      const bool hx = lane >= SHY && lane < 2 * SHY && j0 + (lane - SHY) < je;
      if (ex_ok) load_basis_dir<0>(BASIS + cx, N, b);
      else if (hx) load_basis_dir<0>(BASIS + cx - 33, N, b);
      else benign_basis(b);
      one_sided_states_reg<0>(g, b, qm, Rx);
      bnx = b[BFACE + 0];
      if (lane >= SHY && lane < 2 * SHY) {
#pragma unroll
        for (int v = 0; v < 7; ++v) S.xL[v][(lane - SHY) * 32] = qm[v];
      }
    }
    if (last) break;  // uniform
    __syncthreads();  // (1) published states are visible

    double L[7], R[7], f[5], bn;
    // ---- x faces ----
    if (main_warp) {
#pragma unroll
      for (int v = 0; v < 7; ++v) { L[v] = S.xL[v][tid]; R[v] = S.xR[v][tid]; }
      bn = S.xbn[tid];
    } else {
#pragma unroll
      for (int v = 0; v < 7; ++v) { L[v] = S.xcol[lane & (SHY - 1)][v]; R[v] = Rx[v]; }
      bn = bnx;
    }
    hlld7(g, L, R, bn, f);
    if (main_warp) {
      double fh[5];
#pragma unroll
      for (int v = 0; v < 5; ++v) fh[v] = __shfl_down_sync(0xffffffffu, f[v], 1);
      if (lane == 31) {  // the far x-face comes from the extra warp, which solved it while this warp did its own
        mbar_wait(&S.fxx_bar, (unsigned)(k - k0) & 1u);
#pragma unroll
        for (int v = 0; v < 5; ++v) fh[v] = S.fxx[w][v];
      }
      // x faces: (rho,E,mx,my,mz) <- (0,1,2,3,4)
      u[0] += f[0] * dtdx; u[1] += f[1] * dtdx; u[2] += f[2] * dtdx; u[3] += f[3] * dtdx; u[4] += f[4] * dtdx;
      u[0] -= fh[0] * dtdx; u[1] -= fh[1] * dtdx; u[2] -= fh[2] * dtdx; u[3] -= fh[3] * dtdx; u[4] -= fh[4] * dtdx;
    } else {
      if (lane < SHY) {
#pragma unroll
        for (int v = 0; v < 5; ++v) S.fxx[lane][v] = f[v];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.fxx_bar);  // release: the fluxes above are visible to whoever observes the phase
    }
    // ---- y faces ----
    if (main_warp) {
#pragma unroll
      for (int v = 0; v < 7; ++v) { L[v] = S.yrow[w][v][lane]; R[v] = S.yR[v][tid]; }
      bn = S.ybn[tid];
    } else {
#pragma unroll
      for (int v = 0; v < 7; ++v) { L[v] = S.yrow[SHY][v][lane]; R[v] = Ry[v]; }
      bn = bny;
    }
    hlld7(g, L, R, bn, f);
#pragma unroll
    for (int v = 0; v < 5; ++v) S.fy[w][v][lane] = f[v];
    __syncthreads();  // (2) y-face fluxes are visible; the operand slots may be rewritten
    if (main_warp) {
      // y faces, rotated frame: normal=my (2), t1=mx (3), t2=mz (4)
      u[0] += f[0] * dtdy; u[1] += f[1] * dtdy; u[2] += f[3] * dtdy; u[3] += f[2] * dtdy; u[4] += f[4] * dtdy;
      u[0] -= S.fy[w + 1][0][lane] * dtdy; u[1] -= S.fy[w + 1][1][lane] * dtdy; u[2] -= S.fy[w + 1][3][lane] * dtdy;
      u[3] -= S.fy[w + 1][2][lane] * dtdy; u[4] -= S.fy[w + 1][4][lane] * dtdy;
      // z-lo face: normal=mz (2), t1=my (3), t2=mx (4)
      u[0] += fz[0] * dtdz; u[1] += fz[1] * dtdz; u[2] += fz[4] * dtdz; u[3] += fz[3] * dtdz; u[4] += fz[2] * dtdz;
    }
  }
}

// UpdateEmfFunctor3D (MHDRunFunctors3D.h:2549-2628) alone: the field components of interior cells, expression order
// of :2602-2616 (the streamed pipeline's hydro variables are written by k_hydro)
__global__ void __launch_bounds__(256) k_update_ct(const GridParams g, const StepState *__restrict__ stp,
                                                   const double *__restrict__ Uin, double *__restrict__ Uout,
                                                   const double *__restrict__ EMF) {
  const int gw = g.gw;
  const int k = gw + blockIdx.y;
  const unsigned ni = g.nx;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned jj = t / ni;
  if (jj >= (unsigned)g.ny) return;
  const int i = gw + (int)(t - jj * ni), j = gw + (int)jj;
  const long long N = g.ncell, sj = g.isize, sk = (long long)g.isize * g.jsize;
  const long long c = cidx(g, i, j, k);
  const double dtdx = stp->dtdx, dtdy = stp->dtdy, dtdz = stp->dtdz;
  double a = Uin[c + IA * N], b = Uin[c + IB * N], cc = Uin[c + IC * N];
  const double *Ez = EMF, *Ey = EMF + N, *Ex = EMF + 2 * N;
  const double ez = Ez[c], ey = Ey[c], ex = Ex[c];
  a += (Ez[c + sj] - ez) * dtdy;
  const long long cx = (g.wrap_x && i + 1 == gw + g.nx) ? c + 1 - g.nx : c + 1;  // see k_update
  b -= (Ez[cx] - ez) * dtdx;
  a -= (Ey[c + sk] - ey) * dtdz;
  b += (Ex[c + sk] - ex) * dtdz;
  cc += (Ey[cx] - ey) * dtdx;
  cc -= (Ex[c + sj] - ex) * dtdy;
  Uout[c + IA * N] = a; Uout[c + IB * N] = b; Uout[c + IC * N] = cc;
}

// Column i = nx+gw of a flux / EMF component <- column i = gw (x periodic): only ppk_mhd3d_debug_array needs it, the
// update reads the wrapped column directly.
__global__ void k_wrap_x_column(const GridParams g, double *__restrict__ A, const int ncomp) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long rows = (long long)g.jsize * g.ksize;
  if (t >= rows * ncomp) return;
  const int comp = (int)(t / rows);
  const long long r = t - comp * rows;
  const long long c = (long long)g.isize * r + comp * g.ncell;
  A[c + g.gw + g.nx] = A[c + g.gw];
}

// Diagnostics: per-variable sums over the interior and max |div B| (needs filled upper ghosts).
// Two-stage: per-block partials with warp shuffles, then double atomicAdd / ordered-bits atomicMax.
__global__ void __launch_bounds__(256) k_diagnostics(const GridParams g, const double *__restrict__ U, double *out9) {
  const int gw = g.gw;
  const int k = gw + blockIdx.y;
  const unsigned ni = g.nx;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned jj = t / ni;
  double s[NBVAR], m = 0.0;
#pragma unroll
  for (int v = 0; v < NBVAR; ++v) s[v] = 0.0;
  if (jj < (unsigned)g.ny) {
    const int i = gw + (int)(t - jj * ni), j = gw + (int)jj;
    const long long N = g.ncell, sj = g.isize, sk = (long long)g.isize * g.jsize;
    const long long c = cidx(g, i, j, k);
#pragma unroll
    for (int v = 0; v < NBVAR; ++v) s[v] = U[c + v * N];
    m = fabs((U[c + 1 + IA * N] - s[IA]) / g.dx + (U[c + sj + IB * N] - s[IB]) / g.dy + (U[c + sk + IC * N] - s[IC]) / g.dz);
  }
  __shared__ double red[8][NBVAR + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < NBVAR; ++v)
    for (int o = 16; o > 0; o >>= 1) s[v] += __shfl_xor_sync(0xffffffffu, s[v], o);
  m = warp_max(m);
  if (lane == 0) {
    for (int v = 0; v < NBVAR; ++v) red[wid][v] = s[v];
    red[wid][NBVAR] = m;
  }
  __syncthreads();
  if (threadIdx.x < NBVAR) {
    double a = 0.0;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) a += red[w2][threadIdx.x];
    atomicAdd(&out9[threadIdx.x], a);
  } else if (threadIdx.x == NBVAR) {
    double a = 0.0;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) a = fmax(a, red[w2][NBVAR]);
    atomicMax((unsigned long long *)&out9[NBVAR], (unsigned long long)__double_as_longlong(a));
  }
}

// the reciprocal / square-root primitives of this build against which tests/ compare IEEE results
__global__ void k_fastmath_selftest(int n, const double *x, double *rcp, double *sq, double *rsq) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
#if PPK_EXACT
  rcp[t] = 1.0 / x[t]; sq[t] = sqrt(x[t]); rsq[t] = 1.0 / sqrt(x[t]);
#else
  rcp[t] = frcp(x[t]); sq[t] = fsqrt(x[t]); rsq[t] = frsqrt(x[t]);
#endif
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

static void l_boundary(const GridParams &g, double *U, int dir, int k0, int k1, cudaStream_t s) {
  if (dir == 2) { k0 = 0; k1 = g.ksize; }
  if (k1 <= k0) return;
  const long long e0 = dir == 0 ? g.jsize : g.isize, e1 = dir == 2 ? g.jsize : k1 - k0;
  const long long total = 2LL * g.gw * e0 * e1;
  const int bs = 256;
  if (dir == 0) k_boundary<0><<<cdiv(total, bs), bs, 0, s>>>(g, U, k0, k1 - k0);
  else if (dir == 1) k_boundary<1><<<cdiv(total, bs), bs, 0, s>>>(g, U, k0, k1 - k0);
  else k_boundary<2><<<cdiv(total, bs), bs, 0, s>>>(g, U, 0, g.ksize);
}
static void l_face_copy(const GridParams &g, double *U, double *buf, int dir, int c0, int pack, cudaStream_t s) {
  const long long per_var = (long long)g.gw * (dir == 0 ? g.jsize : g.isize) * g.ksize;
  const int bs = 256;
  if (dir == 0 && pack) k_face_copy<0, true><<<cdiv(per_var, bs), bs, 0, s>>>(g, U, buf, c0);
  else if (dir == 0) k_face_copy<0, false><<<cdiv(per_var, bs), bs, 0, s>>>(g, U, buf, c0);
  else if (pack) k_face_copy<1, true><<<cdiv(per_var, bs), bs, 0, s>>>(g, U, buf, c0);
  else k_face_copy<1, false><<<cdiv(per_var, bs), bs, 0, s>>>(g, U, buf, c0);
}
static void l_prim_dt(const GridParams &g, const double *U, double *Q, StepState *st, int k0, int k1, cudaStream_t s) {
  if (k1 <= k0) return;
  const int bs = 256;
  dim3 grid(cdiv((long long)g.isize * g.jsize, bs), k1 - k0);
  k_prim_dt<true><<<grid, bs, 0, s>>>(g, U, Q, st, k0);
}
// the CFL reduction alone (planes k in [k0, k1); only interior cells contribute): the tiled pipeline rebuilds the
// primitives on chip and needs dt before its producer kernel starts
static void l_dt_only(const GridParams &g, const double *U, StepState *st, int k0, int k1, cudaStream_t s) {
  if (k0 < g.gw) k0 = g.gw;
  if (k1 > g.ksize - g.gw) k1 = g.ksize - g.gw;
  if (k1 <= k0) return;
  const int bs = 256;
  dim3 grid(cdiv((long long)g.isize * g.jsize, bs), k1 - k0);
  k_prim_dt<false><<<grid, bs, 0, s>>>(g, U, nullptr, st, k0);
}
// The plane-sweeping kernels (grid x = cells of a k-plane, y = k) reuse the planes k-1, k, k+1 of their inputs from L2 --
// as long as one k-plane of ALL their streams fits there. At 512^3 a plane of the trace kernel's 46 streams is 99 MB and
// every z-neighbour is fetched from HBM again, so the sweep is cut into y-slabs of `rows` rows (grid z; CTAs are
// dispatched x-fastest, then k, then slab) that keep ~24 MB per plane. At 256^3 one slab covers the plane.
static int slab_rows(const GridParams &g, int nstreams, int nrows) {
  static const int forced = getenv("PPK_SLAB_ROWS") ? atoi(getenv("PPK_SLAB_ROWS")) : 0;
  long long rows = forced > 0 ? forced : (24LL << 20) / ((long long)nstreams * g.isize * 8);
  if (rows >= nrows) return nrows;
  rows &= ~7LL;
  return rows < 8 ? 8 : (int)rows;
}
static void l_finalize_dt(const GridParams &g, StepState *st, cudaStream_t s) { k_finalize_dt<<<1, 1, 0, s>>>(g, st); }
static void l_advance_time(StepState *st, cudaStream_t s) { k_advance_time<<<1, 1, 0, s>>>(st); }
static void l_elec_dbf(const GridParams &g, const double *U, const double *Q, double *E, double *DBF, cudaStream_t s) {
  const int bs = 256;
  const int rows = slab_rows(g, 15, g.jsize - 2);
  dim3 grid(cdiv((long long)(g.isize - 2) * rows, bs), g.ksize - 2, cdiv(g.jsize - 2, rows));
  static const int minb = getenv("PPK_ELEC_MINB") ? atoi(getenv("PPK_ELEC_MINB")) : 1;  // 1: 80 registers (more loads in flight per thread): 0.50 -> 0.46 ms against the 48-register default; 40 / 32 registers are slower
  if (minb == 8) k_elec_dbf<8><<<grid, bs, 0, s>>>(g, U, Q, E, DBF, rows);
  else if (minb == 6) k_elec_dbf<6><<<grid, bs, 0, s>>>(g, U, Q, E, DBF, rows);
  else if (minb == 1) k_elec_dbf<1><<<grid, bs, 0, s>>>(g, U, Q, E, DBF, rows);
  else k_elec_dbf<0><<<grid, bs, 0, s>>>(g, U, Q, E, DBF, rows);
}
static void l_trace_plain(const GridParams &g, const StepState *st, const double *U, const double *Q, const double *E,
                          double *BASIS, cudaStream_t s) {
  const int bs = 128;
  const int rows = slab_rows(g, 46, g.jsize - 4);
  dim3 grid(cdiv((long long)(g.isize - 4) * rows, bs), g.ksize - 4, cdiv(g.jsize - 4, rows));
  static const int minb = getenv("PPK_TRACE_MINB") ? atoi(getenv("PPK_TRACE_MINB")) : 4;
  if (minb == 5) k_trace<5><<<grid, bs, 0, s>>>(g, st, U, Q, E, BASIS, rows);
  else if (minb == 6) k_trace<6><<<grid, bs, 0, s>>>(g, st, U, Q, E, BASIS, rows);
  else if (minb == 8) k_trace<8><<<grid, bs, 0, s>>>(g, st, U, Q, E, BASIS, rows);
  else if (minb == 3) k_trace<3><<<grid, bs, 0, s>>>(g, st, U, Q, E, BASIS, rows);
  else k_trace<4><<<grid, bs, 0, s>>>(g, st, U, Q, E, BASIS, rows);
}
// ---- TMA tensor maps (host) ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct TmaCtx {
  CUtensorMap emfB[3], emfD[3], fluxB[3];
  CUtensorMap xzB, xzD;  // 34 x 1 x 5 boxes of the basis / the face-field slopes (mhd_xzgroup.inc)
  bool xz_ok = false;
  RiemannMaps rall;  // the same nine maps, as the single kernel parameter of k_riemann_all
  unsigned *counter = nullptr;  // work-item counter of k_riemann_pers (device memory, zeroed before every launch)
  int sms = 148;
  // k_trace_tma: maps of Q and E (encoded at the first launch: both arrays are allocated lazily) and of the face fields of
  // every conservative array that has been a step input (U, U2, the staging arrays of ppk_mhd3d_stage_*)
  EncodeTiledFn enc = nullptr;
  const double *tq = nullptr, *te = nullptr;
  CUtensorMap trace_q, trace_e;
  int ntu = 0;
  const double *tu[4] = {nullptr, nullptr, nullptr, nullptr};
  CUtensorMap trace_u[4][3];
  bool trace_attr = false;
};
static bool encode_map(EncodeTiledFn enc, CUtensorMap *m, const GridParams &g, const double *base, int ncomp, int xb, int yb, int zb) {
  const cuuint64_t dims[4] = {(cuuint64_t)g.isize, (cuuint64_t)g.jsize, (cuuint64_t)g.ksize, (cuuint64_t)ncomp};
  const cuuint64_t strides[3] = {(cuuint64_t)g.isize * 8, (cuuint64_t)g.isize * g.jsize * 8, (cuuint64_t)g.ncell * 8};
  const cuuint32_t box[4] = {(cuuint32_t)xb, (cuuint32_t)yb, (cuuint32_t)zb, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template <class Cfg> static bool encode_cfg(EncodeTiledFn enc, CUtensorMap *m, const GridParams &g, const double *base, int ncomp) {
  return encode_map(enc, m, g, base, ncomp, Cfg::XB, Cfg::YB, Cfg::ZB);
}
// returns nullptr when TMA staging cannot be used (odd isize => global strides not 16-byte multiples, or
// PPK_TMA=0): the plain kernels then do all the work
static void l_tma_destroy(void *ctx);
static void *l_tma_create(const GridParams &g, const double *BASIS, const double *DBF) {
  if (getenv("PPK_TMA") && atoi(getenv("PPK_TMA")) == 0) return nullptr;
  if ((g.isize & 1) || g.nx < 32) return nullptr;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return nullptr;
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  TmaCtx *c = new TmaCtx();
  c->enc = enc;
  bool ok = encode_cfg<EmfCfg<0>>(enc, &c->emfB[0], g, BASIS, NBASIS) && encode_cfg<EmfCfg<1>>(enc, &c->emfB[1], g, BASIS, NBASIS) &&
            encode_cfg<EmfCfg<2>>(enc, &c->emfB[2], g, BASIS, NBASIS) && encode_cfg<EmfCfg<0>>(enc, &c->emfD[0], g, DBF, NDBF) &&
            encode_cfg<EmfCfg<1>>(enc, &c->emfD[1], g, DBF, NDBF) && encode_cfg<EmfCfg<2>>(enc, &c->emfD[2], g, DBF, NDBF) &&
            encode_cfg<FluxCfg<0>>(enc, &c->fluxB[0], g, BASIS, NBASIS) && encode_cfg<FluxCfg<1>>(enc, &c->fluxB[1], g, BASIS, NBASIS) &&
            encode_cfg<FluxCfg<2>>(enc, &c->fluxB[2], g, BASIS, NBASIS);
  ok = ok && cudaFuncSetAttribute(k_emf_tma<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EmfCfg<0>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_emf_tma<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EmfCfg<1>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_emf_tma<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EmfCfg<2>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<0, false, RIEMANN_HLLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<0>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<0, false, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<0>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<1, false, RIEMANN_HLLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<1>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<1, false, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<1>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<2, false, RIEMANN_HLLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<2>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<2, false, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<2>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_emf_tma<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EmfCfg<0>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_emf_tma<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EmfCfg<1>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_emf_tma<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EmfCfg<2>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<0, true, RIEMANN_HLLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<0>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<0, true, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<0>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<1, true, RIEMANN_HLLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<1>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<1, true, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<1>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<2, true, RIEMANN_HLLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<2>::SMEM_BYTES) == cudaSuccess &&
       cudaFuncSetAttribute(k_flux_tma<2, true, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FluxCfg<2>::SMEM_BYTES) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(k_riemann_all<RIEMANN_HLLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, RALL_SMEM) == cudaSuccess &&
       cudaFuncSetAttribute(k_riemann_all<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RALL_SMEM) == cudaSuccess;
  if (!ok) { delete c; return nullptr; }
  c->xz_ok = encode_map(enc, &c->xzB, g, BASIS, NBASIS, XZGroup::XB, 1, XZGroup::ZB) &&
             encode_map(enc, &c->xzD, g, DBF, NDBF, XZGroup::XB, 1, XZGroup::ZB) &&
             cudaFuncSetAttribute(k_xz_group<RIEMANN_HLLD, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, XZGroup::SMEM_BYTES) == cudaSuccess &&
             cudaFuncSetAttribute(k_xz_group<RIEMANN_HLLD, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, XZGroup::SMEM_BYTES) == cudaSuccess &&
             cudaFuncSetAttribute(k_xz_group<RIEMANN_HLLD, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, XZGroup::SMEM_BYTES) == cudaSuccess &&
             cudaFuncSetAttribute(k_xz_group<-1, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, XZGroup::SMEM_BYTES) == cudaSuccess;
  for (int d = 0; d < 3; ++d) { c->rall.fluxB[d] = c->fluxB[d]; c->rall.emfB[d] = c->emfB[d]; c->rall.emfD[d] = c->emfD[d]; }
  if (!encode_cfg<EmfXCfg3>(enc, &c->rall.emfB[0], g, BASIS, NBASIS) || !encode_cfg<EmfXCfg3>(enc, &c->rall.emfD[0], g, DBF, NDBF)) {
    delete c;
    return nullptr;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaFuncSetAttribute(k_riemann_pers<RIEMANN_HLLD, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, RALL_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_riemann_pers<RIEMANN_HLLD, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, RALL_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_riemann_pers<-1, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, RALL_SMEM) != cudaSuccess ||
      cudaMalloc(&c->counter, sizeof(unsigned)) != cudaSuccess) {
    l_tma_destroy(c);
    return nullptr;
  }
  return c;
}
static void l_tma_destroy(void *ctx) {
  TmaCtx *c = (TmaCtx *)ctx;
  if (c && c->counter) cudaFree(c->counter);
  delete c;
}

// ComputeTrace: the TMA-staged kernel (mhd_trace_tma.inc) where its boxes exist (even isize, nx >= 32, a TMA context), else the
// plain-load kernel. PPK_TRACE_TMA=0 forces the plain kernel (A/B).
static void l_trace(const GridParams &g, const StepState *st, const double *U, const double *Q, const double *E, double *BASIS,
                    void *tma_, cudaStream_t s) {
  TmaCtx *c = (TmaCtx *)tma_;
  static const int use_tma = getenv("PPK_TRACE_TMA") ? atoi(getenv("PPK_TRACE_TMA")) : 1;
  if (!c || !use_tma || !c->enc || (unsigned long long)g.ncell * 8ull > 0xFFFFFFFFull) return l_trace_plain(g, st, U, Q, E, BASIS, s);
  using T = TraceTile;
  if (c->tq != Q || c->te != E) {  // (first launch: Q and E are allocated when a schedule first needs them)
    if (!encode_map(c->enc, &c->trace_q, g, Q, NBVAR, T::QX, T::QY, T::QZ) || !encode_map(c->enc, &c->trace_e, g, E, NELEC, T::EX, T::EY, T::EZ))
      return l_trace_plain(g, st, U, Q, E, BASIS, s);
    c->tq = Q; c->te = E;
  }
  if (!c->trace_attr) {
    if (cudaFuncSetAttribute(k_trace_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES) != cudaSuccess)
      return l_trace_plain(g, st, U, Q, E, BASIS, s);
    c->trace_attr = true;
  }
  int m = -1;
  for (int a = 0; a < c->ntu; ++a)
    if (c->tu[a] == U) m = a;
  if (m < 0) {  // an array that has not been a step input yet
    if (c->ntu >= 4) return l_trace_plain(g, st, U, Q, E, BASIS, s);
    m = c->ntu;
    if (!encode_map(c->enc, &c->trace_u[m][0], g, U, NBVAR, T::QX, T::TY, 1) || !encode_map(c->enc, &c->trace_u[m][1], g, U, NBVAR, T::QX, T::EY, 1) ||
        !encode_map(c->enc, &c->trace_u[m][2], g, U, NBVAR, T::QX, T::TY, 2))
      return l_trace_plain(g, st, U, Q, E, BASIS, s);
    c->tu[m] = U;
    c->ntu = m + 1;
  }
  TraceMaps maps;
  maps.q = c->trace_q; maps.e = c->trace_e;
  maps.ua = c->trace_u[m][0]; maps.ub = c->trace_u[m][1]; maps.uc = c->trace_u[m][2];
  TracePlan pl;
  const int nrows = g.jsize - 4, nk = g.ksize - 4;
  pl.ntx = (int)cdiv(g.isize - 4, T::TX);
  pl.rows = slab_rows(g, 46, nrows);
  pl.rows = (pl.rows + T::TY - 1) / T::TY * T::TY;
  pl.nslab = (int)cdiv(nrows, pl.rows);
  pl.per_plane = (unsigned)pl.ntx * (unsigned)(pl.rows / T::TY);
  pl.per_slab = (unsigned)nk * pl.per_plane;
  const unsigned long long total = (unsigned long long)pl.nslab * pl.per_slab;
  if (total > 0x7FFFFFFFull) return l_trace_plain(g, st, U, Q, E, BASIS, s);
  k_trace_tma<<<(unsigned)total, 128, T::SMEM_BYTES, s>>>(g, st, maps, pl, BASIS);
}

template <int D>
static void launch_flux(const GridParams &g, const double *BASIS, double *F, const TmaCtx *tma, cudaStream_t s) {
  static const int minb = getenv("PPK_FLUX_MINB") ? atoi(getenv("PPK_FLUX_MINB")) : 5;
  const int ni = g.nx + (D == 0), nj = g.ny + (D == 1), nk = g.nz + (D == 2);
  int done = 0;  // x-columns covered by full TMA tiles
  if (tma) {
    using Cfg = FluxCfg<D>;
    // a remainder of a few columns would make the plain kernel fetch 32-byte sectors for 8 or 16 useful bytes:
    // hand it a whole extra tile instead so that its rows stay coalesced
    int ntx = ni / Cfg::TX;
    if (ni - ntx * Cfg::TX == 1 && g.wrap_x && D == 0) {
      // nx a multiple of 32 and x periodic on both sides: the faces at i = nx+gw are bit-identical to those at i = gw
      // (ghost cells are copies, hence so are their basis numbers) and the update reads them there (k_update):
      // no launch for that column, whose strided 32-byte sectors cost 0.1 ms per kernel
      done = ni;
    } else {
      if (ni % Cfg::TX != 0 && ni % Cfg::TX < 8 && ntx > 1) --ntx;
      done = ntx * Cfg::TX;
    }
    const int nty = cdiv(nj, Cfg::TY), ntz = cdiv(nk, Cfg::TZ);
    // y-slabs only pay where a tile has a z halo (the plane below is re-read by the next z-tile)
    int yslab = Cfg::HZ > 0 ? slab_rows(g, 20 * Cfg::ZB, nj) / Cfg::TY : nty;
    if (yslab < 1) yslab = 1;
    if (yslab * Cfg::TY >= nj - Cfg::TY) yslab = nty;  // no sliver slabs
    if ((long long)yslab * ntz > 65535) yslab = 65535 / ntz > 0 ? 65535 / ntz : 1;  // grid.y limit
    dim3 grid(ntx, yslab * ntz, cdiv(nty, yslab));
    const unsigned ymagic = 0xFFFFFFFFu / (unsigned)yslab + 1u;
    if (g.riemann == RIEMANN_HLLD) {
      if (yslab == nty) k_flux_tma<D, false, RIEMANN_HLLD><<<dim3(ntx, nty, ntz), Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(g, tma->fluxB[D], F, 0, 0u);
      else k_flux_tma<D, true, RIEMANN_HLLD><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(g, tma->fluxB[D], F, yslab, ymagic);
    } else {
      if (yslab == nty) k_flux_tma<D, false, -1><<<dim3(ntx, nty, ntz), Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(g, tma->fluxB[D], F, 0, 0u);
      else k_flux_tma<D, true, -1><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(g, tma->fluxB[D], F, yslab, ymagic);
    }
  }
  if (done < ni) {
    const int bs = 128;
    dim3 grid(cdiv((long long)(ni - done) * nj, bs), (unsigned)nk);
    if (minb == 4) k_flux<D, 4><<<grid, bs, 0, s>>>(g, BASIS, F, done, ni - done);
    else if (minb == 6) k_flux<D, 6><<<grid, bs, 0, s>>>(g, BASIS, F, done, ni - done);
    else k_flux<D, 5><<<grid, bs, 0, s>>>(g, BASIS, F, done, ni - done);
  }
}
static void l_flux(const GridParams &g, int dir, const double *BASIS, double *F, const void *tma, cudaStream_t s) {
  if (dir == 0) launch_flux<0>(g, BASIS, F, (const TmaCtx *)tma, s);
  else if (dir == 1) launch_flux<1>(g, BASIS, F, (const TmaCtx *)tma, s);
  else launch_flux<2>(g, BASIS, F, (const TmaCtx *)tma, s);
}
template <int E>
static void launch_emf(const GridParams &g, const double *BASIS, const double *DBF, double *EMF, const TmaCtx *tma, cudaStream_t s) {
  static const int minb = getenv("PPK_EMF_MINB") ? atoi(getenv("PPK_EMF_MINB")) : 5;
  const int ni = g.nx + (E != 0), nj = g.ny + (E != 1), nk = g.nz + (E != 2);
  int done = 0;
  if (tma) {
    using Cfg = EmfCfg<E>;
    int ntx = ni / Cfg::TX;
    if (ni - ntx * Cfg::TX == 1 && g.wrap_x && E != 0) {
      // nx a multiple of 32 and x periodic on both sides: the edges at i = nx+gw are bit-identical to those at i = gw
      // (ghost cells are copies, hence so are their basis numbers) and the update reads them there (k_update):
      // no launch for that column, whose strided 32-byte sectors cost 0.1 ms per kernel
      done = ni;
    } else {
      if (ni % Cfg::TX != 0 && ni % Cfg::TX < 8 && ntx > 1) --ntx;
      done = ntx * Cfg::TX;
    }
    const int nty = cdiv(nj, Cfg::TY), ntz = cdiv(nk, Cfg::TZ);
    // y-slabs only pay where a tile has a z halo (the plane below is re-read by the next z-tile)
    int yslab = Cfg::HZ > 0 ? slab_rows(g, 20 * Cfg::ZB, nj) / Cfg::TY : nty;
    if (yslab < 1) yslab = 1;
    if (yslab * Cfg::TY >= nj - Cfg::TY) yslab = nty;  // no sliver slabs
    if ((long long)yslab * ntz > 65535) yslab = 65535 / ntz > 0 ? 65535 / ntz : 1;  // grid.y limit
    dim3 grid(ntx, yslab * ntz, cdiv(nty, yslab));
    if (yslab == nty) k_emf_tma<E, false><<<dim3(ntx, nty, ntz), Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(g, tma->emfB[E], tma->emfD[E], EMF, 0, 0u);
    else k_emf_tma<E, true><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(g, tma->emfB[E], tma->emfD[E], EMF, yslab, 0xFFFFFFFFu / (unsigned)yslab + 1u);
  }
  if (done < ni) {
    const int bs = 128;
    dim3 grid(cdiv((long long)(ni - done) * nj, bs), (unsigned)nk);
    if (minb == 4) k_emf<E, 4><<<grid, bs, 0, s>>>(g, BASIS, DBF, EMF, done, ni - done);
    else if (minb == 6) k_emf<E, 6><<<grid, bs, 0, s>>>(g, BASIS, DBF, EMF, done, ni - done);
    else k_emf<E, 5><<<grid, bs, 0, s>>>(g, BASIS, DBF, EMF, done, ni - done);
  }
}
static void l_emf(const GridParams &g, int e, const double *BASIS, const double *DBF, double *EMF, const void *tma, cudaStream_t s) {
  if (e == 0) launch_emf<0>(g, BASIS, DBF, EMF, (const TmaCtx *)tma, s);
  else if (e == 1) launch_emf<1>(g, BASIS, DBF, EMF, (const TmaCtx *)tma, s);
  else launch_emf<2>(g, BASIS, DBF, EMF, (const TmaCtx *)tma, s);
}
// x-faces + y-faces + z-edges of every plane in one launch on shared tiles (mhd_pgroup.inc); returns -1 when unavailable
// (no TMA context, PPK_PGROUP=0): the caller then launches the three kernels one by one.
#ifndef PPK_PGROUP_MINB_DEFAULT
#  define PPK_PGROUP_MINB_DEFAULT 4
#endif
static int l_plane_group(const GridParams &g, const double *BASIS, const double *DBF, double *F0, double *F1, double *EMF,
                         const void *tma_, cudaStream_t s) {
  const TmaCtx *tma = (const TmaCtx *)tma_;
  static const int on = getenv("PPK_PGROUP") ? atoi(getenv("PPK_PGROUP")) : 1;
  if (!tma || !on) return -1;
  const bool wrap = g.wrap_x && g.nx % 32 == 0;
  int ntx = g.nx / 32;
  if (!wrap && (g.nx + 1 - ntx * 32) < 8 && ntx > 1) --ntx;  // keep the left-over rows of the plain kernels coalesced
  if (ntx < 1) return -1;
  const int done = ntx * 32;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    if (cudaFuncSetAttribute(k_plane_group<RIEMANN_HLLD, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, PGroup::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_plane_group<RIEMANN_HLLD, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, PGroup::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(k_plane_group<-1, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, PGroup::SMEM_BYTES) != cudaSuccess)
      return -1;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  dim3 grid(ntx, cdiv(g.ny + 1, PGroup::TY), g.nz);
  // register target (read at every launch, A/B): 4 CTAs per SM at 128 registers without spills (512^3: 11.4 ms) beat 5 at 96
  // registers with 124 B of spills (12.4 ms)
  const char *e_minb = getenv("PPK_PGROUP_MINB");
  const int minb = e_minb ? atoi(e_minb) : PPK_PGROUP_MINB_DEFAULT;
  if (g.riemann != RIEMANN_HLLD) k_plane_group<-1, 5><<<grid, PGroup::THREADS, PGroup::SMEM_BYTES, s>>>(g, tma->fluxB[1], tma->emfD[2], F0, F1, EMF);
  else if (minb == 4) k_plane_group<RIEMANN_HLLD, 4><<<grid, PGroup::THREADS, PGroup::SMEM_BYTES, s>>>(g, tma->fluxB[1], tma->emfD[2], F0, F1, EMF);
  else k_plane_group<RIEMANN_HLLD, 5><<<grid, PGroup::THREADS, PGroup::SMEM_BYTES, s>>>(g, tma->fluxB[1], tma->emfD[2], F0, F1, EMF);
  if (!wrap) {
    const int bs = 128;
    if (done < g.nx + 1) k_flux<0, 5><<<dim3(cdiv((long long)(g.nx + 1 - done) * g.ny, bs), g.nz), bs, 0, s>>>(g, BASIS, F0, done, g.nx + 1 - done);
    if (done < g.nx) k_flux<1, 5><<<dim3(cdiv((long long)(g.nx - done) * (g.ny + 1), bs), g.nz), bs, 0, s>>>(g, BASIS, F1, done, g.nx - done);
    if (done < g.nx + 1) k_emf<2, 5><<<dim3(cdiv((long long)(g.nx + 1 - done) * (g.ny + 1), bs), g.nz), bs, 0, s>>>(g, BASIS, DBF, EMF, done, g.nx + 1 - done);
  }
  return 0;
}

// z-faces + y-edges of every row in one launch on shared x-z tiles (mhd_xzgroup.inc); returns -1 when unavailable (no TMA
// context, PPK_XZGROUP=0): the caller then launches flux(2) and emf(1).
#ifndef PPK_XZGROUP_DEFAULT
#  define PPK_XZGROUP_DEFAULT 1
#endif
static int l_xz_group(const GridParams &g, const double *BASIS, const double *DBF, double *F2, double *EMF, const void *tma_,
                      cudaStream_t s) {
  const TmaCtx *tma = (const TmaCtx *)tma_;
  // (the three knobs of this launcher are read at every launch, so that one process can A/B them: profiles/r2/ab_xz.py)
  const char *e_on = getenv("PPK_XZGROUP"), *e_slab = getenv("PPK_XZ_SLAB_MB"), *e_minb = getenv("PPK_XZ_MINB");
  const int on = e_on ? atoi(e_on) : PPK_XZGROUP_DEFAULT;
  if (!tma || !on || !tma->xz_ok) return -1;
  const bool wrap = g.wrap_x && g.nx % 32 == 0;
  int ntx = g.nx / 32;
  if (!wrap && (g.nx + 1 - ntx * 32) < 8 && ntx > 1) --ntx;  // keep the left-over rows of the plain kernels coalesced
  if (ntx < 1) return -1;
  const int done = ntx * 32;
  const int ntz = (int)cdiv(g.nz + 1, XZGroup::TZ);
  // y-slabs: the four planes a z-tile of a slab reads (23 streams each) pass through the L2 between two uses of its top plane
  const int slab_mb = e_slab && atoi(e_slab) > 0 ? atoi(e_slab) : 48;
  long long rows = ((long long)slab_mb << 20) / ((long long)XZGroup::TZ * XZGroup::NSLOT * g.isize * 8);
  rows &= ~7LL;
  if (rows < 8) rows = 8;
  if (rows >= g.ny - 8) rows = g.ny;  // no sliver slab
  if (rows * ntz > 65535) rows = 65535 / ntz;  // grid.y limit
  if (rows < 1) return -1;
  const int yslab = (int)rows;
  dim3 grid(ntx, (unsigned)(yslab * ntz), cdiv(g.ny, yslab));
  const unsigned ymagic = yslab > 1 ? 0xFFFFFFFFu / (unsigned)yslab + 1u : 0u;
  const int minb = e_minb ? atoi(e_minb) : 4;  // register target: 4 CTAs per SM at 128 registers without spills (512^3: 8.67 ms) beat 5 at 96 (9.3 ms) and 6 at 80 (11.5 ms)
  if (g.riemann != RIEMANN_HLLD) k_xz_group<-1, 5><<<grid, XZGroup::THREADS, XZGroup::SMEM_BYTES, s>>>(g, tma->xzB, tma->xzD, F2, EMF, yslab, ymagic);
  else if (minb == 5) k_xz_group<RIEMANN_HLLD, 5><<<grid, XZGroup::THREADS, XZGroup::SMEM_BYTES, s>>>(g, tma->xzB, tma->xzD, F2, EMF, yslab, ymagic);
  else if (minb == 6) k_xz_group<RIEMANN_HLLD, 6><<<grid, XZGroup::THREADS, XZGroup::SMEM_BYTES, s>>>(g, tma->xzB, tma->xzD, F2, EMF, yslab, ymagic);
  else k_xz_group<RIEMANN_HLLD, 4><<<grid, XZGroup::THREADS, XZGroup::SMEM_BYTES, s>>>(g, tma->xzB, tma->xzD, F2, EMF, yslab, ymagic);
  if (!wrap) {
    const int bs = 128;
    if (done < g.nx) k_flux<2, 5><<<dim3(cdiv((long long)(g.nx - done) * g.ny, bs), g.nz + 1), bs, 0, s>>>(g, BASIS, F2, done, g.nx - done);
    if (done < g.nx + 1) k_emf<1, 5><<<dim3(cdiv((long long)(g.nx + 1 - done) * g.ny, bs), g.nz + 1), bs, 0, s>>>(g, BASIS, DBF, EMF, done, g.nx + 1 - done);
  }
  return 0;
}

// The six flux / EMF tasks in one L2-ordered launch (k_riemann_all); returns -1 when the TMA context is missing
// (the caller then launches the tasks one by one). Columns beyond the last full 32-wide tile go through the plain
// kernels, task by task, exactly like launch_flux / launch_emf.
static int l_riemann_all(const GridParams &g, const double *BASIS, const double *DBF, double *F0, double *F1, double *F2,
                         double *EMF, const void *tma_, cudaStream_t s) {
  const TmaCtx *tma = (const TmaCtx *)tma_;
  if (!tma) return -1;
  static const int off = getenv("PPK_RALL") ? atoi(getenv("PPK_RALL")) == 0 : 0;
  if (off) return -1;
  const bool wrap = g.wrap_x && g.nx % 32 == 0;
  int ntx = g.nx / 32;
  if (!wrap && (g.nx + 1 - ntx * 32) < 8 && ntx > 1) --ntx;  // keep the left-over rows of the plain kernels coalesced
  const int done = ntx * 32;
  RiemannPlan pl;
  pl.ntx = ntx;
  pl.nrows = g.ny + 1;
  // y-slabs: two planes (k-1, k) of the 38 numbers a slab of rows needs stay in the L2 while the six tasks run
  static const int slab_mb = getenv("PPK_RALL_SLAB_MB") ? atoi(getenv("PPK_RALL_SLAB_MB")) : 40;
  long long rows = ((long long)slab_mb << 20) / (2LL * 38 * g.isize * 8);
  rows = rows / 12 * 12;
  if (rows < 12) rows = 12;
  if (rows >= pl.nrows - 12) rows = (pl.nrows + 11) / 12 * 12;  // no sliver slab
  pl.rows = (int)rows;
  const unsigned nslab = cdiv(pl.nrows, rows);
  pl.per_task4 = (unsigned)ntx * (unsigned)(rows / 4);
  pl.per_task3 = (unsigned)ntx * (unsigned)(rows / 3);
  // persistent CTAs (mhd_rpers.inc, PPK_RALL_PERS=1) enumerate the six tasks one by one; the single-shot kernel merges the
  // three plane-k tasks of a tile into one work item (PPK_RALL_MERGE=0: six items, for A/B)
  static const int pers_env = getenv("PPK_RALL_PERS") ? atoi(getenv("PPK_RALL_PERS")) : 0;
  static const int merge_env = getenv("PPK_RALL_MERGE") ? atoi(getenv("PPK_RALL_MERGE")) : 1;
  pl.merged = (!pers_env && merge_env) ? 1 : 0;
  pl.per_plane = (pl.merged ? 3u : 5u) * pl.per_task4 + pl.per_task3;
  pl.per_slab = (unsigned)(g.nz + 1) * pl.per_plane;
  const unsigned long long total = (unsigned long long)nslab * pl.per_slab;
  if (total > 0x7FFFFFFFull || ntx < 1) return -1;
  // PPK_RALL_PERS=1 selects the persistent form (mhd_rpers.inc): built, bit-identical, and slower on B200 (256^3: 4.55 -
  // 5.26 ms against 3.86 ms; the gathered operands of a solve have to live in registers across the hand-off, which costs
  // either a CTA per SM or spills, and the solves are latency-bound, not load-bound: profiles/r2/README.md)
  static const int pers = getenv("PPK_RALL_PERS") ? atoi(getenv("PPK_RALL_PERS")) : 0;
  static const int pers_ctas = getenv("PPK_RALL_CTAS") ? atoi(getenv("PPK_RALL_CTAS")) : 5;
  static const int xmode = getenv("PPK_RALL_XMODE") ? atoi(getenv("PPK_RALL_XMODE")) : 0;
  if (pers && total > (unsigned long long)tma->sms * pers_ctas) {
    const unsigned G = (unsigned)(tma->sms * pers_ctas);
    cudaMemsetAsync(tma->counter, 0, sizeof(unsigned), s);
    if (g.riemann == RIEMANN_HLLD && pers_ctas == 4) k_riemann_pers<RIEMANN_HLLD, 4><<<G, 128, RALL_SMEM, s>>>(g, pl, tma->rall, F0, F1, F2, EMF, (unsigned)total, tma->counter, xmode);
    else if (g.riemann == RIEMANN_HLLD) k_riemann_pers<RIEMANN_HLLD, 5><<<G, 128, RALL_SMEM, s>>>(g, pl, tma->rall, F0, F1, F2, EMF, (unsigned)total, tma->counter, xmode);
    else k_riemann_pers<-1, 5><<<G, 128, RALL_SMEM, s>>>(g, pl, tma->rall, F0, F1, F2, EMF, (unsigned)total, tma->counter, xmode);
  } else if (g.riemann == RIEMANN_HLLD) k_riemann_all<RIEMANN_HLLD><<<(unsigned)total, 128, RALL_SMEM, s>>>(g, pl, tma->rall, F0, F1, F2, EMF);
  else k_riemann_all<-1><<<(unsigned)total, 128, RALL_SMEM, s>>>(g, pl, tma->rall, F0, F1, F2, EMF);
  if (!wrap) {
    const int bs = 128;
    const int nif[3] = {g.nx + 1, g.nx, g.nx}, nie[3] = {g.nx, g.nx + 1, g.nx + 1};
    if (done < nif[0]) k_flux<0, 5><<<dim3(cdiv((long long)(nif[0] - done) * g.ny, bs), g.nz), bs, 0, s>>>(g, BASIS, F0, done, nif[0] - done);
    if (done < nif[1]) k_flux<1, 5><<<dim3(cdiv((long long)(nif[1] - done) * (g.ny + 1), bs), g.nz), bs, 0, s>>>(g, BASIS, F1, done, nif[1] - done);
    if (done < nif[2]) k_flux<2, 5><<<dim3(cdiv((long long)(nif[2] - done) * g.ny, bs), g.nz + 1), bs, 0, s>>>(g, BASIS, F2, done, nif[2] - done);
    if (done < nie[2]) k_emf<2, 5><<<dim3(cdiv((long long)(nie[2] - done) * (g.ny + 1), bs), g.nz), bs, 0, s>>>(g, BASIS, DBF, EMF, done, nie[2] - done);
    if (done < nie[1]) k_emf<1, 5><<<dim3(cdiv((long long)(nie[1] - done) * g.ny, bs), g.nz + 1), bs, 0, s>>>(g, BASIS, DBF, EMF, done, nie[1] - done);
    if (done < nie[0]) k_emf<0, 5><<<dim3(cdiv((long long)(nie[0] - done) * (g.ny + 1), bs), g.nz + 1), bs, 0, s>>>(g, BASIS, DBF, EMF, done, nie[0] - done);
  }
  return 0;
}
static void l_update(const GridParams &g, const StepState *st, const double *Uin, double *Uout, const double *Fx,
                     const double *Fy, const double *Fz, const double *EMF, int k0, int k1, cudaStream_t s) {
  if (k1 <= k0) return;
  const int bs = 256;
  const int rows = slab_rows(g, 34, g.jsize);
  dim3 grid(cdiv((long long)g.isize * rows, bs), k1 - k0, cdiv(g.jsize, rows));
  static const int minb = getenv("PPK_UPDATE_MINB") ? atoi(getenv("PPK_UPDATE_MINB")) : 0;
  if (minb == 6) k_update<6><<<grid, bs, 0, s>>>(g, st, Uin, Uout, Fx, Fy, Fz, EMF, k0, rows);
  else if (minb == 5) k_update<5><<<grid, bs, 0, s>>>(g, st, Uin, Uout, Fx, Fy, Fz, EMF, k0, rows);
  else k_update<0><<<grid, bs, 0, s>>>(g, st, Uin, Uout, Fx, Fy, Fz, EMF, k0, rows);
}
static void l_consume(const GridParams &g, const StepState *st, const double *BASIS, const double *DBF, const double *Uin,
                      double *Uout, int split, cudaStream_t s) {
  dim3 block(CBX, CBY);
  // z-chunks: as few as give >= ~8 waves of CTAs on 148 SMs x 2 CTAs (each chunk re-solves one plane of z-problems)
  const long long tiles = (long long)cdiv(g.nx, CBX - 1) * cdiv(g.ny, CBY - 1);
  int nchunk = (int)((8LL * 296 + tiles - 1) / tiles);
  const int max_chunks = g.nz / 16 > 0 ? g.nz / 16 : 1;
  if (nchunk > max_chunks) nchunk = max_chunks;
  if (nchunk < 1) nchunk = 1;
  const int zchunk = (g.nz + nchunk - 1) / nchunk;
  dim3 grid(cdiv(g.nx, CBX - 1), cdiv(g.ny, CBY - 1), cdiv(g.nz, zchunk));
  if (!split) k_consume<true, true><<<grid, block, 0, s>>>(g, st, BASIS, DBF, Uin, Uout, zchunk);
  else {
    k_consume<true, false><<<grid, block, 0, s>>>(g, st, BASIS, DBF, Uin, Uout, zchunk);
    k_consume<false, true><<<grid, block, 0, s>>>(g, st, BASIS, DBF, Uin, Uout, zchunk);
  }
}
static void l_hydro(const GridParams &g, const StepState *st, const double *BASIS, const double *Uin, double *Uout, cudaStream_t s) {
  static const int zc_env = getenv("PPK_HYDRO_ZCHUNK") ? atoi(getenv("PPK_HYDRO_ZCHUNK")) : 0;
  static const int minb = getenv("PPK_HYDRO_MINB") ? atoi(getenv("PPK_HYDRO_MINB")) : 1;  // 1: 168 registers, no spills (fastest measured)
  const long long tiles = (long long)cdiv(g.nx, 32) * cdiv(g.ny, SHY);
  // z-chunks only exist to give the grid enough CTAs (each one re-solves one plane of z-faces): ~10 waves of 148 x 2
  int nchunk = (int)((10LL * 296 + tiles - 1) / tiles);
  const int max_chunks = g.nz / 8 > 0 ? g.nz / 8 : 1;
  if (nchunk > max_chunks) nchunk = max_chunks;
  if (nchunk < 1) nchunk = 1;
  int zchunk = (g.nz + nchunk - 1) / nchunk;
  if (zc_env > 0) zchunk = zc_env;
  dim3 grid(cdiv(g.nx, 32), cdiv(g.ny, SHY), cdiv(g.nz, zchunk));
  // the dynamic shared-memory limit is a per-device attribute of the function: set it once on every device used
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    if (cudaFuncSetAttribute(k_hydro<168>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HydroSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_hydro<112>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HydroSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_hydro<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HydroSmem)) != cudaSuccess)
      return;  // the launch below would fail anyway; the sticky error reaches the caller through cudaGetLastError
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  if (minb == 1) k_hydro<168><<<grid, SH_THREADS, sizeof(HydroSmem), s>>>(g, st, BASIS, Uin, Uout, zchunk);
  else if (minb == 3) k_hydro<96><<<grid, SH_THREADS, sizeof(HydroSmem), s>>>(g, st, BASIS, Uin, Uout, zchunk);
  else k_hydro<112><<<grid, SH_THREADS, sizeof(HydroSmem), s>>>(g, st, BASIS, Uin, Uout, zchunk);
}
static void l_update_ct(const GridParams &g, const StepState *st, const double *Uin, double *Uout, const double *EMF, cudaStream_t s) {
  const int bs = 256;
  dim3 grid(cdiv((long long)g.nx * g.ny, bs), g.nz);
  k_update_ct<<<grid, bs, 0, s>>>(g, st, Uin, Uout, EMF);
}
static void l_wrap_x_column(const GridParams &g, double *A, int ncomp, cudaStream_t s) {
  if (!g.wrap_x) return;
  const long long total = (long long)g.jsize * g.ksize * ncomp;
  k_wrap_x_column<<<cdiv(total, 256), 256, 0, s>>>(g, A, ncomp);
}
static void l_diagnostics(const GridParams &g, const double *U, double *out9, cudaStream_t s) {
  const int bs = 256;
  dim3 grid(cdiv((long long)g.nx * g.ny, bs), g.nz);
  k_diagnostics<<<grid, bs, 0, s>>>(g, U, out9);
}

static void l_fastmath_selftest(int n, const double *x, double *rcp, double *sq, double *rsq, cudaStream_t s) {
  k_fastmath_selftest<<<(n + 255) / 256, 256, 0, s>>>(n, x, rcp, sq, rsq);
}

#include "mhd_prod.inc"
#include "mhd2d_kernels.inc"

static const KernelTable table = {
#if PPK_EXACT
  "exact",
#else
  "fast",
#endif
  l_boundary, l_prim_dt, l_finalize_dt, l_advance_time, l_elec_dbf, l_trace, l_flux, l_emf, l_update, l_diagnostics, l_fastmath_selftest, l_consume, l_tma_create, l_tma_destroy, l_hydro, l_update_ct, l_wrap_x_column,
  l2_boundary, l2_prim_dt, l2_trace, l2_flux_emf, l2_update,
  l_dt_only, l_prod_create, l_prod_destroy, l_producer, l_riemann_all, l_face_copy, l_plane_group, l_xz_group,
};

}  // namespace PPK_NS

#if PPK_EXACT
const KernelTable *kernel_table_exact() { return &ppk_exact::table; }
#else
const KernelTable *kernel_table_fast() { return &ppk_fast::table; }
#endif

}  // namespace ppk
