// gr2m_math.cuh -- FP64 device arithmetic of the GR2M step (sm_100a).
//
// Three formulations of the same month step (airGR frun_GR2M.f90 MOD_GR2M, reached
// from R/Run_GR2MSemiDistr.R:181-184 and R/Optim_GR2MSemiDistr.R:174-177):
//   step_literal : IEEE division, libdevice exp()/pow(), operation order as written in airGR.
//   step_fast    : same algebra with the two tanh divisions folded into the store updates
//                  (3 divisions instead of 12), a table-driven exp, a Halley-refined reciprocal cube
//                  root and Newton-refined MUFU reciprocals; returns every series (simulation mode).
//   step_fast4   : calibration mode (only Q leaves the thread): production store carried as S/X1, exp
//                  arguments in units of a 2048-entry table's index, halved tanh terms straight from
//                  the table, rint by the conversion unit, one division step less.
// Every building block is accurate to 1-2 ulp, so they agree to ~1e-14 relative; tests hold all
// of them to 1e-10 against the CPU oracle.
// The forcing correction round(f*x, 1) (R/Run_GR2MSemiDistr.R:101-102) is BIT-EXACT in both: a wrong
// rounding decision is a 0.1 mm error, not an ulp.
#pragma once

#include "common.cuh"

// 2^(j/32), j = 0..31 (correctly rounded); copied to shared memory by the kernels that use exp_tab
__constant__ double c_exp2_tab[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

// 2^(j/256), j = 0..255: the table the hot kernels stage in shared memory (2 KiB)
__constant__ double c_exp2_tab256[256] = {
#include "exp2_tab256.inc"
};

// 2^(j/2048), j = 0..2047 (16 KiB, global memory; staged in shared memory by k_gr2m_calib2)
__device__ const double g_exp2_tab2048[2048] = {
#include "exp2_tab2048.inc"
};

#define GR2M_TIE_THR 0.49999            /* see r_round1 */
#define GR2M_MAGIC 6755399441055744.0   /* 1.5 * 2^52: (x + M) - M == rint(x) for |x| < 2^51 */

// 64-bit literals that do not fit an instruction immediate.  Kept in constant memory so DFMA/DMUL
// read them as c[bank][offset] operands instead of re-materialising them with two MOVs per use
// (the first build spent ~60 of its ~200 instructions per month on such moves).
struct MathK {
    double magic, nmagic, tenth, inv_ln2_32, nln2_32_hi, ln2_32_lo, c6, c5, c4, c3, two_ninths, third, tie_thr,
        nf32_delta, inv_ln2_256, nln2_256_hi, nln2_256_lo, inv_ln2_2048, nln2_2048, e1, e2, e3, tie_thr_tab;
};
__constant__ MathK KC = {GR2M_MAGIC,
                         -GR2M_MAGIC,
                         0.1,
                         0x1.71547652b82fep+5,    // 32 / ln 2
                         -0x1.62e42ff000000p-6,   // -(ln2/32) high part (29 significant bits)
                         0x1.718432a1b0e26p-40,   // -(ln2/32) low part
                         0x1.6c16c16c16c17p-10,   // 1/720
                         0x1.1111111111111p-7,    // 1/120
                         0x1.5555555555555p-5,    // 1/24
                         0x1.5555555555555p-3,    // 1/6
                         0x1.c71c71c71c71cp-3,    // 2/9
                         0x1.5555555555555p-2,    // 1/3
                         GR2M_TIE_THR,
                         -9.9341074625651041667e-9,   // -((double)(1.f/3.f) - 1/3)
                         0x1.71547652b82fep+8,        // 256 / ln 2
                         -0x1.62e42fefa0000p-9,       // -(ln2/256) high part (38 significant bits)
                         -0x1.cf79abc9e3b3ap-48,      // -(ln2/256) low part
                         0x1.71547652b82fep+11,       // 2048 / ln 2
                         -0x1.62e42fefa39efp-12,      // -(ln2/2048) rounded to nearest
                         0x1.62e42fefa39efp-12,       // L = ln2/2048
                         0x1.ebfbdff82c58fp-25,       // L^2/2
                         0x1.c6b08d704a0c0p-38,       // L^3/6
                         0.49999999};                 // tie window of the round-2 table-mode step (|10 f x| < 2^16), gr2m_calib3.cuh

// ---- R round(x, 1) ------------------------------------------------------------------------------
// Correctly rounded n/10 for integer-valued |n| < 2^52 (Markstein: 0.1 is RN(1/10), q0 faithful).
__device__ __forceinline__ double div10_rn(double n)
{
    double q = __dmul_rn(n, KC.tenth);
    double r = __fma_rn(q, -10.0, n);
    return __fma_rn(r, KC.tenth, q);
}

// R >= 4.0 algorithm verbatim: candidates floor/ceil(10x)/10, closer one wins in
// double arithmetic, even last digit on a tie.  Any finite or non-finite x.
__device__ __noinline__ double r_round1_exact(double x)
{
    double a = fabs(x);
    if (!(a < 1e14) || a == 0.0) return x;
    double x10 = __dmul_rn(10.0, a);
    double i10 = floor(x10);
    double xd = div10_rn(i10), xu = div10_rn(ceil(x10));
    double du = __dsub_rn(xu, a), dd = __dsub_rn(a, xd);
    bool even = __dsub_rn(i10, __dmul_rn(2.0, floor(__dmul_rn(0.5, i10)))) == 0.0;
    double r = (dd < du || (dd == du && even)) ? xd : xu;
    return copysign(r, x);
}

// Same result, cheap common case: away from a decimal tie the winner is rint(10x)/10 (for either sign).
// R's candidate comparison can only disagree with rint(10x) when 10x is within a few ulp(10x) of a tie, so
// the cheap path is taken when |10x - rint(10x)| < GR2M_TIE_THR = 0.49999 and |10x| < 1e9 (ulp(1e9) = 1.2e-7,
// 80 times inside the 1e-5 window); everything else -- ties, huge values, NaN -- takes the exact path.  `thr`
// is GR2M_TIE_THR, or -1 to send every value down the exact path.  With the earlier 1e-3 window one warp in
// eight (64 corrections per warp and month) left the fast path; now one in 800.
__device__ __forceinline__ double r_round1(double x, double thr)
{
    double x10 = __dmul_rn(x, 10.0);
    double n = __dadd_rn(__dadd_rn(x10, KC.magic), KC.nmagic);
    double f = __dsub_rn(x10, n);
    if (!(fabs(f) < thr) || !(fabs(x10) < 1e9)) return r_round1_exact(x);
    return div10_rn(n);
}

// Same, with rint(10x) taken by the conversion unit (F2I + I2F on the XU pipe) instead of two DADDs on
// the FP64 pipe.  |10x| >= 2^31 saturates the conversion, which makes |f| huge and falls to the exact path.
__device__ __forceinline__ double r_round1_cvt(double x, double thr)
{
    double x10 = __dmul_rn(x, 10.0);
    double n = (double)__double2int_rn(x10);
    double f = __dsub_rn(x10, n);
    if (!(fabs(f) < thr)) return r_round1_exact(x);
    return div10_rn(n);
}

// general digits (1..3 used), IEEE division; for the objective kernel (not hot)
__device__ double r_round_dig(double x, int dig)
{
    double a = fabs(x);
    if (!(a < 1.7e308) || a == 0.0) return x;      // NaN, Inf, 0
    int e10 = (int)floor(log10(a)) + 1;
    if (dig + e10 > 15) return x;
    double p10 = dig == 1 ? 10.0 : (dig == 2 ? 100.0 : 1000.0);
    double x10 = __dmul_rn(p10, a);
    double i10 = floor(x10);
    double xd = __ddiv_rn(i10, p10), xu = __ddiv_rn(ceil(x10), p10);
    double du = __dsub_rn(xu, a), dd = __dsub_rn(a, xd);
    bool even = fmod(i10, 2.0) == 0.0;
    double r = (dd < du || (dd == du && even)) ? xd : xu;
    return copysign(r, x);
}

// ---- fast building blocks -----------------------------------------------------------------------
// exp(x) for moderate |x| (here 0 <= x <= 26): k = rint(32 x / ln2), r = x - k ln2/32 (|r| <= ln2/64),
// exp(x) = 2^(k>>5) * 2^((k&31)/32) * (1 + r + ... + r^6/720); truncation 3.5e-18, total ~1 ulp.
__device__ __forceinline__ double exp_tab(double x, const double *__restrict__ tab)
{
    double t = __fma_rn(x, KC.inv_ln2_32, KC.magic);
    int ki = __double2loint(t);
    double kd = __dadd_rn(t, KC.nmagic);
    double r = __fma_rn(kd, KC.nln2_32_hi, x);
    r = __fma_rn(kd, KC.ln2_32_lo, r);
    double p = __fma_rn(r, KC.c6, KC.c5);
    p = __fma_rn(p, r, KC.c4);
    p = __fma_rn(p, r, KC.c3);
    p = __fma_rn(p, r, 0.5);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    double v = __dmul_rn(p, tab[ki & 31]);
    return __hiloint2double(__double2hiint(v) + ((ki >> 5) << 20), __double2loint(v));
}

// Same scheme with a 256-entry table: |r| <= ln2/512 = 1.35e-3, degree-4 Taylor (remainder 3.8e-17),
// 9 FP64-pipe instructions instead of 11.
__device__ __forceinline__ double exp_tab256(double x, const double *__restrict__ tab)
{
    double t = __fma_rn(x, KC.inv_ln2_256, KC.magic);
    int ki = __double2loint(t);
    double kd = __dadd_rn(t, KC.nmagic);
    double r = __fma_rn(kd, KC.nln2_256_hi, x);
    r = __fma_rn(kd, KC.nln2_256_lo, r);
    double p = __fma_rn(r, KC.c4, KC.c3);
    p = __fma_rn(p, r, 0.5);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    double v = __dmul_rn(p, tab[ki & 255]);
    return __hiloint2double(__double2hiint(v) + ((ki >> 8) << 20), __double2loint(v));
}

// 2048-entry table: |r| <= ln2/4096 = 1.7e-4, degree-3 Taylor (remainder 3.4e-17) and a one-FMA argument
// reduction (here 0 <= x <= 26, so |k| <= 76 800 and the rounding of ln2/2048 costs <= 1.4e-15 on r):
// 7 FP64-pipe instructions.
__device__ __forceinline__ double exp_tab2048(double x, const double *__restrict__ tab)
{
    double t = __fma_rn(x, KC.inv_ln2_2048, KC.magic);
    int ki = __double2loint(t);
    double kd = __dadd_rn(t, KC.nmagic);
    double r = __fma_rn(kd, KC.nln2_2048, x);      // |kd| <= 76 800: the constant's rounding costs <= 1.4e-15 on r
    double p = __fma_rn(r, KC.c3, 0.5);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    double v = __dmul_rn(p, tab[ki & 2047]);
    return __hiloint2double(__double2hiint(v) + ((ki >> 11) << 20), __double2loint(v));
}

__device__ __forceinline__ double rcp_seed(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    return r;
}

// n / d for normal-range d (denominators here are >= 2): MUFU seed, one Newton step on the
// reciprocal, one residual correction on the quotient -> ~0.5-1 ulp.
__device__ __forceinline__ double div_fast(double n, double d)
{
    double r = rcp_seed(d);
    double e = __fma_rn(-d, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-d, r, 1.0);
    r = __fma_rn(r, e, r);
    double q = __dmul_rn(n, r);
    double rem = __fma_rn(-d, q, n);
    return __fma_rn(rem, r, q);
}

// one Newton step only: enough if the MUFU seed is good to ~2^-20 (error 2^-40 on r, squared away by
// the quotient correction).  Measured against IEEE division in tests before it is used.
__device__ __forceinline__ double div_fast1(double n, double d)
{
    double r = rcp_seed(d);
    double e = __fma_rn(-d, r, 1.0);
    r = __fma_rn(r, e, r);
    double q = __dmul_rn(n, r);
    double rem = __fma_rn(-d, q, n);
    return __fma_rn(rem, r, q);
}

// cubic (Halley-type) refinement of the seed, no residual correction: r1 = r0 (1 + e + e^2), e = 1 - d r0;
// 4 FP64-pipe instructions, error ~ e^3 + 1.5 ulp.  Needs a seed good to ~2^-18 (checked in tests).
__device__ __forceinline__ double div_fast3(double n, double d)
{
    double r = rcp_seed(d);
    double e = __fma_rn(-d, r, 1.0);
    double t = __fma_rn(e, e, e);
    r = __fma_rn(r, t, r);
    return __dmul_rn(n, r);
}

// Same quotient with the numerator multiplied in early: q0 = n r0 runs beside e and t, q = q0 + q0 t.  Three dependent
// FP64 instructions after the seed instead of four (the recursion is a latency chain), same instruction count.
__device__ __forceinline__ double div_fast3q(double n, double d)
{
    double r = rcp_seed(d);
    double q0 = __dmul_rn(n, r);
    double e = __fma_rn(-d, r, 1.0);
    double t = __fma_rn(e, e, e);
    return __fma_rn(q0, t, q0);
}

// c^(-1/3) for c in [1, 2] (works for any normal c > 0): FP32 MUFU seed (~2^-21), one Halley step
// y <- y + y e (1/3 + 2/9 e), e = 1 - c y^3   (cubic: residual ~ e^3)
__device__ __forceinline__ double rcbrt_fast(double c)
{
    float cf = __double2float_rn(c);
    float l, y0;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(cf));
    l *= -0.333333343f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(l));
    double y = (double)y0;
    double y3 = __dmul_rn(__dmul_rn(y, y), y);
    double e = __fma_rn(-c, y3, 1.0);
    double q = __fma_rn(e, KC.two_ninths, KC.third);
    return __fma_rn(__dmul_rn(y, e), q, y);
}

// Same for c in [1, 2] only, without the conversion and transcendental units: the FP32 view of c by bit surgery
// (truncation, exact to 2^-23), a degree-6 FP32 polynomial in (c - 1.5) for the seed (Chebyshev-node fit, 2^-19.8
// incl. FP32 rounding) widened back by bit surgery, then the same Halley step: final error <= 1 ulp (tests).
// 12 ALU/FMA-pipe instructions instead of F2F + LG2 + FMUL + EX2 + F2F on the XU pipe, and a shorter chain.
__device__ __forceinline__ double rcbrt_fast_poly(double c)
{
    const float cf = __int_as_float(__funnelshift_l(__double2loint(c), __double2hiint(c), 3) + 0x40000000);
    const float t = cf - 1.5f;
    float y0 = 0.01006671879440546f;
    y0 = fmaf(y0, t, -0.016902167350053787f);
    y0 = fmaf(y0, t, 0.024652300402522087f);
    y0 = fmaf(y0, t, -0.04440813511610031f);
    y0 = fmaf(y0, t, 0.08628594130277634f);
    y0 = fmaf(y0, t, -0.19413940608501434f);
    y0 = fmaf(y0, t, 0.8735804557800293f);
    const int yb = __float_as_int(y0);
    const double y = __hiloint2double((int)((unsigned)yb >> 3) + 0x38000000, yb << 29);
    double y3 = __dmul_rn(__dmul_rn(y, y), y);
    double e = __fma_rn(-c, y3, 1.0);
    double q = __fma_rn(e, KC.two_ninths, KC.third);
    return __fma_rn(__dmul_rn(y, e), q, y);
}

// (a, b) of a piecewise-linear FP32 seed of c^(-1/3) on [1, 2] (256 intervals; scripts/gen_rcbrt_table.py), 4 KiB
__device__ const float2 g_rcbrt_lin[512] = {
#include "rcbrt_lin256.inc"
};

// Same for c in [1, 2] with the seed read from a shared-memory copy of g_rcbrt_lin at the 4 KiB-aligned shared address
// `tab`: the entry index is bits 20..12 of the high word (lowest exponent bit + 8 mantissa bits, so c = 2.0 has its own
// entry), seed = a + b c in one FFMA, 2^-20.8; then the same Halley step (final error <= 1 ulp).  6 instructions
// before the Halley step instead of the polynomial's 10.
__device__ __forceinline__ double rcbrt_fast_tab(double c, uint32_t tab)
{
    const uint32_t hi = (uint32_t)__double2hiint(c);
    const float cf = __int_as_float(__funnelshift_l(__double2loint(c), hi, 3) + 0x40000000);
    float a, b;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(((hi >> 9) & 0xff8u) | tab));
    const double y = (double)fmaf(b, cf, a);       // one F2F on the conversion unit (two integer instructions measured the same)
    double y3 = __dmul_rn(__dmul_rn(y, y), y);
    double e = __fma_rn(-c, y3, 1.0);
    double q = __fma_rn(e, KC.two_ninths, KC.third);
    return __fma_rn(__dmul_rn(y, e), q, y);
}

// ---- the month step -----------------------------------------------------------------------------
struct CellPar {
    double X1, X2, iX1, i2X1, fp, fe, area, thr;   // i2X1 = 2/X1; thr: see r_round1
    double fp10, fe10, c01, g, cP, kappa;          // 10 fp, 10 fe, 0.1 * 2/X1, c01 * 2048/ln2, 0.1/X1, 60/X1 (calibration kernel only)
};

struct StepOut {
    double Pc, AE, Q;   // corrected precipitation, actual evaporation, runoff [mm]; S, R updated in place
};

#define GR2M_F32_EXP_DELTA 9.9341074625651041667e-9   /* (double)(1.f/3.f) - 1/3 */

// min(x, 26) for the exp argument, decided on the high word by the integer ALU instead of a DSETP on the
// FP64 pipe: for x >= 26 (or NaN, like fmin) the high word is >= that of 26.0, whose low word is 0;
// negative x has a negative high word and passes through.
__device__ __forceinline__ double clamp26(double x)
{
    return __double2hiint(x) >= 0x403A0000 ? 26.0 : x;
}

__device__ __forceinline__ StepOut step_fast(const CellPar &p, double Praw, double Eraw, double &S, double &R,
                                             bool f32exp, const double *__restrict__ tab)
{
    StepOut o;
    double Pc = r_round1(__dmul_rn(p.fp, Praw), p.thr);
    double Ec = r_round1(__dmul_rn(p.fe, Eraw), p.thr);
    // production store: S1 = (S + X1 T)/(1 + S/X1 T), T = tanh(w) = (a-1)/(a+1), a = exp(2w), w <= 13
    double a = exp_tab(clamp26(Pc * p.i2X1), tab);
    double am1 = a - 1.0, ap1 = a + 1.0;
    double s = S * p.iX1;
    double S1 = div_fast1(__fma_rn(p.X1, am1, S * ap1), __fma_rn(s, am1, ap1));
    double P1 = (Pc + S) - S1;
    // evaporation: S2 = S1 (1-T)/(1 + (1 - S1/X1) T) = 2 S1 / ((b+1) + (1 - S1/X1)(b-1)), b = exp(2w)
    double b = exp_tab(clamp26(Ec * p.i2X1), tab);
    double s1 = S1 * p.iX1;
    double S2 = div_fast1(S1 + S1, __fma_rn(1.0 - s1, b - 1.0, b + 1.0));
    o.AE = S1 - S2;
    // percolation: S = S2 / (1 + (S2/X1)^3)^(1/3)
    double sr = S2 * p.iX1;
    double c = __fma_rn(sr * sr, sr, 1.0);
    double y = rcbrt_fast(c);
    if (f32exp) y = __fma_rn(KC.nf32_delta * (double)__logf(__double2float_rn(c)), y, y);
    S = S2 * y;
    // routing store with exchange: R2 = X2 (R + P1 + P2), Q = R2^2/(R2 + 60)
    double R2 = p.X2 * (R + (P1 + (S2 - S)));
    double Q = div_fast1(R2 * R2, R2 + 60.0);
    R = R2 - Q;
    o.Pc = Pc;
    o.Q = Q;
    return o;
}

// exp in table-index units: u = x * 2048/ln2 >= 0, returns exp(x)/2.  k = rint(u) by the magic add, f = u - k
// exactly (|f| <= 1/2), exp(f L) = 1 + f L + (f L)^2/2 + (f L)^3/6 with L = ln2/2048 folded into the
// coefficients (remainder 3.4e-17): 2 DADD-class + 1 exact DADD + 3 DFMA + 1 DMUL, no separate reduction FMA.
__device__ __forceinline__ double exp_idx2048_half(double u, const double *__restrict__ tab)
{
    double t = __dadd_rn(u, KC.magic);
    int ki = __double2loint(t);
    double kd = __dadd_rn(t, KC.nmagic);
    double f = __dsub_rn(u, kd);
    double p = __fma_rn(f, KC.e3, KC.e2);
    p = __fma_rn(p, f, KC.e1);
    p = __fma_rn(p, f, 1.0);
    double v = __dmul_rn(p, tab[ki & 2047]);
    return __hiloint2double(__double2hiint(v) + (((ki >> 11) - 1) << 20), __double2loint(v));
}

// min(u, Uc) on the high word by the integer ALU (see clamp26).  Uc = 76820.6875 is the first double with a
// zero low word at or above 26 * 2048/ln2 = 76820.6255: the clamped argument is 26 + 2.1e-5, which moves
// tanh(13) = 1 - 1.0e-11 by 2e-16.
__device__ __forceinline__ double clamp_u26(double u)
{
    return __double2hiint(u) >= 0x40F2C14B ? 76820.6875 : u;
}

// Calibration-mode step (shipped).  Two reductions of the per-month work:
//  * the forcing correction is reduced to its decision: n = rint(10 f x) is the integer R's round() settles
//    on (ties and anything within 1e-3 of one still go through the exact candidate comparison); the corrected
//    value n/10 is only ever consumed inside 1e-16-tolerant arithmetic here (it is not an output in this
//    mode), so it enters as n * 0.1 inside an FMA and the exp argument as n * g, g = 0.2/X1 * 2048/ln2;
//  * the production store is carried as s = S/X1: s1 = (s ap + am)/(ap + s am), s2 = s1/(bp + (1 - s1) bm),
//    s' = s2 (1 + s2^3)^(-1/3), with am = (a-1)/2, ap = (a+1)/2 (a = e^{2P/X1}; the table gives a/2 directly),
//  * the routing store is carried as rho = R/X1 as well: rho2 = X2 (rho + Pc/X1 + (s - s1) + (s2 - s')),
//    q = rho2^2/(rho2 + 60/X1) = Q/X1, rho' = rho2 - q; the caller scales q by area * X1 once per sum.
template <bool F32EXP>
__device__ __forceinline__ double step_fast4(const CellPar &p, double Praw, double Eraw, double &s, double &rho,
                                             const double *__restrict__ tab)
{
    double xp10 = __dmul_rn(p.fp10, Praw), xe10 = __dmul_rn(p.fe10, Eraw);
    // rint in one conversion-unit instruction (FRND.F64).  thr >= 0 implies |10 f x| < 1e9 (forcing bounded at
    // upload, factors at parameter load), where ulp(10 f x) is far inside the tie window
    double np_ = rint(xp10), ne_ = rint(xe10);
    double fp_ = __dsub_rn(xp10, np_), fe_ = __dsub_rn(xe10, ne_);
    if (!(fabs(fp_) < p.thr) || !(fabs(fe_) < p.thr)) {   // near a decimal tie (or NaN, or unbounded input): exact path
        np_ = __dmul_rn(r_round1_exact(__dmul_rn(p.fp, Praw)), 10.0);
        ne_ = __dmul_rn(r_round1_exact(__dmul_rn(p.fe, Eraw)), 10.0);
    }
    double a2 = exp_idx2048_half(clamp_u26(np_ * p.g), tab);
    double b2 = exp_idx2048_half(clamp_u26(ne_ * p.g), tab);
    double am = a2 - 0.5, ap = a2 + 0.5;
    double s1 = div_fast3(__fma_rn(s, ap, am), __fma_rn(s, am, ap));
    double bm = b2 - 0.5, bp = b2 + 0.5;
    double s2 = div_fast3(s1, __fma_rn(1.0 - s1, bm, bp));
    double c = __fma_rn(s2 * s2, s2, 1.0);
    double y = rcbrt_fast(c);
    if (F32EXP) y = __fma_rn(KC.nf32_delta * (double)__logf(__double2float_rn(c)), y, y);
    double sn = s2 * y;
    double d = (s - s1) + (s2 - sn);
    double rho2 = p.X2 * (__fma_rn(np_, p.cP, rho) + d);
    double q = div_fast3(rho2 * rho2, rho2 + p.kappa);
    rho = rho2 - q;
    s = sn;
    return q;
}

__device__ __forceinline__ StepOut step_literal(const CellPar &p, double Praw, double Eraw, double &S,
                                                double &R, bool f32exp)
{
    StepOut o;
    double Pc = r_round1_exact(__dmul_rn(p.fp, Praw));
    double Ec = r_round1_exact(__dmul_rn(p.fe, Eraw));
    double WS = Pc / p.X1;
    if (WS > 13.0) WS = 13.0;
    double ev = exp(WS);
    double T = (ev - 1.0 / ev) / (ev + 1.0 / ev);
    double Sr = S / p.X1;
    double S1 = (S + p.X1 * T) / (1.0 + Sr * T);
    double P1 = Pc + S - S1;
    WS = Ec / p.X1;
    if (WS > 13.0) WS = 13.0;
    ev = exp(WS);
    T = (ev - 1.0 / ev) / (ev + 1.0 / ev);
    double S2 = S1 * (1.0 - T) / (1.0 + (1.0 - S1 / p.X1) * T);
    o.AE = S1 - S2;
    Sr = S2 / p.X1;
    Sr = Sr * Sr * Sr + 1.0;
    S = S2 / pow(Sr, f32exp ? (double)(1.0f / 3.0f) : 1.0 / 3.0);
    double P2 = S2 - S;
    double P3 = P1 + P2;
    double R1 = R + P3;
    double R2 = p.X2 * R1;
    double Q = R2 * R2 / (R2 + 60.0);
    R = R2 - Q;
    o.Pc = Pc;
    o.Q = Q;
    return o;
}
