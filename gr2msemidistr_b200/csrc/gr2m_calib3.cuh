// gr2m_calib3.cuh -- calibration-mode recursion with per-set power tables (included by gr2m.cu).
//
// In calibration mode the corrected forcing is n/10 with n = rint(10 f x) an integer (R's round(f*x, 1),
// R/Run_GR2MSemiDistr.R:101-102), and with one region all lanes of a warp share X1, so both exponentials of a
// month are exp(n c) with ONE c = 0.2/X1 per warp:  exp(n c) = exp(256 i c) * exp(j c),  n = 256 i + j.
// Each warp builds T1h[i] = exp(min(256 i c, 26))/2 and T2[j] = exp(j c) (2 x 256 doubles, 4 KiB) once and a
// month's exp becomes two shared-memory reads whose product enters the store updates through FMAs, instead of 8
// FP64-pipe instructions and a clamp (airGR's min(WS, 13) is an integer min on n).  The rounding decision itself is
// made in fixed point: ONE FMA per forcing value leaves n in the high word and the distance to the next decimal tie
// in the low word of a double (step_tab_fx); a chunk in which any lane came within 2^-26 of a tie is repeated with
// R's exact candidate comparison.  A tile where some lane could see n outside [0, 65535] (huge forcing with a large
// X1, negative forcing or factors, unbounded input) or whose set has X1 < 1 (c > 0.2: the clamped argument n_clamp c
// could overshoot 26 by more than 0.2) takes step_fast4, the general formulation.
// Outlet sums go to global memory by fire-and-forget FP64 reductions (RED.ADD.F64): every (set, month) word is
// only ever updated by one lane of one warp, in program order, so the sums are deterministic; this frees the
// 30 KiB of per-CTA accumulators and lets the tables in.
#pragma once

#ifndef GR2M_WHATIF
#define GR2M_WHATIF 0     /* timing experiments only (scripts/whatif.sh): drop one instruction group, results are wrong */
#endif
constexpr int PT_N = 256;                 // entries per table level
constexpr int PT_WARP_DBL = 2 * PT_N;     // doubles per warp
constexpr unsigned CALIB4_ALIGN_PAD = 4096;   // k_gr2m_calib4 aligns its tables to 4 KiB (entry offsets are ORed into the base)
constexpr int CALIB4_NST = 2;                 // forcing stages in flight (2, 3 and 4 measured alike; 2 leaves room for the tables)
constexpr int RCB_DBL = 512;                  // cube-root seed table: 512 float2 = 4 KiB

// per cell: frange[2c] = max_t P, frange[2c+1] = max_t E; +inf when any value is negative or not finite
__global__ void k_cell_range(const double *__restrict__ F, int ncell, int ntime, int ntime_pad, double *__restrict__ frange)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ncell) return;
    const double *f = F + (size_t)(cell >> 5) * ntime_pad * 64 + (cell & 31);
    double pm = 0.0, em = 0.0;
    bool bad = false;
    for (int t = 0; t < ntime; ++t) {
        double p = f[(size_t)t * 64], e = f[(size_t)t * 64 + 32];
        bad |= !(p >= 0.0) || !(e >= 0.0);
        pm = fmax(pm, p);
        em = fmax(em, e);
    }
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    frange[2 * cell] = bad ? inf : pm;
    frange[2 * cell + 1] = bad ? inf : em;
}

// Table-mode month step: same algebra as step_fast4 (gr2m_math.cuh), exponentials from the warp's tables.
// t1 / t2: shared-memory byte addresses of the warp's T1h / T2; nclamp: smallest n with n c >= 26.
// exp(n c)/2 from the warp's tables: T1h[n >> 8] * T2[n & 255]; addresses by one shift / mask and one multiply-add each
__device__ __forceinline__ void pow_tab_hl(int n, uint32_t t1, uint32_t t2, double &h, double &l)
{
    uint32_t a1, a2;
    asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(a1) : "r"((uint32_t)n >> 8), "r"(t1));
    asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(a2) : "r"((uint32_t)n & (PT_N - 1)), "r"(t2));
    asm("ld.shared.f64 %0, [%1];" : "=d"(h) : "r"(a1));
    asm("ld.shared.f64 %0, [%1];" : "=d"(l) : "r"(a2));
}

// The part of a month that does not depend on the rounding decision's branch: stores, percolation, routing.
// am = a2 - 1/2, ap = a2 + 1/2 (a2 = e^{2P/X1}/2), Y = 1/2 - b2 (b2 = e^{2E/X1}/2), each one FMA on the table product.
// Evaporation: s2 = s1 / (2 b2 + s1 (1/2 - b2)) with 2 b2 = 1 - 2 Y, so the denominator is fma(s1, Y, fma(-2, Y, 1)):
// three FMAs from the table entries (was a product, two additions and an FMA), one of them after s1; both terms are
// positive, nothing cancels.  ctab: shared address of the cube-root seed table (rcbrt_fast_tab).
template <bool F32EXP>
__device__ __forceinline__ double step_tab_core2(const CellPar &p, double np_, double am, double ap, double Y, double &s, double &rho,
                                                 uint32_t ctab)
{
    double s1 = div_fast3q(__fma_rn(s, ap, am), __fma_rn(s, am, ap));
    double s2 = div_fast3q(s1, __fma_rn(s1, Y, __fma_rn(-2.0, Y, 1.0)));
    double c = __fma_rn(s2 * s2, s2, 1.0);
#if GR2M_WHATIF == 1
    double y = __fma_rn(c, -0.2, 1.0);            // no cube root: -5 FP64, -4 XU, -1 FMUL
#elif GR2M_WHATIF == 7
    double y = rcbrt_fast_poly(c);               // degree-6 FP32 polynomial seed (round 2)
#else
    double y = rcbrt_fast_tab(c, ctab);          // c = 1 + s2^3 in [1, 2]
#endif
    if (F32EXP) y = __fma_rn(KC.nf32_delta * (double)__logf(__double2float_rn(c)), y, y);
    double sn = s2 * y;
    double d = (s - s1) + (s2 - sn);
    double rho2 = p.X2 * (__fma_rn(np_, p.cP, rho) + d);
#if GR2M_WHATIF == 5
    double q = rho2 * rho2 * 0.01;               // no third division: -4 FP64, -1 XU, -1 MOV
#else
    double q = div_fast3q(rho2 * rho2, rho2 + p.kappa);
#endif
    rho = rho2 - q;
    s = sn;
    return q;
}

// RQS: round(area Q / den, 1) decided by rint.  y = area X1 q (the caller's product), c10 = 10 / den[t]: x10 = 10 x up to
// one rounding; away from a decimal tie (|x10 - rint(x10)| < GR2M_TIE_THR) and below 1e9 the result is the correctly
// rounded n / 10 (div10_rn), which is what R's candidate comparison returns there; anything else (ties, huge values,
// NaN) raises `tie` and the caller repeats the chunk through r_round1.
__device__ __forceinline__ double round_qs_fast(double y, double c10, bool &tie)
{
    const double x10 = __dmul_rn(y, c10);
    const double n = __dadd_rn(__dadd_rn(x10, KC.magic), KC.nmagic);
    tie |= !(fabs(__dsub_rn(x10, n)) < KC.tie_thr) | !(x10 < 1e9);
    return div10_rn(n);
}

// Table lookups of the fixed-point variant: `hi` is the high word of t = x + FX_C (below), whose bits 0..15 hold
// n = floor(x + 1/2 + w); each table level sits at a 2 KiB-aligned shared address, so an entry's address is two logic
// instructions (shift, and-or) per level and n never has to exist as an integer of its own.
__device__ __forceinline__ void pow_tab_ld(uint32_t hi, uint32_t t1, uint32_t t2, double &h, double &l)
{
    const uint32_t a1 = ((hi >> 5) & 0x7f8u) | t1, a2 = ((hi << 3) & 0x7f8u) | t2;
    asm("ld.shared.f64 %0, [%1];" : "=d"(h) : "r"(a1));
    asm("ld.shared.f64 %0, [%1];" : "=d"(l) : "r"(a2));
}

// 1.5 * 2^20 + 1/2 + 2^-26: for 0 <= x < 65535, t = x + FX_C lies in [2^20, 2^21), where one ulp is 2^-32: the low
// word of t is the fraction of x + 1/2 + 2^-26 in units of 2^-32 and the low 16 bits of the high word are its integer
// part.  A low word below 2^7 (= 2 * 2^-26) means x is within 2^-26 = 1.5e-8 of a decimal tie (the FMA's own rounding,
// 2^-33, cannot hide one: a sum rounded up to the next integer has a zero low word); everywhere else
// floor(x + 1/2 + 2^-26) = rint(x) = the integer R's round(f x, 1) settles on (ulp(x) <= 1.5e-11 here).
#define GR2M_FX_C 1572864.500000014901161193847656250
#define GR2M_FX_TIE 128u

// Table-mode month, branch-free, rounding decision in fixed point (shipped): per forcing value one
// FP64-pipe instruction, fma(10 f, x, FX_C) (was DMUL, DADD, DSETP + FRND and F2I on the conversion unit), the clamp
// min(n, nclamp) as an integer min on the high word, the tie test as a running unsigned min of the low words (`tmin`,
// compared with GR2M_FX_TIE once per chunk by the caller); rint(10 f P) as a double, which the routing store needs, is
// one integer-to-double conversion of the high word's low 16 bits.
// CLAMP = false: the caller knows that no month of the tile pair reaches nclamp (k_cell_range's maxima), the two integer
// mins are left out.
template <bool F32EXP, bool CLAMP>
__device__ __forceinline__ double step_tab_fx(const CellPar &p, double Praw, double Eraw, double &s, double &rho,
                                              uint32_t t1, uint32_t t2, uint32_t ctab, uint32_t hiclamp, uint32_t &tmin)
{
    // t = 10 f x + FX_C in ONE rounding (to 2^-32): integer part and tie distance of the exact product
    const double tp = __fma_rn(p.fp10, Praw, GR2M_FX_C), te = __fma_rn(p.fe10, Eraw, GR2M_FX_C);
    const uint32_t hp = (uint32_t)__double2hiint(tp), he = (uint32_t)__double2hiint(te);
    const double np_ = __int2double_rn((int)(hp & 0xffffu));        // rint(10 f P) for the routing store (I2F)
    tmin = min(tmin, min((uint32_t)__double2loint(tp), (uint32_t)__double2loint(te)));
    double h, l;
    pow_tab_ld(CLAMP ? min(hp, hiclamp) : hp, t1, t2, h, l);
    const double am = __fma_rn(h, l, -0.5), ap = __fma_rn(h, l, 0.5);
    pow_tab_ld(CLAMP ? min(he, hiclamp) : he, t1, t2, h, l);
    return step_tab_core2<F32EXP>(p, np_, am, ap, __fma_rn(-h, l, 0.5), s, rho, ctab);
}

// Table-mode month, branch-free (round-2 form, kept as an experiment switch GR2M_CALIB_MAGIC=0|1): the rounding
// decision is rint(10 f x); `tie` collects the lanes that were within
// 1e-8 of a decimal tie (|10 f x| < 2^16 here, ulp <= 1.5e-11: R's candidate comparison cannot disagree with rint
// outside that window).  The caller repeats the 8-month chunk with step_tab when any lane raised it.
template <bool F32EXP, int MAGIC>
__device__ __forceinline__ double step_tab_nb(const CellPar &p, double Praw, double Eraw, double &s, double &rho,
                                              uint32_t t1, uint32_t t2, uint32_t ctab, int nclamp, bool &tie)
{
    double xp10 = __dmul_rn(p.fp10, Praw), xe10 = __dmul_rn(p.fe10, Eraw);
    double np_, ne_;
    int ip, ie;
    if (MAGIC) {      // rint and the integer from one magic add each (FP64 pipe) instead of FRND + F2I (XU pipe)
        double tp = __dadd_rn(xp10, KC.magic), te = __dadd_rn(xe10, KC.magic);
        ip = __double2loint(tp); ie = __double2loint(te);
        np_ = __dadd_rn(tp, KC.nmagic); ne_ = __dadd_rn(te, KC.nmagic);
    } else {
        np_ = rint(xp10); ne_ = rint(xe10);
        ip = __double2int_rn(xp10); ie = __double2int_rn(xe10);
    }
    const bool okp = fabs(__dsub_rn(xp10, np_)) < KC.tie_thr_tab, oke = fabs(__dsub_rn(xe10, ne_)) < KC.tie_thr_tab;
    tie |= !(okp & oke);
    double h, l, am, ap;
    pow_tab_hl(min(ip, nclamp), t1, t2, h, l);
    am = __fma_rn(h, l, -0.5); ap = __fma_rn(h, l, 0.5);
    pow_tab_hl(min(ie, nclamp), t1, t2, h, l);
    return step_tab_core2<F32EXP>(p, np_, am, ap, __fma_rn(-h, l, 0.5), s, rho, ctab);
}

// Same month with the exact rounding path taken by the lanes that need it (R's candidate comparison).
template <bool F32EXP>
__device__ __forceinline__ double step_tab(const CellPar &p, const double *__restrict__ par, double Praw, double Eraw,
                                           double &s, double &rho, uint32_t t1, uint32_t t2, uint32_t ctab, int nclamp)
{
    double xp10 = __dmul_rn(p.fp10, Praw), xe10 = __dmul_rn(p.fe10, Eraw);
    double np_ = rint(xp10), ne_ = rint(xe10);
    int ip = __double2int_rn(xp10), ie = __double2int_rn(xe10);
    const bool okp = fabs(__dsub_rn(xp10, np_)) < KC.tie_thr_tab, oke = fabs(__dsub_rn(xe10, ne_)) < KC.tie_thr_tab;
    if (!(okp & oke)) {            // factors re-read from memory: rare path
        np_ = __dmul_rn(r_round1_exact(__dmul_rn(__ldg(par + 2), Praw)), 10.0);
        ne_ = __dmul_rn(r_round1_exact(__dmul_rn(__ldg(par + 3), Eraw)), 10.0);
        ip = __double2int_rn(np_);
        ie = __double2int_rn(ne_);
    }
    double h, l, am, ap;                   // the same fused half terms as the branch-free pass: same bits either way
    pow_tab_hl(min(ip, nclamp), t1, t2, h, l);
    am = __fma_rn(h, l, -0.5); ap = __fma_rn(h, l, 0.5);
    pow_tab_hl(min(ie, nclamp), t1, t2, h, l);
    return step_tab_core2<F32EXP>(p, np_, am, ap, __fma_rn(-h, l, 0.5), s, rho, ctab);
}

// ---------------------------------------------------------------------------------------------------------
// k_gr2m_calib4: the same table-mode recursion with CPT cells per thread and canonical partial sums.
//  * CPT = 2: a lane carries the stores of two cells (tiles t and t+1); the two recursions are independent, so the
//    instruction stream of a warp holds two dependency chains to interleave and the per-chunk work (cross-lane
//    reduction, reductions to global memory, stage hand-over) is shared by twice the units.
//  * Outlet sums are accumulated per GROUP of a.group_tiles consecutive tiles (a function of the number of cells only,
//    calib_group_tiles): partial[group][set][month].  The groups do not depend on how a launch is cut into blocks, so a set's sums are the same bits whatever the batch size,
//    the number of ranks or the SM count (k_objective adds the groups in index order).
// ---------------------------------------------------------------------------------------------------------

//  * MAGIC = 2 (shipped): rounding decision in fixed point (step_tab_fx); 0 / 1: the round-2 forms (step_tab_nb).
//  * RQS (routed objective at a gauge, gr2m_objective_routed_batch): every subbasin's discharge area Q / den[t] is rounded
//    to 0.1 with R's rule BEFORE it is summed (Run:229-235 rounds QS, Routing then accumulates the rounded values); the
//    cells outside the gauge's upstream set come in with area 0.  The decision is rint(10 x) wherever 10 x is at
//    least 1e-5 away from a decimal tie and below 1e9, anything else raises the chunk's redo flag and goes through
//    r_round1 month by month.  conv10[t] = 10 / den[t] sits in shared memory.
//  * MASKED (same caller): a tile pair whose 64 cells all carry area 0 is consumed without arithmetic.
template <bool F32EXP, int CTAS, int CPT, int MAGIC, bool RQS = false, bool MASKED = false>
__global__ void __launch_bounds__(SPB * 32, CTAS) k_gr2m_calib4(CalibArgs a)
{
    constexpr int NST = CALIB4_NST, STG = CPT * STAGE2_DBL;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the tables come first, at a 4 KiB-aligned shared address (pow_tab_ld / rcbrt_fast_tab OR the entry offset into
    // the base); the launch asks for CALIB4_ALIGN_PAD more bytes than the layout needs
    unsigned char *smem_al = smem_raw + ((0u - smem_u32(smem_raw)) & (CALIB4_ALIGN_PAD - 1));
    double *rcb = reinterpret_cast<double *>(smem_al);               // RCB_DBL: cube-root seed table (float2 entries)
    double *ptab = rcb + RCB_DBL;                                    // SPB * PT_WARP_DBL
    double *stage = ptab + SPB * PT_WARP_DBL;                        // NST * STG
    double *saved = stage + NST * STG;                               // 2 * CPT * SPB * 32: stores at the start of a chunk
    uint64_t *full = reinterpret_cast<uint64_t *>(saved + 2 * CPT * SPB * 32);   // NST
    int *left = reinterpret_cast<int *>(full + NST);                 // NST: warps that have left the stage
    double *conv10 = reinterpret_cast<double *>(left + NST + (NST & 1));         // RQS only: a.ntime_pad doubles

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sg = blockIdx.x % a.nsg, cs = blockIdx.x / a.nsg;
    const int set = sg * SPB + warp;
    const bool set_ok = set < a.nsets;
    const int tile0 = cs * a.tiles_per_split;                        // a multiple of a.group_tiles
    const int tile1 = min(a.ntiles, tile0 + a.tiles_per_split);
    const int nchunk = a.ntime_pad / TM2;
    const int npair = (tile1 - tile0 + CPT - 1) / CPT;
    const int total = npair * nchunk;

    if (tid < NST) left[tid] = 0;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < RCB_DBL; i += SPB * 32) reinterpret_cast<float2 *>(rcb)[i] = g_rcbrt_lin[i];
    if (RQS)
        for (int i = tid; i < a.ntime_pad; i += SPB * 32) conv10[i] = i < a.ntime ? 10.0 * a.tconv[i] : 0.0;
    const uint32_t ctab = smem_u32(rcb);
    __syncthreads();
    // stage gi = 8 months (chunk gi % nchunk) of the CPT tiles of pair gi / nchunk: one 4 KiB bulk copy per tile
    auto issue = [&](int gi) {
        const int pi = gi / nchunk, ch = gi - pi * nchunk, st = gi % NST;
        const int tA = tile0 + pi * CPT, nvalid = min(CPT, tile1 - tA);
        mbar_arrive_expect_tx(&full[st], nvalid * STAGE2_DBL * 8);
        for (int c = 0; c < nvalid; ++c)
            bulk_g2s(stage + st * STG + c * STAGE2_DBL, a.F + ((size_t)(tA + c) * a.ntime_pad + ch * TM2) * 64, STAGE2_DBL * 8,
                     &full[st]);
    };
    if (tid == 0)
        for (int gi = 0; gi < NST && gi < total; ++gi) issue(gi);

    // one region: every lane of the warp has the set's parameters (only the area differs)
    const double *par = a.params + (size_t)(set_ok ? set : 0) * 4;
    CellPar p = load_par(par, 1, 0, 0.0, a.forcing_ok);
    // smallest n whose argument n c reaches the clamp 26 (airGR: min(P/X1, 13)); tables hold exp(min(., 26))
    const double U26 = 0x1.2c14a02335a6fp+16;                       // 26 * 2048 / ln 2
    const double ncd = ceil(U26 / p.g);
    const int nclamp = ncd < 1073741824.0 ? (int)ncd : 1073741824;   // NaN parameters: results are NaN anyway
    double *mytab = ptab + warp * PT_WARP_DBL;
    for (int i = lane; i < PT_N; i += 32) {
        mytab[i] = exp_idx2048_half(clamp_u26((double)(i * PT_N) * p.g), g_exp2_tab2048);            // T1h
        mytab[PT_N + i] = 2.0 * exp_idx2048_half(clamp_u26((double)i * p.g), g_exp2_tab2048);        // T2
    }
    __syncwarp();
    const uint32_t t1 = smem_u32(mytab), t2 = t1 + PT_N * 8;
    const bool set_tab = p.thr >= 0.0 && p.X1 >= 1.0 && p.fp10 >= 0.0 && p.fe10 >= 0.0;
    // high word of min(nclamp, 65535) + FX_C: the clamp is an unsigned min on the high word of x + FX_C
    const uint32_t hiclamp = (uint32_t)__double2hiint((double)min(nclamp, 65535) + GR2M_FX_C);
    const double nclamp_m1 = (double)nclamp - 1.0;                   // 10 f x below it: floor(10 f x + 1/2 + 2^-26) < nclamp

    const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const bool writer = set_ok && (lane & 3) == 0;
    int g = 0, nredo = 0;
    bool branchy = false;
    for (int pi = 0; pi < npair; ++pi) {
        const int tA = tile0 + pi * CPT;
        double *myout = a.partial + ((size_t)(tA / a.group_tiles) * a.nsets + (set_ok ? set : 0)) * a.ntime;
        double aX1[CPT], S[CPT], R[CPT];
        int soff[CPT];                  // a missing last tile re-reads the first one's forcing with area 0
        bool lane_tab = set_tab, lane_noclamp = MAGIC == 3;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const bool has = tA + c < tile1;
            const int cell = (tA + c) * 32 + lane;
            const bool cell_ok = has && cell < a.ncell;
            aX1[c] = (cell_ok ? a.area[cell] : 0.0) * p.X1;
            soff[c] = (has ? c : 0) * STAGE2_DBL + lane;
            // table mode needs 0 <= 10 f x < 65535 for every month of every lane (+inf marks negative / non-finite forcing)
            if (cell_ok) {
                const double xpm = p.fp10 * a.frange[2 * cell], xem = p.fe10 * a.frange[2 * cell + 1];
                lane_tab = lane_tab && (xpm < 65535.0) && (xem < 65535.0);
                lane_noclamp = lane_noclamp && (xpm < nclamp_m1) && (xem < nclamp_m1);
            }
            S[c] = 0.3;                                               // stores divided by X1; airGR starts at 0.3 X1, 0.5 X2
            R[c] = 0.5 * p.X2 * p.iX1;
        }
        const bool use_tab = __all_sync(0xffffffffu, lane_tab);
        const bool no_clamp = MAGIC == 3 && __all_sync(0xffffffffu, lane_noclamp);   // no month of the pair reaches nclamp
        // a pair whose 64 cells all carry area 0 adds nothing to the sums (the routed objective masks the subbasins
        // outside the gauge's upstream set this way): its stages are consumed without arithmetic
        bool lane_live = !MASKED;
#pragma unroll
        for (int c = 0; c < CPT; ++c) lane_live = lane_live || !(aX1[c] == 0.0);
        const bool pair_live = !MASKED || __any_sync(0xffffffffu, lane_live);
        for (int ch = 0; ch < nchunk; ++ch, ++g) {
            const int st = g % NST;
            mbar_wait(&full[st], (g / NST) & 1);
            if (set_ok && pair_live) {
                const double *sp = stage + st * STG;
                bool redo = true;
                if (use_tab && !branchy) {
                    // branch-free pass over the 8 months; repeated below if a lane sat on a rounding tie
#pragma unroll
                    for (int c = 0; c < CPT; ++c) {
                        saved[(2 * c) * SPB * 32 + tid] = S[c];
                        saved[(2 * c + 1) * SPB * 32 + tid] = R[c];
                    }
                    bool tie = false;
                    uint32_t tmin = 0xffffffffu;
                    double v[8];
#pragma unroll
                    if (MAGIC == 3 && no_clamp) {
#pragma unroll
                        for (int m = 0; m < 8; ++m) {
#pragma unroll
                            for (int c = 0; c < CPT; ++c) {
                                const double Pr = sp[soff[c] + m * 64], Er = sp[soff[c] + m * 64 + 32];
                                const double qc = step_tab_fx<F32EXP, false>(p, Pr, Er, S[c], R[c], t1, t2, ctab, hiclamp, tmin);
                                if (RQS) {
                                    const double qs = round_qs_fast(aX1[c] * qc, conv10[ch * TM2 + m], tie);
                                    v[m] = c == 0 ? qs : v[m] + qs;
                                } else {
                                    v[m] = c == 0 ? aX1[0] * qc : __fma_rn(aX1[c], qc, v[m]);
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int m = 0; m < 8; ++m) {
#pragma unroll
                            for (int c = 0; c < CPT; ++c) {
                                const double Pr = sp[soff[c] + m * 64], Er = sp[soff[c] + m * 64 + 32];
                                const double qc = MAGIC >= 2 ? step_tab_fx<F32EXP, true>(p, Pr, Er, S[c], R[c], t1, t2, ctab, hiclamp, tmin)
                                                             : step_tab_nb<F32EXP, MAGIC>(p, Pr, Er, S[c], R[c], t1, t2, ctab, nclamp, tie);
                                if (RQS) {
                                    const double qs = round_qs_fast(aX1[c] * qc, conv10[ch * TM2 + m], tie);
                                    v[m] = c == 0 ? qs : v[m] + qs;
                                } else {
                                    v[m] = c == 0 ? aX1[0] * qc : __fma_rn(aX1[c], qc, v[m]);
                                }
                            }
                        }
                    }
                    if (MAGIC >= 2) tie |= tmin < GR2M_FX_TIE;
                    redo = __any_sync(0xffffffffu, tie);
                    if (!redo) {
                        double s = warp_reduce8(v, lane);
                        const int t = ch * TM2 + ridx;
                        if (writer && t < a.ntime) atomicAdd(myout + t, RQS ? s : s * __ldg(a.tconv + t));
                    } else {
#pragma unroll
                        for (int c = 0; c < CPT; ++c) {
                            S[c] = saved[(2 * c) * SPB * 32 + tid];
                            R[c] = saved[(2 * c + 1) * SPB * 32 + tid];
                        }
                        // parameter sets whose factors make ties common (e.g. three-decimal factors on one-decimal
                        // forcing: one value in a thousand) go month by month for the rest of the launch
                        nredo += 16;
                        branchy = nredo > g + 64;
                    }
                }
                if (redo) {
                    // month by month (ties, or the general formulation); the butterfly adds in warp_reduce8's order
#pragma unroll 1
                    for (int m = 0; m < 8; ++m) {
                        double q = 0.0;
#pragma unroll
                        for (int c = 0; c < CPT; ++c) {
                            const double Pr = sp[soff[c] + m * 64], Er = sp[soff[c] + m * 64 + 32];
                            const double qc = use_tab ? step_tab<F32EXP>(p, par, Pr, Er, S[c], R[c], t1, t2, ctab, nclamp)
                                                      : step_fast4<F32EXP>(p, Pr, Er, S[c], R[c], g_exp2_tab2048);
                            if (RQS) {
                                const int tt = ch * TM2 + m;
                                const double x = __dmul_rn(aX1[c] * qc, tt < a.ntime ? __ldg(a.tconv + tt) : 0.0);
                                const double qs = r_round1(x, fabs(x) < 1e8 ? GR2M_TIE_THR : -1.0);
                                q = c == 0 ? qs : q + qs;
                            } else {
                                q = c == 0 ? aX1[0] * qc : __fma_rn(aX1[c], qc, q);
                            }
                        }
                        q = warp_sum_f64(q);
                        const int t = ch * TM2 + m;
                        if (lane == 0 && t < a.ntime) atomicAdd(myout + t, RQS ? q : q * __ldg(a.tconv + t));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                // this warp's reads of the stage are complete; release / acquire on left[st] orders them before the refill
                __threadfence_block();
                if (atomicAdd(&left[st], 1) == SPB - 1) {
                    __threadfence_block();
                    left[st] = 0;
                    if (g + NST < total) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        issue(g + NST);
                    }
                }
            }
        }
    }
}
