// rmt_kernels.cu — hand-written FP64 kernels for the pseudo-homogeneous packed-bed
// reactor models N1 (steady-state axial ODE) and N2 (dynamic, method of lines).
//
// This file is the model-independent half of the NVRTC translation unit.  The
// model-dependent half, "rmt_model.cuh", is generated per (components,
// reactions, kinetics) by rmt_app_b200/codegen.py and provides the sizes, the
// component literals, the traced kinetics `rmt_rates` / `rmt_rates_jac` and the
// Rosenbrock tableau.  Target: sm_100a (B200).  No tensor cores: the path is not
// a contraction; the bound is the FP64 vector pipe (fused integrator) or HBM
// (stand-alone RHS / Jacobian kernels).
//
// Reference arithmetic being restated (file:line relative to
// /root/reference/PyREMOT/): setup docs/pbHomoReactor.py:2744-2852 (N1) and
// :3370-3507 (N2); RHS modelEquationN1 :3017-3314, modelEquationN2 :3706-4134;
// thermo docs/rmtThermo.py:16-101,258-369; utilities docs/rmtUtility.py:405-496;
// viscosity docs/gasTransPor.py:137-274; un-scaling solvers/solResultAnalysis.py:191-301.
//
// Data layout: structure-of-arrays with the instance index fastest everywhere
// (inputs [row][B], constants [row][B], states [var][B] or [var][node][B],
// outputs [point][var][B]) so that a warp's lanes touch consecutive doubles.

// Division.  IEEE `a/b` on sm_100a is a ~14-instruction sequence with a branch to a slow path
// (BSSY/BSYNC + call) — with ~35 divisions per RHS that is a fifth of the instruction stream and a
// steady source of instruction-fetch stalls.  The hot code therefore multiplies by a refined
// reciprocal: MUFU.RCP64H seed (~2^-20) + one third-order correction (3 FMAs) -> <= 2 ulp, branch free.  Zero / Inf /
// NaN operands still end in Inf / NaN, which the integrator's domain guard turns into a rejected
// step.  -DRMT_EXACT_DIV=1 restores IEEE division everywhere.
#ifndef RMT_EXACT_DIV
#define RMT_EXACT_DIV 0
#endif
__device__ __forceinline__ double rmt_rcp(const double x)
{
#if RMT_EXACT_DIV
    return 1.0/x;
#else
    // seed r0 = (1/x)(1 - e), |e| <~ 2^-20; one third-order step r0 (1 + e + e^2) = (1/x)(1 - e^3): three dependent
    // FMAs instead of the four of two Newton steps, |e|^3 < 2^-58
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
#endif
}
// |x| limited to the double whose high word is `hi_limit` (the low word of x is kept: limit <= |result| < limit*(1 + 2^-20)),
// sign kept; NaN / Inf come out as +-limit like the fmin(fmax()) pair it replaces.  Integer pipe only: DMNMX does not
// exist on sm_100a, fmin/fmax of doubles are a DSETP + selects each, and the DSETP occupies an FP64 issue slot.
__device__ __forceinline__ double rmt_clamp_abs(const double x, const int hi_limit)
{
    const int hx = __double2hiint(x);
    return __hiloint2double((hx & 0x80000000) | min(hx & 0x7fffffff, hi_limit), __double2loint(x));
}
#if RMT_EXACT_DIV
#define RMT_DIV(a, b) ((a)/(b))
#else
#define RMT_DIV(a, b) ((a)*rmt_rcp(b))
#endif

// Transcendentals.  libdevice's exp / log / sqrt each contain a branch to a special-case path; besides the
// extra instructions, those branches cut the instruction stream into basic blocks, so the independent
// Arrhenius / equilibrium exponentials of a kinetics section (all functions of T only) cannot be interleaved
// by the scheduler.  The versions below are branch free (select instead of branch) and accurate to ~1 ulp
// on the normal range: exp clamps its argument to +-708 (finite huge/tiny instead of Inf/0), log returns NaN outside
// the positive normal range, sqrt NaN for negative arguments.  -DRMT_EXACT_MATH=1 uses libdevice everywhere.
#ifndef RMT_EXACT_MATH
#define RMT_EXACT_MATH 0
#endif
// (Literal coefficients on purpose: fetching them from the constant bank was measured — no gain for the
// integrator, and the one-shot RHS kernel lost 25 % to the extra constant-load latency.)
// 1 = Estrin evaluation of the exp polynomial (13 FMA + 3 MUL, depth 5) instead of Horner (14 FMA, depth 14)
#ifndef RMT_EXP_ESTRIN
#define RMT_EXP_ESTRIN 0
#endif
__device__ __forceinline__ double rmt_exp_reduced(const double r, const int k)
{
    // exp(r) for |r| <= ln2/2 by the degree-13 Taylor polynomial (truncation 4e-18), times 2^k
#if RMT_EXP_ESTRIN
    const double r2 = r*r, r4 = r2*r2, r8 = r4*r4;
    const double a0 = fma(r, 1.0, 1.0);
    const double a1 = fma(r, 0.16666666666666666, 0.5);
    const double a2 = fma(r, 0.008333333333333333, 0.041666666666666664);
    const double a3 = fma(r, 0.0001984126984126984, 0.001388888888888889);
    const double a4 = fma(r, 2.7557319223985893e-06, 2.48015873015873e-05);
    const double a5 = fma(r, 2.505210838544172e-08, 2.755731922398589e-07);
    const double a6 = fma(r, 1.6059043836821613e-10, 2.08767569878681e-09);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double c0 = fma(b1, r4, b0), c1 = fma(a6, r4, b2);
    const double pe = fma(c1, r8, c0);
    return __hiloint2double(__double2hiint(pe) + (k << 20), __double2loint(pe));
#endif
    double p = 1.6059043836821613e-10;                 // 1/13!
    p = fma(p, r, 2.08767569878681e-09);               // 1/12!
    p = fma(p, r, 2.505210838544172e-08);              // 1/11!
    p = fma(p, r, 2.755731922398589e-07);              // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);             // 1/9!
    p = fma(p, r, 2.48015873015873e-05);               // 1/8!
    p = fma(p, r, 0.0001984126984126984);              // 1/7!
    p = fma(p, r, 0.001388888888888889);               // 1/6!
    p = fma(p, r, 0.008333333333333333);               // 1/5!
    p = fma(p, r, 0.041666666666666664);               // 1/4!
    p = fma(p, r, 0.16666666666666666);                // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
// Table form: exp(x) = 2^k * 2^(j/32) * exp(r), |r| <= ln2/64, so a degree-6 polynomial is enough (truncation
// 3.5e-18) — 13 FP64 instructions instead of 20 and a dependent chain of 7 instead of 14; the 32 table entries
// (two cache lines) are read through the read-only path.  Same accuracy (<= 2 ulp), but measured no faster in the
// integrator (16.09 vs 16.05 ms per 2^20 solves: the table load's latency eats the shorter chain) and 1 % slower
// in the one-shot RHS kernel, so the polynomial-only form stays the default.
#ifndef RMT_EXP_TABLE
#define RMT_EXP_TABLE 0
#endif
__device__ const double RMT_EXP2_TAB[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072,
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002};
__device__ __forceinline__ double rmt_exp_tab(const double r, const int n)
{
    // 2^(n/32) * exp(r): T*(1 + q), q = r + r^2 (1/2 + r/6 + ...), evaluated as fma(T, q, T)
    double a = fma(r, 1.3888888888888889e-03, 8.3333333333333332e-03);   // 1/720, 1/120
    a = fma(a, r, 4.1666666666666664e-02);
    a = fma(a, r, 1.6666666666666666e-01);
    a = fma(a, r, 0.5);
    const double q = fma(r, a*r, r);
    const double T = __ldg(&RMT_EXP2_TAB[n & 31]);
    const double v = fma(T, q, T);
    return __hiloint2double(__double2hiint(v) + ((n >> 5) << 20), __double2loint(v));
}
__device__ __forceinline__ double rmt_exp(double x)
{
#if RMT_EXACT_MATH
    return exp(x);
#elif RMT_EXP_TABLE
    x = rmt_clamp_abs(x, 0x40862000);                                     // |x| <= 708
    const double t = fma(x, 46.16624130844683, 6755399441055744.0);      // round(x*32/ln2) in the low word
    const double nd = t - 6755399441055744.0;
    double r = fma(nd, -0.02166084938653512, x);                          // ln2_hi/32 (exact product)
    r = fma(nd, -5.9631716539705866e-12, r);                              // ln2_lo/32
    return rmt_exp_tab(r, __double2loint(t));
#else
    x = rmt_clamp_abs(x, 0x40862000);                                     // |x| <= 708: 2^k stays a normal number
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);     // round(x*log2(e)) in the low word
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    return rmt_exp_reduced(r, __double2loint(t));
#endif
}
__device__ __forceinline__ double rmt_exp10(double x)
{
#if RMT_EXACT_MATH
    return exp10(x);
#elif RMT_EXP_TABLE
    x = rmt_clamp_abs(x, 0x40733000);                                     // |x| <= 307
    const double t = fma(x, 106.30169903639559, 6755399441055744.0);     // round(x*32*log2(10))
    const double nd = t - 6755399441055744.0;
    const double xh = x*2.302585092994045901e+00;
    double r = fma(nd, -0.02166084938653512, xh);
    r = fma(nd, -5.9631716539705866e-12, r);
    r = fma(x, -2.1707562233822494e-16, r);                               // ln10 - ln10_hi
    r += fma(x, 2.302585092994045901e+00, -xh);
    return rmt_exp_tab(r, __double2loint(t));
#else
    x = rmt_clamp_abs(x, 0x40733000);                                     // |x| <= 307
    const double t = fma(x, 3.3219280948873622, 6755399441055744.0);     // round(x*log2(10))
    const double kd = t - 6755399441055744.0;
    // r = x*ln10 - k*ln2 with ln10 and ln2 split hi/lo and the rounding error of x*ln10_hi recovered
    const double xh = x*2.302585092994045901e+00;
    double r = fma(kd, -6.93147180369123816490e-01, xh);
    r = fma(kd, -1.90821492927058770002e-10, r);
    r = fma(x, -2.1707562233822494e-16, r);                               // ln10 - ln10_hi
    r += fma(x, 2.302585092994045901e+00, -xh);
    return rmt_exp_reduced(r, __double2loint(t));
#endif
}
__device__ __forceinline__ double rmt_log(const double x)
{
#if RMT_EXACT_MATH
    return log(x);
#else
    // fdlibm e_log.c without the special cases: x = 2^k * m, m in [sqrt(1/2), sqrt(2))
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    // positive normal numbers only (one integer range test instead of a DSETP): negative, zero, subnormal, Inf, NaN -> NaN
    const bool ok = (unsigned)(hx - 0x00100000) < (unsigned)(0x7ff00000 - 0x00100000);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    const double m = __hiloint2double(hx | (i ^ 0x3ff00000), lx);
    k += i >> 20;
    const double f = m - 1.0;
    double rr;
    {   // s = f/(2+f) by a refined reciprocal
        const double d = 2.0 + f;
        rr = rmt_rcp(d);
    }
    const double s = f*rr;
    const double z = s*s, w = z*z;
    const double t1 = w*fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z*fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01),
                                   2.857142874366239149e-01), 6.666666666666735130e-01);
    const double R = t2 + t1;
    const double hfsq = 0.5*f*f;
    const double dk = (double)k;
    const double res = dk*6.93147180369123816490e-01 - ((hfsq - fma(s, hfsq + R, dk*1.90821492927058770002e-10)) - f);
    return ok ? res : __longlong_as_double(0x7ff8000000000000LL);
#endif
}
__device__ __forceinline__ double rmt_log10(const double x)
{
#if RMT_EXACT_MATH
    return log10(x);
#else
    return rmt_log(x)*4.342944819032518167e-01;
#endif
}
__device__ __forceinline__ double rmt_sqrt(const double x)
{
#if RMT_EXACT_MATH
    return sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x*y, hh = 0.5*y;
    double r = fma(-hh, g, 0.5);
    g = fma(g, r, g); hh = fma(hh, r, hh);
    r = fma(-hh, g, 0.5);
    g = fma(g, r, g); hh = fma(hh, r, hh);
    const double d = fma(-g, g, x);
    g = fma(d, hh, g);
    // 0 -> 0; +Inf -> +Inf and NaN -> NaN (x + x) instead of the NaN / garbage the iteration makes of them
    const bool special = (__double2hiint(x) & 0x7fffffff) >= 0x7ff00000;
    return x == 0.0 ? 0.0 : (special ? x + x : g);
#endif
}
// x^a for the step-size controllers (x > 0): exp(a*log(x)) with the branch-free pair above.  libdevice's pow is a
// ~180-instruction function with branches, called 2.2 times per step attempt — 6.5 % of the integrator's
// instructions (ncu, profiles/r01_ncu_n1_solve_v4_extents.csv); the controller needs a few digits only.
__device__ __forceinline__ double rmt_powc(const double x, const double a)
{
#if RMT_EXACT_MATH
    return pow(x, a);
#else
    return rmt_exp(a*rmt_log(fmax(x, 1e-300)));
#endif
}
#define RMT_EXP(x) rmt_exp(x)
#define RMT_EXP10(x) rmt_exp10(x)
#define RMT_LOG(x) rmt_log(x)
#define RMT_LOG10(x) rmt_log10(x)
#define RMT_SQRT(x) rmt_sqrt(x)

// Partial derivatives of the traced rates.  With IEEE special values (RMT_EXACT_MATH) a saturated factor — exp() overflowed
// to Inf in a denominator, so the rate is exactly 0 like in NumPy — makes the SYMBOLIC derivative Inf/Inf or 0*Inf = NaN
// although the function is flat there; the reference's integrators difference the RHS numerically and see slope 0.  The
// exact-math build therefore takes a NaN partial as 0 (a NaN rate itself still rejects the step and fails loudly).
#if RMT_EXACT_MATH
#define RMT_DERIV(x) rmt_nan_to_zero(x)
__device__ __forceinline__ double rmt_nan_to_zero(const double v) { return v == v ? v : 0.0; }
#else
#define RMT_DERIV(x) (x)
#endif

// full-mantissa literals of the generated kinetics: constant-bank operands (1) or instruction immediates (0)
#ifndef RMT_USE_CBANK
#define RMT_USE_CBANK 1
#endif

#include "rmt_model.cuh"

// err^(1/order) for the step-size controller.  For the fourth-order tableaux this is a fourth root of a number in
// [1e-10, 1e30], and the controller needs three digits of it: two MUFU.RSQ in single precision (rsqrt(rsqrt(x)) =
// x^(1/4), relative error ~1e-6) instead of a double-precision log + exp — about 60 FP64 instructions less per call,
// 2.2 calls per step attempt.
__device__ __forceinline__ double rmt_root_order(const double x, const double inv_order)
{
#if !RMT_EXACT_MATH && RMT_ROS_ORDER == 4
    (void)inv_order;
    return (double)rsqrtf(rsqrtf((float)fmax(x, 1e-37)));
#else
    return rmt_powc(x, inv_order);
#endif
}

// max(x, 1e-30) (the reference's concentration clamp, pbHomoReactor.py:3897-3904) and its indicator on the integer pipe:
// fmax of doubles is a DSETP + NaN test + selects (~10 instructions, 12 clamps per node evaluation — 7 % of the
// stage-pipelined N2 kernel's instructions); comparing the HIGH WORD alone is one ISETP + two selects.  Values whose high
// word equals that of 1e-30 (|x/1e-30 - 1| < 2^-20) count as "above" and pass unchanged, negative numbers and zero give
// 1e-30, a NaN stays a NaN (the step is then rejected by the error test).
#define RMT_EPS_HI 0x39B4484B                       // high word of 1e-30 = 0x39B4484BFEEBC2A0
#ifndef RMT_FP_CLAMP
#define RMT_FP_CLAMP 0
#endif
__device__ __forceinline__ bool rmt_above_eps(const double x)
{
#if RMT_EXACT_MATH || RMT_FP_CLAMP
    return x > 1e-30;
#else
    return __double2hiint(x) >= RMT_EPS_HI;
#endif
}
__device__ __forceinline__ double rmt_clamp_eps(const double x)
{
#if RMT_EXACT_MATH || RMT_FP_CLAMP
    return fmax(x, 1e-30);
#else
    return rmt_above_eps(x) ? x : 1e-30;
#endif
}

#define RMT_R_CONST 8.314472            // core/constants.py:8
#define RMT_TREF 298.15                 // core/constants.py:17-23
#define RMT_PI 3.141592653589793        // core/constants.py:14
#define RMT_EPS_CONST 1e-30             // core/constants.py:11

// steady-state axial models share the integrator: N1 (dimensionless, [Ci..., P, (T)]) and its dimensional
// twin M7 = pbReactor.runM3 ([Ci..., T, P], docs/pbReactor.py:1170-1575)
#if defined(RMT_MODEL_N1) || defined(RMT_MODEL_M7)
#define RMT_STEADY 1
#endif
#if defined(RMT_MODEL_M7) || defined(RMT_MODEL_M9)
#define RMT_DIMENSIONAL 1                        // pbReactor.py twins: no scaling, viscosity and exchange area are inputs
#endif
#if defined(RMT_MODEL_N2) || defined(RMT_MODEL_M9)
#define RMT_DYNAMIC 1                            // method-of-lines models served by the rmt_n2_* kernels
#endif
#if defined(RMT_MODEL_M7)
#define RMT_N (RMT_NC + 2)                       // Ci [mol/m^3]..., T [K], P [Pa]
#define RMT_IT RMT_NC
#define RMT_IP (RMT_NC + 1)
#elif defined(RMT_MODEL_N1)
#define RMT_N (RMT_NC + (RMT_ISO ? 1 : 2))      // Ci..., P, (T)
#define RMT_IP RMT_NC                            // N1: index of P-hat
#define RMT_IT (RMT_NC + 1)                      // N1: index of T-hat
#else
#define RMT_N (RMT_NC + (RMT_ISO ? 0 : 1))      // Ci..., (T) per node
#endif
#define RMT_NCONST_VALUE (29 + RMT_NC + RMT_NKP)

// Linear-system dimension of the steady-state integrator.  The species balances of N1 and M7 are
// f_C = c(y) * nu^T R(y): every stage increment of a Rosenbrock method therefore lies in the range of
// E = [[nu^T, 0], [0, I]] and K_i = E k_i with (I/(h*gamma) - G E) k_i = g(Y_i) + sum_j (c_ij/h) k_j, where
// g = (c R_1..c R_nr, f_P, f_T) and G = dg/dy.  Solving for the k_i (dimension nr + 2) instead of the K_i
// (dimension nc + 2) gives the same iterates — error norm, guards and outputs are evaluated on the expanded
// state — with a smaller Jacobian, LU and triangular solves.  Used when nr < nc.
#ifndef RMT_REDUCED
#define RMT_REDUCED 0
#endif
#if defined(RMT_STEADY) && RMT_REDUCED
#define RMT_M (RMT_NR + RMT_N - RMT_NC)
#elif defined(RMT_STEADY)
#define RMT_M RMT_N
#endif

#ifndef RMT_BLOCK
#define RMT_BLOCK 256
#endif
// N2 integrator: lanes per reactor (nodes evaluated in parallel), see rmt_n2_solve; 0 = the stage-pipelined
// mapping (one thread per reactor and Rosenbrock stage, see "stage pipeline" below)
#ifndef RMT_N2_G
#define RMT_N2_G 1
#endif
#if RMT_N2_G == 0
#define RMT_N2_WF 1
#else
#define RMT_N2_WF 0
#endif
// warps of a block are kept in (loose) lockstep so that they share instruction-cache lines:
// 0 = free running, 1 = one block barrier per step attempt, 2 = one per Rosenbrock stage
#ifndef RMT_SYNC
#define RMT_SYNC 1
#endif
// 1 = the Rosenbrock stage loop is a real loop (one copy of the RHS code), 0 = fully unrolled
#ifndef RMT_ROLL
#define RMT_ROLL 1
#endif
// with RMT_SYNC == 1: a block barrier every RMT_SYNC_EVERY-th step attempt
// read each stage vector once per stage (argument and c-sum together) at the cost of 2n live registers
#ifndef RMT_MERGE_KLOADS
#define RMT_MERGE_KLOADS 1
#endif
#ifndef RMT_SYNC_EVERY
#define RMT_SYNC_EVERY 1
#endif
// a warp fetches new reactors when at least RMT_REFILL_MIN of its lanes are idle, or one has idled for
// more than RMT_REFILL_WAIT step attempts
#ifndef RMT_REFILL_MIN
#define RMT_REFILL_MIN 1
#endif
#ifndef RMT_REFILL_WAIT
#define RMT_REFILL_WAIT 2
#endif
// Lockstep groups of the steady-state integrator (RMT_SYNC == 1): the warps of a block are kept in step in groups of
// four (128 threads, one named barrier per group) instead of all together.  Four warps on the same instructions are
// enough to share the fetched instruction lines; the groups drift apart on their own (they pick up reactors at
// different times), so that while one group is in the FP64-heavy kinetics another factorises or solves — the FP64 pipe
// and the shared-memory pipe are loaded more evenly, and a barrier waits for 4 warps instead of 12.  Measured on 2^20
// config-3 reactors (384 threads): one group 11.95 ms, 2 groups 11.34, 3 groups 11.20, 4 groups 11.33, 6 groups 11.67;
// a deliberate phase offset at kernel start changes nothing.  Results are bit-identical (a reactor's arithmetic does
// not depend on which lane, warp or group integrates it).
#ifndef RMT_SYNC_GROUPS
#define RMT_SYNC_GROUPS ((RMT_BLOCK % 128 == 0 && RMT_BLOCK > 128) ? RMT_BLOCK/128 : 1)
#endif
// alternative: block-uniform refill every RMT_REFILL_EVERY-th attempt (0/1 = off)
#ifndef RMT_REFILL_EVERY
#define RMT_REFILL_EVERY 3
#endif

typedef long long i64;

// self-description of the module, read back by rmt_module_load()
extern "C" __global__ void rmt_meta(int* out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
#if defined(RMT_MODEL_N1)
    out[0] = 1;
#elif defined(RMT_MODEL_M7)
    out[0] = 7;
#elif defined(RMT_MODEL_M9)
    out[0] = 9;
#else
    out[0] = 2;
#endif
    out[1] = RMT_N; out[2] = RMT_NC; out[3] = RMT_NR; out[4] = RMT_NIN; out[5] = RMT_NCONST_VALUE;
    out[6] = RMT_NKP; out[7] = RMT_ROS_S; out[8] = RMT_BLOCK; out[9] = RMT_ISO;
    out[10] = RMT_FLOPS_RHS_ALG; out[11] = RMT_FLOPS_RHS_WT; out[12] = RMT_FLOPS_JAC_ALG; out[13] = RMT_FLOPS_JAC_WT;
#if defined(RMT_STEADY)
    out[14] = RMT_M;
#else
    out[14] = 0;
#endif
#if defined(RMT_MODEL_N2) || defined(RMT_MODEL_M9)
    out[15] = RMT_N2_G;                     // 0: stage-pipelined mapping (block = 32 reactors x (stages + 1) roles)
#else
    out[15] = 1;
#endif
}

// ---------------------------------------------------------------------------------
// primary per-instance inputs: the modelInput numbers a sweep may vary
// ---------------------------------------------------------------------------------
enum {
    IN_T = 0, IN_P = 1, IN_C0 = 2,
    IN_Q = 2 + RMT_NC, IN_D, IN_L, IN_DP, IN_EPS, IN_U, IN_TM,
    IN_MUG, IN_AEX,            // feed.mixture-viscosity and external-heat.EfHeTrAr: read by M7 / M9 only
    IN_CADE, IN_CASP,          // reactor.CaDe, reactor.CaSpHeCa: M9's energy balance only (pbReactor.py:2619)
    IN_KP0
};
static_assert(IN_KP0 + RMT_NKP == RMT_NIN, "input row count");

struct RmtInputs {
    const double* rows;        // varying rows, [n_rows][B]
    i64 B;
    int map[RMT_NIN];          // row index into `rows`, or -1 -> uniform value u[q]
    double u[RMT_NIN];
};

__device__ __forceinline__ double rmt_in(const RmtInputs& p, const int q, const i64 i)
{
    const int m = p.map[q];
    return m >= 0 ? __ldg(p.rows + (i64)m*p.B + i) : p.u[q];
}

// ---------------------------------------------------------------------------------
// per-instance constants (what runN1/runN2 put into paramsSet)
// ---------------------------------------------------------------------------------
enum {
    K_CMAX = 0, K_TF, K_PF, K_C0, K_UI0, K_US0, K_RHO0, K_CPF, K_GM, K_GH, K_MU,
    K_EPS, K_DP, K_ZF, K_U, K_A, K_TM, K_VF, K_MWF,
    // derived once per reactor for the hot loops (so that picking up a reactor is loads only)
    H_ERGA, H_ERGC, H_UA, H_INVC0, H_INVRHO0, H_INVGM, H_INVGH, H_EPSCPF, H_X0, H_X1,
    K_IV0,
    K_KP0 = K_IV0 + RMT_NC,
    RMT_NCONST = K_KP0 + RMT_NKP
};
static_assert(RMT_NCONST == RMT_NCONST_VALUE, "constant row count");

__device__ __forceinline__ double rmt_cp(const int i, const double T, const double T2, const double T3)
{
    // Cp string "a0 + a1*T + a2*(T**2) + a3*(T**3)", left to right (rmtThermo.py:37)
    double v = RMT_cCP[i][0] + RMT_cCP[i][1]*T + RMT_cCP[i][2]*T2;
    if (RMT_CP[i][3] != 0.0) v = v + RMT_cCP[i][3]*T3;
    return v;
}

// setup: docs/pbHomoReactor.py:2744-2852 (runN1) / :3370-3507 (runN2)
extern "C" __global__ void __launch_bounds__(128) rmt_setup(const RmtInputs in, double* __restrict__ consts)
{
    const i64 i = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= in.B) return;
    const i64 B = in.B;
    const double T = rmt_in(in, IN_T, i), P = rmt_in(in, IN_P, i);
    const double Q0 = rmt_in(in, IN_Q, i), D = rmt_in(in, IN_D, i), L = rmt_in(in, IN_L, i);
    const double dp = rmt_in(in, IN_DP, i), eps = rmt_in(in, IN_EPS, i);
    const double U = rmt_in(in, IN_U, i), Tm = rmt_in(in, IN_TM, i);
    double C0i[RMT_NC], y0[RMT_NC];
    double C0 = 0.0, Cmax = 0.0;
#pragma unroll
    for (int k = 0; k < RMT_NC; ++k) {
        C0i[k] = rmt_in(in, IN_C0 + k, i);
        C0 += C0i[k];
        Cmax = k == 0 ? C0i[k] : fmax(Cmax, C0i[k]);
    }
#pragma unroll
    for (int k = 0; k < RMT_NC; ++k) y0[k] = C0i[k]/C0;
    const double A = RMT_PI*(D*D)/4;                         // :2751
    const double ui0 = Q0/(A*eps);                           // :3137 / :3391
    const double us0 = ui0*eps;
#if defined(RMT_MODEL_N1) || defined(RMT_MODEL_M7)
    const double vf = Q0/A;                                  // :2763
#else
    const double vf = us0;                                   // :3452
#endif
#if defined(RMT_DIMENSIONAL)
    const double a = rmt_in(in, IN_AEX, i);                  // M7 / M9 use EfHeTrAr as given (pbReactor.py:1508, :2589)
#else
    const double a = 4/D;                                    // :2778 (EfHeTrAr input ignored)
#endif
    // pure-gas viscosities at feed T (gasTransPor.py:137-154, dataGasViscosity.py:133)
    double mu[RMT_NC];
#pragma unroll
    for (int k = 0; k < RMT_NC; ++k) {
        if (RMT_VISC_EQ[k] == 1)
            mu[k] = RMT_VISC[k][0]*1e-6*pow(T, RMT_VISC[k][1])/(1 + RMT_VISC[k][2]*(1/T) + RMT_VISC[k][3]*(1.0/(T*T)));
        else
            mu[k] = RMT_VISC[k][0]*pow(T, RMT_VISC[k][1])/(1 + (RMT_VISC[k][2]/T));
    }
    // Wilke mixing rule (gasTransPor.py:229-274)
    double mumix = 0.0;
#pragma unroll
    for (int p = 0; p < RMT_NC; ++p) {
        double den = 0.0;
#pragma unroll
        for (int q = 0; q < RMT_NC; ++q) {
            double phi;
            if (p == q) phi = 1.0;
            else {
                const int lo = p < q ? p : q, hi = p < q ? q : p;
                const double A1 = 1 + sqrt(mu[lo]/mu[hi])*sqrt(sqrt(RMT_MW[hi]/RMT_MW[lo]));
                const double up = (A1*A1)/sqrt(8*(1 + (RMT_MW[lo]/RMT_MW[hi])));
                phi = p < q ? up : (mu[p]/mu[q])*(RMT_MW[q]/RMT_MW[p])*up;
            }
            den += y0[q]*phi;
        }
        mumix += (mu[p]*y0[p])/den;
    }
#if defined(RMT_DIMENSIONAL)
    mumix = rmt_in(in, IN_MUG, i);                           // feed.mixture-viscosity (pbReactor.py:1235, :2069)
#endif
    const double T2 = T*T, T3 = T*T*T;
    double Cpf = 0.0, MWf = 0.0;
#pragma unroll
    for (int k = 0; k < RMT_NC; ++k) {
        const double cpm = (RMT_CPREF[k] + rmt_cp(k, T, T2, T3))*0.50;     // rmtThermo.py:52-75
        Cpf += y0[k]*cpm;
        MWf += y0[k]*RMT_MW[k];
    }
    MWf = MWf*1e-3;                                           // rmtUtility.py:57-95, "kg/mol"
    const double rho0 = MWf*C0;                               // :2796
    const double Gm = (vf/L)*Cmax;                            // :2819-2821 (GaMaCoTe0 == "MAX")
    const double Gh = (rho0*vf*T*(Cpf/MWf)/L);                // :2823
    double* c = consts + i;
    c[(i64)K_CMAX*B] = Cmax; c[(i64)K_TF*B] = T;   c[(i64)K_PF*B] = P;     c[(i64)K_C0*B] = C0;
    c[(i64)K_UI0*B] = ui0;   c[(i64)K_US0*B] = us0; c[(i64)K_RHO0*B] = rho0; c[(i64)K_CPF*B] = Cpf;
    c[(i64)K_GM*B] = Gm;     c[(i64)K_GH*B] = Gh;   c[(i64)K_MU*B] = mumix;  c[(i64)K_EPS*B] = eps;
    c[(i64)K_DP*B] = dp;     c[(i64)K_ZF*B] = L;    c[(i64)K_U*B] = U;       c[(i64)K_A*B] = a;
    c[(i64)K_TM*B] = Tm;     c[(i64)K_VF*B] = vf;   c[(i64)K_MWF*B] = MWf;
    // Ergun coefficients (pbHomoReactor.py:3214-3217): ergA*ergB = H_ERGA*us, ergC*ergD = H_ERGC*rho*us^2
    c[(i64)H_ERGA*B] = 150*mumix/(dp*dp)*(((1 - eps)*(1 - eps))/(eps*eps*eps));
    c[(i64)H_ERGC*B] = 1.75/dp*((1 - eps)/(eps*eps*eps));
    c[(i64)H_UA*B] = U*a;
    c[(i64)H_INVC0*B] = 1.0/C0;   c[(i64)H_INVRHO0*B] = 1.0/rho0;
    c[(i64)H_INVGM*B] = 1.0/Gm;   c[(i64)H_INVGH*B] = 1.0/Gh;
    c[(i64)H_EPSCPF*B] = eps/Cpf;
#if defined(RMT_STEADY)
    c[(i64)H_X0*B] = L/P;                  // 1/(Pf/zf), the Ergun scale (N1)
    c[(i64)H_X1*B] = 0.0;
#elif defined(RMT_MODEL_M9)
    c[(i64)H_X0*B] = 1/eps;                // const_F1, pbReactor.py:2615
    c[(i64)H_X1*B] = (1 - eps)*rmt_in(in, IN_CADE, i)*rmt_in(in, IN_CASP, i);      // catalyst term of const_T2, :2619
    c[(i64)H_EPSCPF*B] = eps;
#else
    c[(i64)H_X0*B] = 1/(eps*(L/vf));       // const_F1, :4075
    c[(i64)H_X1*B] = 1/(L/vf);
#endif
#pragma unroll
    for (int k = 0; k < RMT_NC; ++k)
#if defined(RMT_DIMENSIONAL)
        c[(i64)(K_IV0 + k)*B] = C0i[k];            // dimensional initial state (pbReactor.py:1240-1246, :2090-2103)
#else
        c[(i64)(K_IV0 + k)*B] = C0i[k]/Cmax;       // :2833 / :3489
#endif
#pragma unroll
    for (int k = 0; k < RMT_NKP; ++k) c[(i64)(K_KP0 + k)*B] = rmt_in(in, IN_KP0 + k, i);
}

// hot constants kept in registers by the RHS / integrator
struct Hot {
    double Cmax, Tf, Pf, us0, ergA, ergC, Ua, Tm;
    double invC0, invRho0, invGm, invGh, epsCpf;     // 1/C0, 1/rho0, 1/Gm, 1/Gh, eps/Cpf
#if defined(RMT_STEADY)
    double invBeta;              // zf/Pf
#else
    double F1, invZv;            // N2: 1/(eps*(zf/vf)), vf/zf;  M9: 1/eps, (1-eps)*CaDe*CaSpHeCa (epsCpf holds eps)
    double iv[RMT_NC];           // inlet boundary values C0_i/Cmax (M9: C0_i)
#if defined(RMT_MODEL_M9)
    double zf;                   // reactor length [m]: the grid is dimensional
#endif
#endif
    double kp[RMT_NKP > 0 ? RMT_NKP : 1];
};

__device__ __forceinline__ void rmt_load_hot(const double* __restrict__ consts, const i64 B, const i64 i, Hot& h)
{
    // loads only (independent, issued back to back): a lane picks up a new reactor while the other
    // lanes of its warp — and, through the block barrier, the other warps — wait for it
    const double* c = consts + i;
    h.Cmax = __ldg(c + (i64)K_CMAX*B); h.Tf = __ldg(c + (i64)K_TF*B); h.Pf = __ldg(c + (i64)K_PF*B);
    h.us0 = __ldg(c + (i64)K_US0*B); h.Tm = __ldg(c + (i64)K_TM*B);
    h.ergA = __ldg(c + (i64)H_ERGA*B); h.ergC = __ldg(c + (i64)H_ERGC*B); h.Ua = __ldg(c + (i64)H_UA*B);
    h.invC0 = __ldg(c + (i64)H_INVC0*B); h.invRho0 = __ldg(c + (i64)H_INVRHO0*B);
    h.invGm = __ldg(c + (i64)H_INVGM*B); h.invGh = __ldg(c + (i64)H_INVGH*B);
    h.epsCpf = __ldg(c + (i64)H_EPSCPF*B);
#if defined(RMT_STEADY)
    h.invBeta = __ldg(c + (i64)H_X0*B);
#else
    h.F1 = __ldg(c + (i64)H_X0*B);
    h.invZv = __ldg(c + (i64)H_X1*B);
#pragma unroll
    for (int k = 0; k < RMT_NC; ++k) h.iv[k] = __ldg(c + (i64)(K_IV0 + k)*B);
#if defined(RMT_MODEL_M9)
    h.zf = __ldg(c + (i64)K_ZF*B);
#endif
#endif
#pragma unroll
    for (int k = 0; k < RMT_NKP; ++k) h.kp[k] = __ldg(c + (i64)(K_KP0 + k)*B);
}

// ---------------------------------------------------------------------------------
// point physics shared by N1 and N2: everything of the RHS that depends on the
// local (C_i, T, P) only.  JAC adds the partial derivatives.
// ---------------------------------------------------------------------------------
struct Point {
    double y[RMT_NC];      // mole fractions
    double S, invS;        // total concentration [mol/m^3] and its reciprocal
    double T, P;
    double MWm, rho;       // mixture MW [kg/mol], EOS density
    double R[RMT_NR];      // reaction rates
    double r[RMT_NC];      // formation rates
    double cpm[RMT_NC];    // mean heat capacities
    double Cp;             // mixture mean heat capacity
    double dH[RMT_NR];     // heats of reaction at T
    double q, Qm;          // overall heat of reaction, coolant duty
};

struct PointJac {
    double dRdT[RMT_NR], dRdP[RMT_NR], dRdy[RMT_NR][RMT_NC], dRdC[RMT_NR][RMT_NC];
    double dCpdT;          // sum_i y_i * 0.5 * Cp_i'(T)
    double ddHdT[RMT_NR];  // d(dH_j)/dT
};

template <bool JAC>
__device__ __forceinline__ void rmt_point(const double (&C)[RMT_NC], const double T, const double P,
                                          const Hot& h, Point& p, PointJac& pj)
{
    double S = 0.0;
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) S += C[i];
    p.S = S; p.T = T; p.P = P;
    const double invS = rmt_rcp(S);
    p.invS = invS;
    double mw = 0.0;
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) { p.y[i] = C[i]*invS; mw += p.y[i]*RMT_cMW[i]; }
    p.MWm = mw*1e-3;
    p.rho = (P*p.MWm)*rmt_rcp(RMT_R_CONST*T);                 // P/((R/MW)*T), rmtThermo.py:353-369
    if (JAC) rmt_rates_jac(T, P, p.y, C, h.kp, p.R, pj.dRdT, pj.dRdP, pj.dRdy, pj.dRdC);
    else rmt_rates(T, P, p.y, C, h.kp, p.R);
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) {                        // rmtReaction.py:64-97
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < RMT_NR; ++j) if (RMT_NU[j][i] != 0.0) acc += RMT_NU[j][i]*p.R[j];   // small integers: immediates
        p.r[i] = acc;
    }
    const double T2 = T*T, T3 = T2*T;
    double Cp = 0.0, dCp = 0.0;
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) {
        p.cpm[i] = (RMT_cCPREF[i] + rmt_cp(i, T, T2, T3))*0.50;
        Cp += p.y[i]*p.cpm[i];
        if (JAC) {
            double d = RMT_cCP[i][1] + 2.0*RMT_cCP[i][2]*T;
            if (RMT_CP[i][3] != 0.0) d += 3.0*RMT_cCP[i][3]*T2;
            dCp += p.y[i]*(0.5*d);
        }
    }
    p.Cp = Cp;
    if (JAC) pj.dCpdT = dCp;
    const double dT = T - RMT_TREF;
    double q = 0.0;
#pragma unroll
    for (int j = 0; j < RMT_NR; ++j) {                        // rmtThermo.py:258-312 + StHeRe25
        const double dcp = RMT_cDCP[j][0] + RMT_cDCP[j][1]*T + RMT_cDCP[j][2]*T2 + RMT_cDCP[j][3]*T3;
        p.dH[j] = dcp*dT + RMT_cDH25[j];
        q += p.R[j]*p.dH[j];
        if (JAC) pj.ddHdT[j] = dcp + dT*(RMT_cDCP[j][1] + 2.0*RMT_cDCP[j][2]*T + 3.0*RMT_cDCP[j][3]*T2);
    }
    p.q = q;
    p.Qm = (h.Tm == 0.0) ? 0.0 : h.Ua*(h.Tm - T);             // rmtUtility.py:424-452
}

#if defined(RMT_STEADY)
struct NoJac { __device__ __forceinline__ void operator()(int, int, double) const {} };

#if defined(RMT_MODEL_M7)
// ---------------------------------------------------------------------------------
// M7 right-hand side (modelEquationM3, docs/pbReactor.py:1371-1575): the dimensional twin of N1.
// y = [C_i [mol/m^3]..., T [K], P [Pa]] along z [m].  Differences from N1 besides the scaling: the Ergun
// term uses rho = MW*C (not the EOS density), the mixture viscosity and the exchange area are inputs, and
// there is no adiabatic switch on MeTe.
// ---------------------------------------------------------------------------------
template <bool JAC, class JS>
__device__ __forceinline__ void n1_eval(const double (&yv)[RMT_N], const Hot& h, double (&f)[RMT_N], JS&& J)
{
    double C[RMT_NC];
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) C[i] = yv[i];
    const double T = yv[RMT_IT], P = yv[RMT_IP];
    Point p; PointJac pj;
    rmt_point<JAC>(C, T, P, h, p, pj);
    const double invP = rmt_rcp(P);
    const double w = (p.S*h.invC0)*(h.Pf*invP);               // rmtUtility.calGaVeFromEOS (:405-421)
    const double us = h.us0*w;                                // SuGaVe
    const double gade = p.MWm*p.S;                            // calDensityIG(MiMoWe, CoSp)
    f[RMT_IP] = -1*(h.ergA*us + h.ergC*gade*(us*us));
    const double c1 = rmt_rcp(us);
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) f[i] = p.r[i]*c1;
    const double Qm = h.Ua*(h.Tm - T);
    const double invDn = rmt_rcp((p.S*us)*p.Cp);              // 1/(MoFl*CpMeanMixture)
    f[RMT_IT] = (-p.q + Qm)*invDn;
    if (JAC) {
        const double invS = p.invS, invCp = rmt_rcp(p.Cp);
        double sy[RMT_NR];
#pragma unroll
        for (int j = 0; j < RMT_NR; ++j) {
            double a = 0.0;
#if RMT_RATES_DEP_Y
#pragma unroll
            for (int i = 0; i < RMT_NC; ++i) a += pj.dRdy[j][i]*p.y[i];
#endif
            sy[j] = a;
        }
#pragma unroll
        for (int col = 0; col < RMT_N; ++col) {
            const bool isC = col < RMT_NC, isP = col == RMT_IP;
            const int cc = col < RMT_NC ? col : 0;
            double dlnw, dgade = 0.0, dlnS = 0.0, dR[RMT_NR];
            if (isC) {
                dlnw = invS; dlnS = invS; dgade = 1e-3*RMT_cMW[cc];
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) {
                    double a = 0.0;
#if RMT_RATES_DEP_Y
                    a = (pj.dRdy[j][cc] - sy[j])*invS;
#endif
#if RMT_RATES_DEP_C
                    a += pj.dRdC[j][cc];
#endif
                    dR[j] = a;
                }
            } else if (isP) {
                dlnw = -invP;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dR[j] = pj.dRdP[j];
            } else {
                dlnw = 0.0;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dR[j] = pj.dRdT[j];
            }
            const double dus = us*dlnw;
            J(RMT_IP, col, -1*(h.ergA*dus + h.ergC*(dgade*(us*us) + 2.0*gade*us*dus)));
#pragma unroll
            for (int i = 0; i < RMT_NC; ++i) {
                double dr = 0.0;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) if (RMT_NU[j][i] != 0.0) dr += RMT_NU[j][i]*dR[j];
                J(i, col, dr*c1 - f[i]*dlnw);
            }
            double dq = 0.0;
#pragma unroll
            for (int j = 0; j < RMT_NR; ++j) dq += dR[j]*p.dH[j];
            double dCp = 0.0, dQm = 0.0;
            if (isC) dCp = (p.cpm[cc] - p.Cp)*invS;
            else if (!isP) {
                dCp = pj.dCpdT;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dq += p.R[j]*pj.ddHdT[j];
                dQm = -h.Ua;
            }
            J(RMT_IT, col, (-dq + dQm)*invDn - f[RMT_IT]*(dlnS + dlnw + dCp*invCp));
        }
    }
}
#else
// ---------------------------------------------------------------------------------
// N1 right-hand side (modelEquationN1, pbHomoReactor.py:3017-3314) and its analytic
// Jacobian.  `JS` receives J(i, j, value) = d f_i / d yhat_j.
// ---------------------------------------------------------------------------------

template <bool JAC, class JS>
__device__ __forceinline__ void n1_eval(const double (&yh)[RMT_N], const Hot& h, double (&f)[RMT_N], JS&& J)
{
    double C[RMT_NC];
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) C[i] = yh[i]*h.Cmax;    // :3159-3162
    const double P = yh[RMT_IP]*h.Pf;                         // :3173
#if RMT_ISO
    const double T = 0.0*h.Tf + h.Tf;                         // :3155, :3170
#else
    const double T = yh[RMT_IT]*h.Tf + h.Tf;
#endif
    Point p; PointJac pj;
    rmt_point<JAC>(C, T, P, h, p, pj);
    // velocities (rmtUtility.py:405-421; :3180-3186): u/u0 = (C/C0)*(Pf/P) for both the interstitial
    // and the superficial velocity (us = ui*eps, us0 = ui0*eps)
    const double invP = rmt_rcp(P);
    const double w = (p.S*h.invC0)*(h.Pf*invP);
    const double us = h.us0*w;
    const double rhoh = p.rho*h.invRho0;
    // Ergun (:3214-3220)
    f[RMT_IP] = -1*(h.ergA*us + h.ergC*p.rho*(us*us))*h.invBeta;
    const double c1 = rmt_rcp(w);
    const double c1Gm = c1*h.invGm;
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) f[i] = p.r[i]*c1Gm;         // :3283-3289
#if !RMT_ISO
    const double cpeffh = p.Cp*h.epsCpf;                       // (Cp/Cpf)*eps, :3252-3257
    const double invDn = rmt_rcp(rhoh*cpeffh*w);
    f[RMT_IT] = ((-p.q + p.Qm)*h.invGh)*invDn;                 // :3284, :3298
#endif
    if (JAC) {
        const double invS = p.invS;
        const double invT = rmt_rcp(T), invMW = rmt_rcp(p.MWm);
#if !RMT_ISO
        const double invCp = rmt_rcp(p.Cp);
#endif
        // sum_i dRdy[j][i]*y_i, used by every species column
        double sy[RMT_NR];
#pragma unroll
        for (int j = 0; j < RMT_NR; ++j) {
            double a = 0.0;
#if RMT_RATES_DEP_Y
#pragma unroll
            for (int i = 0; i < RMT_NC; ++i) a += pj.dRdy[j][i]*p.y[i];
#endif
            sy[j] = a;
        }
        const double inv_ushGm = c1Gm;
#pragma unroll
        for (int col = 0; col < RMT_N; ++col) {
            const bool isC = col < RMT_NC, isP = col == RMT_IP, isT = (!RMT_ISO) && col == RMT_IT;
            // d ln w, d ln rho, dR_j for this column (already multiplied by the column scale)
            double dlnw, dlnrho, dR[RMT_NR];
            if (isC) {
                dlnw = h.Cmax*invS;
                dlnrho = h.Cmax*(1e-3*RMT_cMW[col < RMT_NC ? col : 0] - p.MWm)*invS*invMW;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) {
                    double a = 0.0;
#if RMT_RATES_DEP_Y
                    a = (pj.dRdy[j][col < RMT_NC ? col : 0] - sy[j])*invS;
#endif
#if RMT_RATES_DEP_C
                    a += pj.dRdC[j][col < RMT_NC ? col : 0];
#endif
                    dR[j] = h.Cmax*a;
                }
            } else if (isP) {
                dlnw = -h.Pf*invP;
                dlnrho = h.Pf*invP;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dR[j] = h.Pf*pj.dRdP[j];
            } else {
                dlnw = 0.0;
                dlnrho = -h.Tf*invT;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dR[j] = h.Tf*pj.dRdT[j];
            }
            const double dus = us*dlnw;
            J(RMT_IP, col, -1*(h.ergA*dus + h.ergC*(p.rho*dlnrho*(us*us) + 2.0*p.rho*us*dus))*h.invBeta);
#pragma unroll
            for (int i = 0; i < RMT_NC; ++i) {
                double dr = 0.0;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) if (RMT_NU[j][i] != 0.0) dr += RMT_NU[j][i]*dR[j];
                J(i, col, dr*inv_ushGm - f[i]*dlnw);
            }
#if !RMT_ISO
            double dq = 0.0;
#pragma unroll
            for (int j = 0; j < RMT_NR; ++j) dq += dR[j]*p.dH[j];
            double dCp, dQm = 0.0;
            if (isC) dCp = h.Cmax*(p.cpm[col < RMT_NC ? col : 0] - p.Cp)*invS;
            else if (isT) {
                dCp = h.Tf*pj.dCpdT;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dq += p.R[j]*(h.Tf*pj.ddHdT[j]);
                dQm = (h.Tm == 0.0) ? 0.0 : -h.Ua*h.Tf;
            } else dCp = 0.0;
            const double dN = (-dq + dQm)*h.invGh;
            const double dlnD = dlnrho + dCp*invCp + dlnw;
            J(RMT_IT, col, dN*invDn - f[RMT_IT]*dlnD);
#endif
        }
    }
}

#endif  // N1 vs M7 right-hand side

// ---------------------------------------------------------------------------------
// System form seen by the integrator: g (dimension RMT_M) and A = dg/dx for its unknowns x.
// Full form: x = y, g = f, A = J.  Reduced form (RMT_REDUCED): x = (xi_1..xi_nr, non-species unknowns) with
// y_C = y_C(0) + nu^T xi, g = (c R_j, f_P, f_T), A = G E — derivatives along the reaction directions.
// ---------------------------------------------------------------------------------
#define RMT_NX (RMT_N - RMT_NC)               // non-species unknowns (P, T)

// d = E x: increments of the integrator's unknowns -> increments of the full state
__device__ __forceinline__ void rmt_expand(const double (&x)[RMT_M], double (&d)[RMT_N])
{
#if RMT_REDUCED
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < RMT_NR; ++j) if (RMT_NU[j][i] != 0.0) acc += RMT_NU[j][i]*x[j];
        d[i] = acc;
    }
#pragma unroll
    for (int q = 0; q < RMT_NX; ++q) d[RMT_NC + q] = x[RMT_NR + q];
#else
#pragma unroll
    for (int i = 0; i < RMT_N; ++i) d[i] = x[i];
#endif
}

#if !RMT_REDUCED
template <bool JAC, class JS>
__device__ __forceinline__ void n1_eval_sys(const double (&y)[RMT_N], const Hot& h, double (&g)[RMT_M], JS&& A)
{
    n1_eval<JAC>(y, h, g, A);
}
#else
#define RMT_RP (RMT_NR + RMT_IP - RMT_NC)     // position of P / T among the reduced unknowns
#define RMT_RT (RMT_NR + RMT_IT - RMT_NC)

// compile-time sums over the stoichiometry of reaction k: net mole change and net molar mass [kg/mol]
__device__ constexpr double rmt_nu_sum(const int k)
{
    double a = 0.0;
    for (int i = 0; i < RMT_NC; ++i) a += RMT_NU[k][i];
    return a;
}
__device__ constexpr double rmt_nu_mw(const int k)
{
    double a = 0.0;
    for (int i = 0; i < RMT_NC; ++i) a += RMT_NU[k][i]*(1e-3*RMT_MW[i]);
    return a;
}

template <bool JAC, class JS>
__device__ __forceinline__ void n1_eval_sys(const double (&yv)[RMT_N], const Hot& h, double (&g)[RMT_M], JS&& A)
{
    double C[RMT_NC];
#if defined(RMT_MODEL_M7)
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) C[i] = yv[i];
    const double T = yv[RMT_IT], P = yv[RMT_IP];
    const double sC = 1.0, sP = 1.0, sT = 1.0;               // column scales d(C,P,T)/d(unknown)
#else
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) C[i] = yv[i]*h.Cmax;
    const double P = yv[RMT_IP]*h.Pf;
#if RMT_ISO
    const double T = 0.0*h.Tf + h.Tf;
#else
    const double T = yv[RMT_IT]*h.Tf + h.Tf;
#endif
    const double sC = h.Cmax, sP = h.Pf, sT = h.Tf;
#endif
    Point p; PointJac pj;
    rmt_point<JAC>(C, T, P, h, p, pj);
    const double invP = rmt_rcp(P);
    const double w = (p.S*h.invC0)*(h.Pf*invP);
    const double us = h.us0*w;
#if defined(RMT_MODEL_M7)
    const double gade = p.MWm*p.S;
    g[RMT_RP] = -1*(h.ergA*us + h.ergC*gade*(us*us));
    const double cR = rmt_rcp(us);                            // species balances: f_i = r_i/us
    const double Qm = h.Ua*(h.Tm - T);
    const double invDn = rmt_rcp((p.S*us)*p.Cp);
    g[RMT_RT] = (-p.q + Qm)*invDn;
#else
    const double rhoh = p.rho*h.invRho0;
    g[RMT_RP] = -1*(h.ergA*us + h.ergC*p.rho*(us*us))*h.invBeta;
    const double cR = rmt_rcp(w)*h.invGm;                     // species balances: f_i = r_i/(w*Gm)
#if !RMT_ISO
    const double cpeffh = p.Cp*h.epsCpf;
    const double invDn = rmt_rcp(rhoh*cpeffh*w);
    g[RMT_RT] = ((-p.q + p.Qm)*h.invGh)*invDn;
#endif
#endif
#pragma unroll
    for (int j = 0; j < RMT_NR; ++j) g[j] = p.R[j]*cR;
    if (JAC) {
        const double invS = p.invS;
#if !defined(RMT_MODEL_M7)
        const double invT = rmt_rcp(T), invMW = rmt_rcp(p.MWm);
#endif
#if !RMT_ISO
        const double invCp = rmt_rcp(p.Cp);
#endif
        double sy[RMT_NR];
#pragma unroll
        for (int j = 0; j < RMT_NR; ++j) {
            double a = 0.0;
#if RMT_RATES_DEP_Y
#pragma unroll
            for (int i = 0; i < RMT_NC; ++i) a += pj.dRdy[j][i]*p.y[i];
#endif
            sy[j] = a;
        }
#pragma unroll
        for (int d = 0; d < RMT_M; ++d) {
            const bool isX = d < RMT_NR, isP = d == RMT_RP;
            const int k = d < RMT_NR ? d : 0;
            // derivatives along direction d: d ln w, d ln rho (N1) or d(MW*C) and d ln S (M7), dR_j, dCp
            double dlnw, dlnrho = 0.0, dgade = 0.0, dlnS = 0.0, dCp = 0.0, dR[RMT_NR];
            if (isX) {
                const double sn = rmt_nu_sum(k);
                dlnw = (sC*sn)*invS;
                dlnS = dlnw;
                dgade = sC*rmt_nu_mw(k);
#if !defined(RMT_MODEL_M7)
                dlnrho = sC*(rmt_nu_mw(k) - p.MWm*sn)*invS*invMW;
#endif
                double scp = 0.0;
#pragma unroll
                for (int i = 0; i < RMT_NC; ++i) if (RMT_NU[k][i] != 0.0) scp += RMT_NU[k][i]*p.cpm[i];
                dCp = sC*(scp - p.Cp*sn)*invS;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) {
                    double a = 0.0;
#if RMT_RATES_DEP_Y
                    double ty = 0.0;
#pragma unroll
                    for (int i = 0; i < RMT_NC; ++i) if (RMT_NU[k][i] != 0.0) ty += RMT_NU[k][i]*pj.dRdy[j][i];
                    a = (ty - sy[j]*sn)*invS;
#endif
#if RMT_RATES_DEP_C
#pragma unroll
                    for (int i = 0; i < RMT_NC; ++i) if (RMT_NU[k][i] != 0.0) a += RMT_NU[k][i]*pj.dRdC[j][i];
#endif
                    dR[j] = sC*a;
                }
            } else if (isP) {
                dlnw = -sP*invP;
#if !defined(RMT_MODEL_M7)
                dlnrho = sP*invP;
#endif
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dR[j] = sP*pj.dRdP[j];
            } else {
                dlnw = 0.0;
#if !defined(RMT_MODEL_M7)
                dlnrho = -sT*invT;
#endif
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dR[j] = sT*pj.dRdT[j];
            }
            const double dus = us*dlnw;
#if defined(RMT_MODEL_M7)
            A(RMT_RP, d, -1*(h.ergA*dus + h.ergC*(dgade*(us*us) + 2.0*gade*us*dus)));
#else
            A(RMT_RP, d, -1*(h.ergA*dus + h.ergC*(p.rho*dlnrho*(us*us) + 2.0*p.rho*us*dus))*h.invBeta);
#endif
#pragma unroll
            for (int j = 0; j < RMT_NR; ++j) A(j, d, dR[j]*cR - g[j]*dlnw);
#if !RMT_ISO
            double dq = 0.0, dQm = 0.0;
#pragma unroll
            for (int j = 0; j < RMT_NR; ++j) dq += dR[j]*p.dH[j];
            if (!isX && !isP) {
                dCp = sT*pj.dCpdT;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dq += p.R[j]*(sT*pj.ddHdT[j]);
#if defined(RMT_MODEL_M7)
                dQm = -h.Ua;
#else
                dQm = (h.Tm == 0.0) ? 0.0 : -h.Ua*sT;
#endif
            }
#if defined(RMT_MODEL_M7)
            A(RMT_RT, d, (-dq + dQm)*invDn - g[RMT_RT]*(dlnS + dlnw + dCp*invCp));
#else
            A(RMT_RT, d, ((-dq + dQm)*h.invGh)*invDn - g[RMT_RT]*(dlnrho + dCp*invCp + dlnw));
#endif
#endif
            (void)dgade; (void)dlnS; (void)dlnrho;
        }
    }
}
#endif  // system form

// stand-alone batched RHS: y [N][B] -> f [N][B]   (parity + "RHS evals/s" kernel)
extern "C" __global__ void __launch_bounds__(128)
rmt_n1_rhs(const double* __restrict__ consts, const i64 B, const double* __restrict__ y, double* __restrict__ f)
{
    const i64 i = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= B) return;
    Hot h; rmt_load_hot(consts, B, i, h);
    double yh[RMT_N], fo[RMT_N];
#pragma unroll
    for (int k = 0; k < RMT_N; ++k) yh[k] = y[(i64)k*B + i];
    n1_eval<false>(yh, h, fo, NoJac());
#pragma unroll
    for (int k = 0; k < RMT_N; ++k) f[(i64)k*B + i] = fo[k];
}

struct GlobalJac {
    double* J; i64 B, i;
    __device__ __forceinline__ void operator()(int r, int c, double v) const { J[(i64)(r*RMT_N + c)*B + i] = v; }
};

// stand-alone batched Jacobian: y [N][B] -> f [N][B], J [N*N][B] (row-major d f_r / d y_c)
// (HBM-bound, n^2 stores per reactor: occupancy buys memory-level parallelism, RMT_JAC_MINBLOCKS blocks per SM)
#ifndef RMT_JAC_MINBLOCKS
#define RMT_JAC_MINBLOCKS 2
#endif
extern "C" __global__ void __launch_bounds__(128, RMT_JAC_MINBLOCKS)
rmt_n1_jac(const double* __restrict__ consts, const i64 B, const double* __restrict__ y,
           double* __restrict__ f, double* __restrict__ J)
{
    const i64 i = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= B) return;
    Hot h; rmt_load_hot(consts, B, i, h);
    double yh[RMT_N], fo[RMT_N];
#pragma unroll
    for (int k = 0; k < RMT_N; ++k) yh[k] = y[(i64)k*B + i];
    GlobalJac gj{J, B, i};
    n1_eval<true>(yh, h, fo, gj);
#pragma unroll
    for (int k = 0; k < RMT_N; ++k) f[(i64)k*B + i] = fo[k];
}

struct GlobalSys {
    double* A; i64 B, i;
    __device__ __forceinline__ void operator()(int r, int c, double v) const { A[(i64)(r*RMT_M + c)*B + i] = v; }
};

// the integrator's system form at given states: y [N][B] -> g [M][B], A [M*M][B] (parity of the
// reduced-coordinate Jacobian: E A == J E and E g == f)
extern "C" __global__ void __launch_bounds__(128)
rmt_n1_sys(const double* __restrict__ consts, const i64 B, const double* __restrict__ y,
           double* __restrict__ g, double* __restrict__ A)
{
    const i64 i = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= B) return;
    Hot h; rmt_load_hot(consts, B, i, h);
    double yh[RMT_N], go[RMT_M];
#pragma unroll
    for (int k = 0; k < RMT_N; ++k) yh[k] = y[(i64)k*B + i];
    GlobalSys gs{A, B, i};
    n1_eval_sys<true>(yh, h, go, gs);
#pragma unroll
    for (int k = 0; k < RMT_M; ++k) g[(i64)k*B + i] = go[k];
}

// ---------------------------------------------------------------------------------
// N1 integrator: one reactor instance per lane, adaptive Rosenbrock (tableau from the
// generated header), SciPy-style error test  rms(err / (atol + rtol*max(|y|,|y_new|))) <= 1.
// The n x n iteration matrix W = I/(h*gamma) - J, its LU factors and the stage vectors
// K_1..K_s live in shared memory, column-interleaved across the block ([slot][tid]) so
// every access is conflict-free.  Lanes pull instances from a global queue, so a lane
// that finishes early immediately starts another reactor instead of idling until the
// slowest reactor of its warp is done.
// ---------------------------------------------------------------------------------
struct SolveArgs {
    const double* consts;      // [NCONST][B]
    i64 B;
    const double* z_eval;      // [n_eval] increasing, last = end of domain
    int n_eval;
    int out_mode;              // 0: raw scaled state, 1: dataYs rows (y_i, P[Pa], T[K]), 2: raw | C_i | dataYs
    double rtol, atol;
    int max_steps;
    int dense;                 // 1: interpolate eval points; 0: step onto each of them
    double* out;               // [n_eval][N][B]
    int* status;               // [B]
    int* stats;                // [4][B]: accepted, rejected, nfev, njev
    unsigned long long* queue; // work counter (zeroed by the host before launch)
    // optional fused objective (parameter estimation): sum_k ((out_k - ref_k)/ref_k)^2 over the outlet
    const double* obj_ref;     // [N] or null
    double* obj;               // [B]
    // step-size controller: safety, max shrink factor, max growth factor, tolerance scale kappa
    // (error test uses kappa*(atol + rtol*|y|)), PI exponent beta (0 = Gustafsson only), initial-step factor
    double ctrl[6];
    // optional step log of one instance: rows of (t, h, err, accepted) — diagnostics only
    double* trace;             // [trace_cap][4] or null
    i64 trace_inst;
    int trace_cap;
    // optional fused reduction of the objective (parameter-estimation populations): the last block to finish folds
    // obj[0..B) and status[0..B) into red[0..3] = sum, min, argmin + red_offset (as a double), number of failed
    // reactors — in a fixed order, so the result is deterministic — and no second kernel is needed before the
    // cross-GPU step.  The block-completion counter sits behind the work counter (queue + 1).
    double* red;               // [4] or null
    i64 red_offset;
};

#define SM(slot) sm[(slot)*RMT_BLOCK]
#define LU(r, c) SM((r)*RMT_M + (c))
#define KS(s, i) SM(RMT_M*RMT_M + (s)*RMT_M + (i))
#define RMT_SMEM_DOUBLES_PER_THREAD (RMT_M*RMT_M + RMT_ROS_S*RMT_M)

struct SmemJac {            // stores -J: the iteration matrix is W = I/(h*gamma) - J
    double* sm;
    __device__ __forceinline__ void operator()(int r, int c, double v) const { LU(r, c) = -v; }
};

// rows written per output point: mode 0 raw scaled state (sol.y); mode 1 dataYs rows
// (y_i, P [Pa], T [K]); mode 2 everything runN1 packs: raw | C_i [mol/m^3] | dataYs rows
__device__ __forceinline__ int n1_out_rows(const int mode) { return mode == 2 ? 2*RMT_N + RMT_NC : RMT_N; }

__device__ __forceinline__ void n1_write_point(const SolveArgs& a, const Hot& h, const i64 inst, const int e,
                                                const double (&v)[RMT_N])
{
    const int rows = n1_out_rows(a.out_mode);
    double* o = a.out + ((i64)e*rows)*a.B + inst;
    if (a.out_mode != 1) {
#pragma unroll
        for (int k = 0; k < RMT_N; ++k) o[(i64)k*a.B] = v[k];
        o += (i64)RMT_N*a.B;
    }
    if (a.out_mode != 0) {
        // sortResult4 (solResultAnalysis.py:191-249) + mole fractions (pbHomoReactor.py:2973-2983);
        // M7: mole fractions from the dimensional concentrations (pbReactor.py:1301-1312)
        double S = 0.0, C[RMT_NC];
#if defined(RMT_MODEL_M7)
#pragma unroll
        for (int k = 0; k < RMT_NC; ++k) { C[k] = v[k]; S += C[k]; }
#else
#pragma unroll
        for (int k = 0; k < RMT_NC; ++k) { C[k] = v[k]*h.Cmax; S += C[k]; }
#endif
        if (a.out_mode == 2) {
#pragma unroll
            for (int k = 0; k < RMT_NC; ++k) o[(i64)k*a.B] = C[k];
            o += (i64)RMT_NC*a.B;
        }
#pragma unroll
        for (int k = 0; k < RMT_NC; ++k) o[(i64)k*a.B] = C[k]/S;
#if defined(RMT_MODEL_M7)
        o[(i64)RMT_IT*a.B] = v[RMT_IT];
        o[(i64)RMT_IP*a.B] = v[RMT_IP];
#else
        o[(i64)RMT_IP*a.B] = v[RMT_IP]*h.Pf;
#if !RMT_ISO
        o[(i64)RMT_IT*a.B] = v[RMT_IT]*h.Tf + h.Tf;
#endif
#endif
    }
}

// register cap of the integrator: from the block size (one block per SM), or explicit (RMT_MAXNREG: ptxas rounds a
// launch-bounds cap down to an occupancy step — 448 threads gives 128 registers although 144 fit)
#ifdef RMT_MAXNREG
#define RMT_SOLVE_BOUNDS __maxnreg__(RMT_MAXNREG)
#else
#define RMT_SOLVE_BOUNDS __launch_bounds__(RMT_BLOCK)
#endif
extern "C" __global__ void RMT_SOLVE_BOUNDS rmt_n1_solve(const SolveArgs a)
{
    extern __shared__ double smem[];
    double* sm = smem + threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;

    i64 inst = -1;
    bool exhausted = false;
    Hot h = {};
    double y[RMT_N];
    double t = 0.0, hstep = 0.0, hacc = 0.0, erracc = 0.0;
    int nacc = 0, nrej = 0, next_e = 0, nanrej = 0;
    // Output grid: everything a lane needs when it picks up a reactor is read here, once per thread, so that the
    // refill itself has no dependent loads: the end of the domain, the number of leading output points at z <= 0
    // (written straight from the initial state) and the first output position inside the domain.
    const double tend = a.z_eval[a.n_eval - 1];
    int first_e = 0;
    while (first_e < a.n_eval && a.z_eval[first_e] <= 0.0) ++first_e;
    const double z_first = first_e < a.n_eval ? a.z_eval[first_e] : tend;
    double znext = z_first;                 // z_eval[next_e] (tend when all points are written)
    bool last_rejected = false, fresh = false;
#if RMT_SYNC && RMT_SYNC_EVERY > 1
    int iter = 0;
#endif
#if RMT_REFILL_MIN > 1
    int idle = 0;
#endif
#if RMT_REFILL_EVERY > 1
    int rslot = 0;
#endif

    while (true) {
        // ---- refill idle lanes from the queue (warp-aggregated atomic) ----
        const bool need = (inst < 0) && !exhausted;
        const unsigned m = __ballot_sync(FULL, need);
#if RMT_REFILL_EVERY > 1
        // block-uniform refill slots: every warp of the block picks up reactors in the same attempt, so the
        // load latency is paid once per slot by everybody at the same time
        const bool go = (rslot++ % RMT_REFILL_EVERY) == 0;
        if (m && go) {
#elif RMT_REFILL_MIN > 1
        // picking up a reactor makes the rest of the warp (and, through the barrier, of the block) wait for
        // the loads: do it for several lanes at once — a lane idles at most RMT_REFILL_WAIT attempts
        if (need) ++idle;
        const bool go = __popc(m) >= RMT_REFILL_MIN || __any_sync(FULL, need && idle > RMT_REFILL_WAIT);
        if (m && go) {
#else
        if (m) {
#endif
            unsigned long long base = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(m));
            base = __shfl_sync(FULL, base, leader);
            if (need) {
                const i64 cand = (i64)base + __popc(m & ((1u << lane) - 1));
#if RMT_REFILL_MIN > 1
                idle = 0;
#endif
                if (cand >= a.B) exhausted = true;
                else {
                    inst = cand;
                    rmt_load_hot(a.consts, a.B, inst, h);
#pragma unroll
                    for (int k = 0; k < RMT_NC; ++k) y[k] = a.consts[(i64)(K_IV0 + k)*a.B + inst];
#if defined(RMT_MODEL_M7)
                    y[RMT_IT] = h.Tf; y[RMT_IP] = h.Pf;       // pbReactor.py:1244-1246
#else
                    y[RMT_IP] = 1.0;                          // P/Pf, :2834
#if !RMT_ISO
                    y[RMT_IT] = 0.0;                          // (T-Tf)/Tf, :2838
#endif
#endif
                    t = 0.0; nacc = 0; nrej = 0; nanrej = 0; next_e = 0; last_rejected = false; fresh = true;
                    hacc = 0.0; erracc = 1e-2;
                    for (; next_e < first_e; ++next_e) n1_write_point(a, h, inst, next_e, y);
                    znext = z_first;
                }
            }
        }
#if RMT_SYNC
#if RMT_SYNC_EVERY > 1
        if ((++iter % RMT_SYNC_EVERY) == 0)
#endif
#if RMT_SYNC_GROUPS > 1
        {   // lockstep groups: named barrier 1 + group index, and-reduction of "this lane has no reactor" over the group
            unsigned allidle;
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %3, 0;\n\tbar.red.and.pred p, %1, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(allidle) : "r"(1 + (int)(threadIdx.x/(RMT_BLOCK/RMT_SYNC_GROUPS))), "r"(RMT_BLOCK/RMT_SYNC_GROUPS), "r"((unsigned)(inst < 0)) : "memory");
            if (allidle) break;
        }
#else
        if (__syncthreads_and(inst < 0)) break;      // block-uniform exit; also re-aligns the warps
#endif
#else
        if (__all_sync(FULL, inst < 0)) break;
#endif
        // Lanes without work: in the free-running / per-attempt modes they skip the attempt; with per-stage
        // barriers (RMT_SYNC == 2) they run it as ghosts on their stale state — all global writes below are
        // predicated on `live` — so that every thread reaches every barrier converged.
        const bool live = inst >= 0;
#if RMT_SYNC != 2
        if (!live) continue;
#endif

        // ---- one step attempt ----
        double f0[RMT_M];
        SmemJac sj{sm};
        n1_eval_sys<true>(y, h, f0, sj);                      // g(y_n), and -A into LU(.,.)

        if (fresh) {
            // initial step (Hairer-Wanner II.4 with the exact y'' = J f)
            double d0 = 0.0, d1 = 0.0, d2 = 0.0;
            double ag[RMT_M], fy[RMT_N], jfy[RMT_N];
#pragma unroll
            for (int i = 0; i < RMT_M; ++i) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < RMT_M; ++j) v -= LU(i, j)*f0[j];
                ag[i] = v;
            }
            rmt_expand(f0, fy);                               // y' = E g
            rmt_expand(ag, jfy);                              // y'' = E A g
#pragma unroll
            for (int i = 0; i < RMT_N; ++i) {
                const double isc = rmt_rcp(a.ctrl[3]*(a.atol + a.rtol*fabs(y[i])));
                d0 += (y[i]*isc)*(y[i]*isc); d1 += (fy[i]*isc)*(fy[i]*isc); d2 += (jfy[i]*isc)*(jfy[i]*isc);
            }
            d0 = rmt_sqrt(d0*(1.0/RMT_N)); d1 = rmt_sqrt(d1*(1.0/RMT_N)); d2 = rmt_sqrt(d2*(1.0/RMT_N));
            const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01*d0*rmt_rcp(d1);
            const double dm = fmax(d1, d2);
            const double h1 = dm <= 1e-15 ? fmax(1e-6, h0*1e-3) : rmt_powc(0.01*rmt_rcp(dm), 1.0/(RMT_ROS_ORDER + 1));
            hstep = fmin(a.ctrl[5]*fmin(100.0*h0, h1), tend);
            fresh = false;
        }
        // clip to the end of the domain / next output point
        double hlim = tend - t;
        const bool dense = RMT_ROS_DENSE && a.dense;
        if (!dense) hlim = znext - t;
        const bool clipped = hstep*1.01 >= hlim;
        const double hh = clipped ? hlim : hstep;

        // W = I/(h*gamma) - J, LU with partial pivoting.  The row permutation lives in registers as the
        // shared-memory address of each (pivoted) row, so every access is [row + immediate].
        const double invh = rmt_rcp(hh);                      // (refined reciprocal: IEEE division costs 2-3x the instructions)
        const double dg = invh*(1.0/RMT_ROS_GAMMA);
#pragma unroll
        for (int i = 0; i < RMT_M; ++i) LU(i, i) += dg;
        double* row[RMT_M];
        int perm[RMT_M];
#pragma unroll
        for (int i = 0; i < RMT_M; ++i) { row[i] = sm + (i*RMT_M)*RMT_BLOCK; perm[i] = i; }
#define ROW(i, j) row[i][(j)*RMT_BLOCK]
#pragma unroll
        for (int k = 0; k < RMT_M; ++k) {
            double best = fabs(ROW(k, k));
            int bi = k;
#pragma unroll
            for (int i = k + 1; i < RMT_M; ++i) {
                const double v = fabs(ROW(i, k));
                if (v > best) { best = v; bi = i; }
            }
#pragma unroll
            for (int i = k + 1; i < RMT_M; ++i)
                if (i == bi) {
                    double* tr = row[k]; row[k] = row[i]; row[i] = tr;
                    const int tp = perm[k]; perm[k] = perm[i]; perm[i] = tp;
                }
            const double piv = rmt_rcp(ROW(k, k));
            ROW(k, k) = piv;                                  // store reciprocal pivot
            double urow[RMT_M];
#pragma unroll
            for (int j = k + 1; j < RMT_M; ++j) urow[j] = ROW(k, j);
#pragma unroll
            for (int i = k + 1; i < RMT_M; ++i) {
                const double l = ROW(i, k)*piv;
                ROW(i, k) = l;
#pragma unroll
                for (int j = k + 1; j < RMT_M; ++j) ROW(i, j) -= l*urow[j];
            }
        }

        // stages
#if RMT_ROS_REUSE
        double flast[RMT_M];                   // latest evaluated f (a re-using stage takes the previous stage's)
#pragma unroll
        for (int i = 0; i < RMT_M; ++i) flast[i] = f0[i];
#else
        const double (&flast)[RMT_M] = f0;
#endif

#if RMT_ROLL
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int s = 0; s < RMT_ROS_S; ++s) {
#if RMT_SYNC == 2
            __syncthreads();
#endif
            double rhs[RMT_M];
#if RMT_ROLL
            if (s == 0 || !RMT_cROS_NEWF[s]) {
#else
            if (s == 0 || !RMT_ROS_NEWF[s]) {
#endif
                // first stage, or a stage with the same argument as the previous one: re-use its function value
                // (flast == f(y_n) until a later stage has evaluated a new one)
#pragma unroll
                for (int i = 0; i < RMT_M; ++i) rhs[i] = flast[i];
#if RMT_ROLL
                for (int j = 0; j < s; ++j) {
                    const double cj = RMT_cROS_C[s][j]*invh;
                    const double* kj = &KS(j, 0);
#pragma unroll
                    for (int i = 0; i < RMT_M; ++i) rhs[i] += cj*kj[i*RMT_BLOCK];
                }
#else
#pragma unroll
                for (int j = 0; j < s; ++j)
                    if (RMT_ROS_C[s][j] != 0.0) {
                        const double cj = RMT_cROS_C[s][j]*invh;
#pragma unroll
                        for (int i = 0; i < RMT_M; ++i) rhs[i] += cj*KS(j, i);
                    }
#endif
            } else {
                double ur[RMT_M];
#pragma unroll
                for (int i = 0; i < RMT_M; ++i) ur[i] = 0.0;
#if RMT_ROLL
                // rolled form: one copy of the RHS / triangular-solve code for all stages (instruction-cache
                // footprint); tableau rows are read from the constant bank with a runtime stage index.
                // Each K_j is read once for both the stage argument and the (c_sj/h) K_j sum.
#if RMT_MERGE_KLOADS
                double vc[RMT_M];
#pragma unroll
                for (int i = 0; i < RMT_M; ++i) vc[i] = 0.0;
#endif
                for (int j = 0; j < s; ++j) {
                    const double aj = RMT_cROS_A[s][j];
                    const double* kj = &KS(j, 0);
#if RMT_MERGE_KLOADS
                    const double cj = RMT_cROS_C[s][j]*invh;
#pragma unroll
                    for (int i = 0; i < RMT_M; ++i) { const double kv = kj[i*RMT_BLOCK]; ur[i] += aj*kv; vc[i] += cj*kv; }
#else
#pragma unroll
                    for (int i = 0; i < RMT_M; ++i) ur[i] += aj*kj[i*RMT_BLOCK];
#endif
                }
#else
#pragma unroll
                for (int j = 0; j < s; ++j)
                    if (RMT_ROS_A[s][j] != 0.0) {
#pragma unroll
                        for (int i = 0; i < RMT_M; ++i) ur[i] += RMT_cROS_A[s][j]*KS(j, i);
                    }
#endif
                double u[RMT_N];
                rmt_expand(ur, u);
#pragma unroll
                for (int i = 0; i < RMT_N; ++i) u[i] += y[i];
                n1_eval_sys<false>(u, h, rhs, NoJac());
#if RMT_ROS_REUSE
#pragma unroll
                for (int i = 0; i < RMT_M; ++i) flast[i] = rhs[i];
#endif
#if RMT_ROLL
#if RMT_MERGE_KLOADS
#pragma unroll
                for (int i = 0; i < RMT_M; ++i) rhs[i] += vc[i];
#else
                for (int j = 0; j < s; ++j) {
                    const double cj = RMT_cROS_C[s][j]*invh;
                    const double* kj = &KS(j, 0);
#pragma unroll
                    for (int i = 0; i < RMT_M; ++i) rhs[i] += cj*kj[i*RMT_BLOCK];
                }
#endif
#else
#pragma unroll
                for (int j = 0; j < s; ++j)
                    if (RMT_ROS_C[s][j] != 0.0) {
                        const double cj = RMT_cROS_C[s][j]*invh;
#pragma unroll
                        for (int i = 0; i < RMT_M; ++i) rhs[i] += cj*KS(j, i);
                    }
#endif
            }
            // solve W k = rhs: the right-hand side is stashed in the K_s slot so that the permuted read
            // is an address, not a register index
            double* ks = &KS(s, 0);
#pragma unroll
            for (int i = 0; i < RMT_M; ++i) ks[i*RMT_BLOCK] = rhs[i];
            double x[RMT_M];
#pragma unroll
            for (int i = 0; i < RMT_M; ++i) {
                double v = ks[perm[i]*RMT_BLOCK];
#pragma unroll
                for (int j = 0; j < i; ++j) v -= ROW(i, j)*x[j];
                x[i] = v;
            }
#pragma unroll
            for (int i = RMT_M - 1; i >= 0; --i) {
                double v = x[i];
#pragma unroll
                for (int j = i + 1; j < RMT_M; ++j) v -= ROW(i, j)*x[j];
                x[i] = v*ROW(i, i);
            }
#pragma unroll
            for (int i = 0; i < RMT_M; ++i) ks[i*RMT_BLOCK] = x[i];
        }
#undef ROW
        // sum m_s k_s and sum e_s k_s in the integrator's unknowns, from the stage vectors in shared memory (accumulating
        // them inside the stage loop would hold 2m doubles in registers across every right-hand-side evaluation)
        double dxs[RMT_M], evs[RMT_M];
#pragma unroll
        for (int i = 0; i < RMT_M; ++i) {
            double dx = 0.0, ev = 0.0;
#pragma unroll
            for (int s = 0; s < RMT_ROS_S; ++s) {
                const double kv = KS(s, i);
                if (RMT_ROS_M[s] == 1.0) dx += kv;
                else if (RMT_ROS_M[s] != 0.0) dx += RMT_cROS_M[s]*kv;
                if (RMT_ROS_E[s] == 1.0) ev += kv;
                else if (RMT_ROS_E[s] != 0.0) ev += RMT_cROS_E[s]*kv;
            }
            dxs[i] = dx; evs[i] = ev;
        }
        double ynew[RMT_N], errv[RMT_N];
        rmt_expand(dxs, ynew);
        rmt_expand(evs, errv);
#pragma unroll
        for (int i = 0; i < RMT_N; ++i) ynew[i] += y[i];

        // error norm (scipy/integrate/_ivp/common.py:63-65 rms norm; radau.py scale)
        double err = 0.0;
        bool bad = false;
#pragma unroll
        for (int i = 0; i < RMT_N; ++i) {
            const double sc = a.ctrl[3]*(a.atol + a.rtol*fmax(fabs(y[i]), fabs(ynew[i])));
            const double e = errv[i]*rmt_rcp(sc);
            err += e*e;
            bad = bad || !(fabs(ynew[i]) <= 1.7e308);
        }
        err = rmt_sqrt(err*(1.0/RMT_N));
        // domain guard.  RMT_POSITIVE[i]: the kinetics have a pole where species i vanishes (it reaches a denominator
        // that can become zero, a logarithm, a negative / non-integer power — decided by a sign analysis of the traced
        // rates and their partials, kinetics.positive_species): such a species must stay strictly positive, a step
        // that violates this is rejected like a failed error test (the reference raises there).  Every other species
        // may sit at exactly zero (zero feed and never formed, e.g. inside an LHHW term 1 + K*p_i) or reach it
        // (irreversible reaction at complete conversion, where the exact zero sits at the rounding floor with a random
        // sign): a value that is negative within the error tolerance IS that zero and is taken as such; only a value
        // further below zero than the tolerance allows is rejected (a negative concentration can run away through
        // second-order terms, e.g. -k*C^2).
#pragma unroll
        for (int i = 0; i < RMT_NC; ++i) {
            if (RMT_POSITIVE[i]) bad = bad || !(ynew[i] > 0.0);
            else {
                bad = bad || (ynew[i] < -(a.ctrl[3]*(a.atol + a.rtol*fabs(y[i]))));
                ynew[i] = fmax(ynew[i], 0.0);
            }
        }
        if (bad || !(err == err)) err = 1e30;

        // step-size controller: Hairer-Wanner with Gustafsson's predictive correction
        if (a.trace && inst == a.trace_inst && nacc + nrej < a.trace_cap) {
            double* tr = a.trace + 4*(nacc + nrej);
            tr[0] = t; tr[1] = hh; tr[2] = err; tr[3] = err <= 1.0 ? 1.0 : 0.0;
        }
        const double ISAFE = rmt_rcp(a.ctrl[0]), FAC1 = a.ctrl[1], FAC2 = rmt_rcp(a.ctrl[2]), BETA = a.ctrl[4];
        const double errc = fmax(err, 1e-10);
        double fac;
        if (BETA > 0.0 && nacc > 0)            // PI controller (Gustafsson 1991): uses the previous accepted error
            fac = rmt_powc(errc, 1.0/(RMT_ROS_ORDER) - 0.75*BETA)*rmt_powc(erracc, -BETA)*ISAFE;   // note erracc^(-beta): small previous error -> grow
        else
            fac = rmt_root_order(errc, 1.0/(RMT_ROS_ORDER))*ISAFE;
        fac = fmax(FAC2, fmin(FAC1, fac));
        double hnew = hh*rmt_rcp(fac);
        int fin = -1;
        if (err <= 1.0) {
            if (nacc > 0 && BETA <= 0.0) {
                double facgus = (hacc*invh)*rmt_root_order(err*err*rmt_rcp(erracc), 1.0/(RMT_ROS_ORDER))*ISAFE;
                facgus = fmax(FAC2, fmin(FAC1, facgus));
                fac = fmax(fac, facgus);
                hnew = hh*rmt_rcp(fac);
            }
            hacc = hh; erracc = fmax(1e-2, err);
            ++nacc; nanrej = 0;
            const double tnew = clipped ? (dense ? tend : znext) : t + hh;
            // output points inside (t, tnew]
            if (dense) {
                while (next_e < a.n_eval && a.z_eval[next_e] <= tnew) {
                    const double ze = a.z_eval[next_e];
                    double v[RMT_N];
                    if (ze >= tnew) {
#pragma unroll
                        for (int i = 0; i < RMT_N; ++i) v[i] = ynew[i];
                    } else {
                        const double th = (ze - t)*invh, th1 = 1.0 - th;
                        double dr[RMT_M], dy[RMT_N];
#pragma unroll
                        for (int i = 0; i < RMT_M; ++i) {
                            double d2 = 0.0, d3 = 0.0;
#pragma unroll
                            for (int s = 0; s < RMT_ROS_S; ++s) {
                                if (RMT_ROS_D[0][s] != 0.0) d2 += RMT_cROS_D[0][s]*KS(s, i);
                                if (RMT_ROS_D[1][s] != 0.0) d3 += RMT_cROS_D[1][s]*KS(s, i);
                            }
                            dr[i] = d2 + th*d3;
                        }
                        rmt_expand(dr, dy);
#pragma unroll
                        for (int i = 0; i < RMT_N; ++i) v[i] = y[i]*th1 + th*(ynew[i] + th1*dy[i]);
                    }
                    if (live) n1_write_point(a, h, inst, next_e, v);
                    ++next_e;
                }
            } else if (clipped && next_e < a.n_eval) {
                if (live) n1_write_point(a, h, inst, next_e, ynew);
                ++next_e;
                znext = next_e < a.n_eval ? a.z_eval[next_e] : tend;
            }
            t = tnew;
#pragma unroll
            for (int i = 0; i < RMT_N; ++i) y[i] = ynew[i];
            if (last_rejected) hnew = fmin(hnew, hh);
            last_rejected = false;
            // a clipped step says nothing about the step the error would allow
            hstep = clipped ? fmax(hnew, hstep) : hnew;
            if (t >= tend) fin = 0;
            else if (nacc + nrej >= a.max_steps) fin = 1;
        } else {
            ++nrej;
            if (err >= 1e29) { ++nanrej; hnew = hh*0.1; }
            last_rejected = true;
            hstep = hnew;
            if (nacc + nrej >= a.max_steps) fin = 1;
            else if (hstep < 1e-14*fmax(tend, 1.0)) fin = 2;
            else if (nanrej > 30) fin = 3;
        }
        if (fin >= 0 && live) {
            a.status[inst] = fin;
            if (a.stats) {                                       // optional (rmt_b200.h)
                a.stats[inst] = nacc; a.stats[a.B + inst] = nrej;
                int nnew = 0;
#pragma unroll
                for (int q = 1; q < RMT_ROS_S; ++q) nnew += RMT_ROS_NEWF[q];
                a.stats[2*a.B + inst] = (nacc + nrej)*nnew; a.stats[3*a.B + inst] = nacc + nrej;
            }
            if (fin != 0) {
                // make failures loud in the data as well: NaN for every point not yet written
                double v[RMT_N];
#pragma unroll
                for (int i = 0; i < RMT_N; ++i) v[i] = __longlong_as_double(0x7ff8000000000000LL);
                const int rows = n1_out_rows(a.out_mode);
                for (; next_e < a.n_eval; ++next_e) {
                    double* o = a.out + ((i64)next_e*rows)*a.B + inst;
                    for (int k = 0; k < rows; ++k) o[(i64)k*a.B] = v[0];
                }
            }
            if (a.obj) {
                double ob = 0.0;
                if (fin == 0) {
                    double S = 0.0, C[RMT_NC];
#pragma unroll
#if defined(RMT_MODEL_M7)
                    for (int k = 0; k < RMT_NC; ++k) { C[k] = y[k]; S += C[k]; }
#else
                    for (int k = 0; k < RMT_NC; ++k) { C[k] = y[k]*h.Cmax; S += C[k]; }     // same arithmetic as the output rows
#endif
#pragma unroll
                    for (int k = 0; k < RMT_NC; ++k) { const double d = (C[k]/S - a.obj_ref[k])/a.obj_ref[k]; ob += d*d; }
#if defined(RMT_MODEL_M7)
                    { const double d = (y[RMT_IT] - a.obj_ref[RMT_IT])/a.obj_ref[RMT_IT]; ob += d*d; }
#elif !RMT_ISO
                    { const double d = ((y[RMT_IT]*h.Tf + h.Tf) - a.obj_ref[RMT_IT])/a.obj_ref[RMT_IT]; ob += d*d; }
#endif
                } else ob = __longlong_as_double(0x7ff0000000000000LL);
                a.obj[inst] = ob;
            }
            inst = -1;
        }
    }
    if (a.red) {
        // ---- fused objective reduction: last block out folds the shard ----
        __shared__ int is_last;
        __threadfence();                                   // this block's obj / status writes are visible device-wide ...
        __syncthreads();
        if (threadIdx.x == 0) {                            // ... before its arrival is counted
            const unsigned arrived = atomicAdd(reinterpret_cast<unsigned int*>(a.queue + 1), 1u);
            is_last = arrived == gridDim.x - 1;
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            const double inf = __longlong_as_double(0x7ff0000000000000LL);
            double sum = 0.0, mn = inf, nbad = 0.0;
            i64 am = -1;
            for (i64 i = threadIdx.x; i < a.B; i += RMT_BLOCK) {          // fixed assignment: deterministic
                const double x = __ldcg(a.obj + i);
                sum += x;
                if (x < mn) { mn = x; am = i + a.red_offset; }
                nbad += __ldcg(a.status + i) != 0 ? 1.0 : 0.0;
            }
            double* ssum = smem; double* smin = smem + RMT_BLOCK; double* sbad = smem + 2*RMT_BLOCK;
            i64* sarg = reinterpret_cast<i64*>(smem + 3*RMT_BLOCK);
            __syncthreads();                               // the integrator's shared-memory slots are free now
            ssum[threadIdx.x] = sum; smin[threadIdx.x] = mn; sbad[threadIdx.x] = nbad; sarg[threadIdx.x] = am;
            __syncthreads();
            if (threadIdx.x == 0) {                        // fixed order over the block's threads
                for (int k = 1; k < RMT_BLOCK; ++k) {
                    sum += ssum[k]; nbad += sbad[k];
                    const double o = smin[k]; const i64 oa = sarg[k];
                    if (o < mn || (o == mn && oa >= 0 && (am < 0 || oa < am))) { mn = o; am = oa; }
                }
                a.red[0] = sum; a.red[1] = mn; a.red[2] = (double)am; a.red[3] = nbad;
            }
        }
    }
}
#endif  // RMT_STEADY

#if defined(RMT_DYNAMIC)
// ---------------------------------------------------------------------------------
// N2: dynamic model by the method of lines (modelEquationN2, pbHomoReactor.py:3706-4134).
// State yhat[(nc+1)][zNo] variable-major as in the reference (:3873); on the device
// [var][node][B].  The nodes are swept in flow direction (1..32 lanes per reactor, see rmt_n2_solve) because
// the pressure is marched node by node (:3979) and the convection is first-order upwind
// (:4082-4128), so node k depends on nodes <= k only.
// ---------------------------------------------------------------------------------
#define RMT_ITN RMT_NC                       // N2: index of T-hat within a node

struct NodeJac {                             // derivative blocks of one node
    double A[RMT_N][RMT_N];                  // d f_k / d u_k
    double g[RMT_N];                         // d f_k / d P_k
    double L[RMT_N];                         // d f_k / d u_{k-1}  (diagonal: upwind differences)
    double e[RMT_N];                         // d E_k / d u_k
    double ep;                               // d E_k / d P_k
#if defined(RMT_MODEL_M9)
    // M9 marches the superficial velocity as well (v_{k+1} = v_k + dz*V_k, pbReactor.py:2606-2612), and V_k sees
    // the temperature of the node before (dT/dz), which puts a T_{k-1} column into every species balance
    double gv[RMT_N];                        // d f_k / d v_k
    double Lt[RMT_N];                        // d f_k / d T_{k-1} beyond the diagonal (species rows)
    double ev;                               // d E_k / d v_k
    double eV[RMT_N];                        // d V_k / d u_k
    double eVb, eVP, eVv;                    // d V_k / d T_{k-1}, d P_k, d v_k
#endif
};

// f_k and the Ergun gradient E_k [Pa/m] at one node.  u: node state, ub: upwind node state (or the
// inlet boundary values), P: pressure at the node.
// (the diagonal block d f_k / d u_k is handed entry by entry to `asink(row, col, value)`, column by column, so that it
// never has to exist in registers as a whole: the integrator stores W_kk = I/(h gamma) - A straight into shared memory)
struct NodeJacSink { NodeJac& nj; __device__ __forceinline__ void operator()(int r, int c, double v) const { nj.A[r][c] = v; } };
struct NoSink { __device__ __forceinline__ void operator()(int, int, double) const {} };

template <bool JAC, class AS>
__device__ __forceinline__ void n2_node(const double (&u)[RMT_N], const double (&ub)[RMT_N], const bool inlet,
                                        const double P, const double invdz, const Hot& h,
                                        double (&f)[RMT_N], double& E, NodeJac& nj, AS&& asink)
{
    double C[RMT_NC];
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) C[i] = rmt_clamp_eps(u[i])*h.Cmax;      // :3897-3904
#if RMT_ISO
    const double T = 0.0*h.Tf + h.Tf;
#else
    const double T = u[RMT_ITN]*h.Tf + h.Tf;                                       // :3914
#endif
    Point p; PointJac pj;
    rmt_point<JAC>(C, T, P, h, p, pj);
    const double us = h.us0;                                                       // v_z frozen, :3937, :4066
    E = -1*(h.ergA*us + h.ergC*p.rho*(us*us));                                     // :3970-3974 (not scaled)
    const double rhoh = p.rho*h.invRho0;
    // mass balances (:4082-4099): F1*(-v/vf*(Ci - Ci_b)/dz + ri/GaMaCoTe0), v/vf == 1
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) {
        const double cb = inlet ? h.iv[i] : rmt_clamp_eps(ub[i]);
        f[i] = h.F1*(-1*((u[i] - cb)*invdz) + p.r[i]*h.invGm);
    }
#if !RMT_ISO
    // energy balance (:4102-4128)
    const double cph_eps = p.Cp*h.epsCpf;                                          // (Cp/Cpf)*eps
    const double invD = rmt_rcp(rhoh*cph_eps);
    const double tb = inlet ? 0.0 : ub[RMT_ITN];                                   // (T0 - Tf)/Tf = 0 at the inlet
    const double dTdz = (u[RMT_ITN] - tb)*invdz;
    const double Nn = (-p.q + p.Qm)*h.invGh;
    f[RMT_ITN] = h.invZv*(-dTdz + Nn*invD);
#endif
    if (JAC) {
        const double invS = p.invS, invT = rmt_rcp(T), invMW = rmt_rcp(p.MWm), invP = rmt_rcp(P);
#if !RMT_ISO
        const double invCp = rmt_rcp(p.Cp);
#endif
        double sy[RMT_NR];
#pragma unroll
        for (int j = 0; j < RMT_NR; ++j) {
            double a = 0.0;
#if RMT_RATES_DEP_Y
#pragma unroll
            for (int i = 0; i < RMT_NC; ++i) a += pj.dRdy[j][i]*p.y[i];
#endif
            sy[j] = a;
        }
        const double F1Gm = h.F1*h.invGm;
        const double ergE = -1*h.ergC*(us*us)*p.rho;            // dE = ergE * dlnrho
        // columns: local species, local temperature, and the pressure (col == RMT_N)
#pragma unroll
        for (int col = 0; col <= RMT_N; ++col) {
            const bool isC = col < RMT_NC, isP = col == RMT_N;
            double dlnrho, dR[RMT_NR];
            if (isC) {
                const int c = col < RMT_NC ? col : 0;
                const double sc = rmt_above_eps(u[c]) ? h.Cmax : 0.0;    // d max(u, eps)/du
                dlnrho = sc*(1e-3*RMT_cMW[c] - p.MWm)*invS*invMW;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) {
                    double a = 0.0;
#if RMT_RATES_DEP_Y
                    a = (pj.dRdy[j][c] - sy[j])*invS;
#endif
#if RMT_RATES_DEP_C
                    a += pj.dRdC[j][c];
#endif
                    dR[j] = sc*a;
                }
            } else if (isP) {
                dlnrho = invP;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dR[j] = pj.dRdP[j];
            } else {
                dlnrho = -h.Tf*invT;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dR[j] = h.Tf*pj.dRdT[j];
            }
            if (isP) nj.ep = ergE*dlnrho; else nj.e[col < RMT_N ? col : 0] = ergE*dlnrho;
#pragma unroll
            for (int i = 0; i < RMT_NC; ++i) {
                double dr = 0.0;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) if (RMT_NU[j][i] != 0.0) dr += RMT_NU[j][i]*dR[j];
                double v = dr*F1Gm;
                if (!isP && i == col) v -= h.F1*invdz;
                if (isP) nj.g[i] = v; else asink(i, col < RMT_N ? col : 0, v);
            }
#if !RMT_ISO
            double dq = 0.0;
#pragma unroll
            for (int j = 0; j < RMT_NR; ++j) dq += dR[j]*p.dH[j];
            double dCp = 0.0, dQm = 0.0;
            if (isC) {
                const int c = col < RMT_NC ? col : 0;
                dCp = (rmt_above_eps(u[c]) ? h.Cmax : 0.0)*(p.cpm[c] - p.Cp)*invS;
            } else if (!isP) {
                dCp = h.Tf*pj.dCpdT;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dq += p.R[j]*(h.Tf*pj.ddHdT[j]);
                dQm = (h.Tm == 0.0) ? 0.0 : -h.Ua*h.Tf;
            }
            const double dN = (-dq + dQm)*h.invGh;
            const double dlnD = dlnrho + dCp*invCp;
            double vT = h.invZv*(dN*invD - Nn*invD*dlnD);
            if (!isP && col == RMT_ITN) vT -= h.invZv*invdz;
            if (isP) nj.g[RMT_ITN] = vT; else asink(RMT_ITN, col < RMT_N ? col : 0, vT);
#endif
        }
        // upwind coupling (diagonal)
#pragma unroll
        for (int i = 0; i < RMT_NC; ++i) nj.L[i] = inlet ? 0.0 : (rmt_above_eps(ub[i]) ? h.F1*invdz : 0.0);
#if !RMT_ISO
        nj.L[RMT_ITN] = inlet ? 0.0 : h.invZv*invdz;
#endif
    }
}

#if defined(RMT_MODEL_M9)
// ---------------------------------------------------------------------------------
// M9 node (modelEquationM5, docs/pbReactor.py:2296-2660): the dimensional twin of the N2 node.  u = (C_i
// [kmol/m^3 in the reference's own script], T [K]); pressure P and superficial velocity v at the node come from the
// marches P_{k+1} = P_k + dz*E_k (:2546) and v_{k+1} = v_k + dz*V_k (:2606-2612).  Differences from N2 besides the
// scaling: rho = MW*C (not the EOS density), the velocity is not frozen, species balances carry -C_i*dv/dz, the
// energy balance has the catalyst's heat capacity, the coolant duty is in kJ.
// ---------------------------------------------------------------------------------
template <bool JAC>
__device__ __forceinline__ void m9_node(const double (&u)[RMT_N], const double (&ub)[RMT_N], const bool inlet,
                                        const double P, const double v, const double invdz, const Hot& h,
                                        double (&f)[RMT_N], double& E, double& V, NodeJac& nj)
{
    double C[RMT_NC];
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) C[i] = rmt_clamp_eps(u[i]);            // :2487-2491
    const double T = u[RMT_ITN];
    Point p; PointJac pj;
    rmt_point<JAC>(C, T, P, h, p, pj);
    const double gade = p.MWm*p.S;                                                 // calDensityIG, :2528
    E = -1*(h.ergA*v + h.ergC*gade*(v*v));                                         // :2534-2544
    double OvR = 0.0;
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) OvR += p.r[i];                                // :2559
    const double Qm = p.Qm*1e-3;                                                   // 'kJ/m^3.s', rmtUtility.py:446-447
    const double tb = inlet ? h.Tf : ub[RMT_ITN];                                  // :2597-2600
    const double dTdz = (T - tb)*invdz;
    const double invT = rmt_rcp(T);
    const double PT2 = P*invT*invT;
    const double a1 = rmt_rcp(p.S*1000);
    const double vR = -v*(1.0/RMT_R_CONST);
    const double br = invT*E - PT2*dTdz;                                           // (1/T)*dP/dz - (P/T^2)*dT/dz
    V = a1*(vR*br + OvR*1000);                                                     // :2606-2608
    double cb[RMT_NC];
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) {
        cb[i] = inlet ? h.iv[i] : rmt_clamp_eps(ub[i]);
        f[i] = h.F1*(-v*((u[i] - cb[i])*invdz) - u[i]*V + p.r[i]);                 // :2626-2640 (centre value un-clamped)
    }
    const double SCp = p.S*p.Cp;
    const double Svc = (p.S*v)*p.Cp;                                               // const_T1 = MoFl*CpMeanMixture
    const double invD = rmt_rcp(SCp*h.epsCpf + h.invZv);                           // const_T2, :2619
    f[RMT_ITN] = invD*(-Svc*dTdz + (-p.q + Qm));                                   // :2651
    if (JAC) {
        const double invS = p.invS;
        double sy[RMT_NR];
#pragma unroll
        for (int j = 0; j < RMT_NR; ++j) {
            double a = 0.0;
#if RMT_RATES_DEP_Y
#pragma unroll
            for (int i = 0; i < RMT_NC; ++i) a += pj.dRdy[j][i]*p.y[i];
#endif
            sy[j] = a;
        }
        nj.ev = -1*(h.ergA + 2.0*h.ergC*gade*v);
        nj.ep = 0.0;
        // columns: local species, local temperature, pressure (col == RMT_N), velocity (col == RMT_N + 1)
#pragma unroll
        for (int col = 0; col <= RMT_N + 1; ++col) {
            const bool isC = col < RMT_NC, isT = col == RMT_ITN, isP = col == RMT_N, isV = col == RMT_N + 1;
            const int c = col < RMT_NC ? col : 0;
            double dS = 0.0, dE = 0.0, dSCp = 0.0, dR[RMT_NR];
#pragma unroll
            for (int j = 0; j < RMT_NR; ++j) dR[j] = 0.0;
            if (isC) {
                const double sc = rmt_above_eps(u[c]) ? 1.0 : 0.0;             // d max(u, eps)/du
                dS = sc;
                dE = -1*h.ergC*(v*v)*(sc*1e-3*RMT_cMW[c]);
                dSCp = sc*p.cpm[c];
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) {
                    double a = 0.0;
#if RMT_RATES_DEP_Y
                    a = (pj.dRdy[j][c] - sy[j])*invS;
#endif
#if RMT_RATES_DEP_C
                    a += pj.dRdC[j][c];
#endif
                    dR[j] = sc*a;
                }
            } else if (isT) {
                dSCp = p.S*pj.dCpdT;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dR[j] = pj.dRdT[j];
            } else if (isP) {
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dR[j] = pj.dRdP[j];
            } else {
                dE = nj.ev;
            }
            double dr[RMT_NC], dOvR = 0.0;
#pragma unroll
            for (int i = 0; i < RMT_NC; ++i) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) if (RMT_NU[j][i] != 0.0) a += RMT_NU[j][i]*dR[j];
                dr[i] = a; dOvR += a;
            }
            double dbr;                                                            // d(vR*br)
            if (isC) dbr = vR*(invT*dE);
            else if (isT) dbr = vR*(-(E*invT*invT) + 2.0*PT2*invT*dTdz - PT2*invdz);
            else if (isP) dbr = vR*(-(invT*invT)*dTdz);
            else dbr = -(1.0/RMT_R_CONST)*br + vR*(invT*dE);
            const double dV = a1*(dbr + 1000*dOvR) - V*dS*invS;
#pragma unroll
            for (int i = 0; i < RMT_NC; ++i) {
                double val = h.F1*(-u[i]*dV + dr[i]);
                if (isC && i == col) val -= h.F1*(v*invdz + V);
                if (isV) val -= h.F1*((u[i] - cb[i])*invdz);
                if (isP) nj.g[i] = val; else if (isV) nj.gv[i] = val; else nj.A[i][col < RMT_N ? col : 0] = val;
            }
            double dq = 0.0;
#pragma unroll
            for (int j = 0; j < RMT_NR; ++j) dq += dR[j]*p.dH[j];
            double dN = -v*dTdz*dSCp - dq;
            if (isT) {
#pragma unroll
                for (int j = 0; j < RMT_NR; ++j) dN -= p.R[j]*pj.ddHdT[j];
                dN += -Svc*invdz + ((h.Tm == 0.0) ? 0.0 : -h.Ua*1e-3);
            }
            if (isV) dN -= SCp*dTdz;
            const double valT = (dN - f[RMT_ITN]*(h.epsCpf*dSCp))*invD;
            if (isP) { nj.g[RMT_ITN] = valT; nj.eVP = dV; }
            else if (isV) { nj.gv[RMT_ITN] = valT; nj.eVv = dV; }
            else { nj.A[RMT_ITN][col < RMT_N ? col : 0] = valT; nj.e[col < RMT_N ? col : 0] = dE; nj.eV[col < RMT_N ? col : 0] = dV; }
        }
        // node before: upwind concentrations (diagonal) and its temperature (through dT/dz in V and in the energy balance)
        const double dVb = inlet ? 0.0 : a1*(vR*(PT2*invdz));
        nj.eVb = dVb;
#pragma unroll
        for (int i = 0; i < RMT_NC; ++i) {
            nj.L[i] = inlet ? 0.0 : (rmt_above_eps(ub[i]) ? h.F1*(v*invdz) : 0.0);
            nj.Lt[i] = h.F1*(-u[i]*dVb);
        }
        nj.L[RMT_ITN] = inlet ? 0.0 : (Svc*invdz)*invD;
        nj.Lt[RMT_ITN] = 0.0;
    }
}
#endif  // RMT_MODEL_M9

// stand-alone batched RHS: y [n][zNo][B] -> f [n][zNo][B]
extern "C" __global__ void __launch_bounds__(64)
rmt_n2_rhs(const double* __restrict__ consts, const i64 B, const int zNo, const double* __restrict__ y, double* __restrict__ f)
{
    const i64 i = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= B) return;
    Hot h; rmt_load_hot(consts, B, i, h);
#if defined(RMT_MODEL_M9)
    const double dz = h.zf/(zNo - 1);                                              // pbReactor.py:2076
#else
    const double dz = 1.0/(zNo - 1);                                               // :3439
#endif
    const double invdz = 1.0/dz;
    double P = h.Pf;                                                               // P_z[0] = P0, :3848
    double ub[RMT_N] = {0}, u[RMT_N], fo[RMT_N], E;
    NodeJac nj;
#if defined(RMT_MODEL_M9)
    double vs = h.us0, V;                                                          // v_z[0] = SuGaVe0, pbReactor.py:2433
#endif
    for (int k = 0; k < zNo; ++k) {
#pragma unroll
        for (int v = 0; v < RMT_N; ++v) u[v] = y[((i64)v*zNo + k)*B + i];
#if defined(RMT_MODEL_M9)
        m9_node<false>(u, ub, k == 0, P, vs, invdz, h, fo, E, V, nj);
        vs = V*dz + vs;                                                            // pbReactor.py:2612
#else
        n2_node<false>(u, ub, k == 0, P, invdz, h, fo, E, nj, NoSink());
#endif
#pragma unroll
        for (int v = 0; v < RMT_N; ++v) { f[((i64)v*zNo + k)*B + i] = fo[v]; ub[v] = u[v]; }
        P = E*dz + P;                                                              // :3979 (dimensionless dz, kept)
    }
}

// ---------------------------------------------------------------------------------
// N2 integrator: the same adaptive Rosenbrock method as N1, applied to the (nc+1)*zNo system.
// W = I/(h*gamma) - J is block lower triangular in the node index: dense n x n diagonal
// blocks, a diagonal sub-diagonal block from the upwind differences, and a rank-structured
// remainder from the pressure march which is carried EXACTLY by one running scalar
// (the linearised pressure dP_k).  Every stage is therefore one forward sweep over the nodes
// with an n x n solve per node — no approximation of the Jacobian, as Rosenbrock methods need.
// Per-instance work arrays live in global memory, [.][thread of the block], so that the lanes of a warp read
// consecutive doubles (layout: see rmt_n2_solve).
// ---------------------------------------------------------------------------------
struct SolveArgsN2 {
    const double* consts;
    i64 B;
    int zNo, tNo;
    double period;
    double rtol, atol;
    int max_steps;
    int out_mode;
    double* out;               // [tNo][rows][zNo][B]
    int* status;               // [B]
    int* stats;                // [4][B]
    double* work;
    unsigned long long* queue;
    double ctrl[6];
};

// rows of the per-node work record
enum {
    W_Y0 = 0, W_Y1 = RMT_N, W_K = 2*RMT_N, W_LU = W_K + RMT_ROS_S*RMT_N, W_L = W_LU + RMT_N*RMT_N,
    W_G = W_L + RMT_N, W_E = W_G + RMT_N, W_EP = W_E + RMT_N,
#if defined(RMT_MODEL_M9)
    W_GV = W_EP + 1, W_LT = W_GV + RMT_N, W_EV = W_LT + RMT_N, W_S4 = W_EV + RMT_N,     // velocity-march couplings
    W_ROWS = W_S4 + 4
#else
    W_ROWS = W_EP + 1
#endif
};                               // W_LU.. W_EP: used by M9 only (N2 keeps these in shared memory, S_* rows below)

// rows of the shared-memory record of the N2 substitution (see rmt_n2_solve); the host sizes the dynamic shared
// memory as N2_SH_ROWS*(RMT_BLOCK + 1) doubles = ((n + 1) n + 3 n + 1)*(block + 1)*8 bytes
enum { S_AUG = 0, S_L = (RMT_N + 1)*RMT_N, S_G = S_L + RMT_N, S_R = S_G + RMT_N, S_DPF = S_R + RMT_N, N2_SH_ROWS = S_DPF + 1 };

// solve with LU factors held in registers (rows permuted in place, reciprocal pivots on the diagonal)
__device__ __forceinline__ void n2_lu_solve(const double (&A)[RMT_N][RMT_N], const int (&perm)[RMT_N],
                                            const double (&b)[RMT_N], double (&x)[RMT_N])
{
#pragma unroll
    for (int r = 0; r < RMT_N; ++r) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < RMT_N; ++q) if (q == perm[r]) v = b[q];
#pragma unroll
        for (int c = 0; c < r; ++c) v -= A[r][c]*x[c];
        x[r] = v;
    }
#pragma unroll
    for (int r = RMT_N - 1; r >= 0; --r) {
        double v = x[r];
#pragma unroll
        for (int c = r + 1; c < RMT_N; ++c) v -= A[r][c]*x[c];
        x[r] = v*A[r][r];
    }
}

// column c of the inverse: the same solve for the unit vector e_c — the permuted right-hand side is one compare per row
__device__ __forceinline__ void n2_lu_solve_unit(const double (&A)[RMT_N][RMT_N], const int (&perm)[RMT_N],
                                                 const int c, double (&x)[RMT_N])
{
#pragma unroll
    for (int r = 0; r < RMT_N; ++r) {
        double v = perm[r] == c ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < r; ++q) v -= A[r][q]*x[q];
        x[r] = v;
    }
#pragma unroll
    for (int r = RMT_N - 1; r >= 0; --r) {
        double v = x[r];
#pragma unroll
        for (int q = r + 1; q < RMT_N; ++q) v -= A[r][q]*x[q];
        x[r] = v*A[r][r];
    }
}

__device__ __forceinline__ int n2_out_rows(const int mode) { return mode == 2 ? 2*RMT_N + RMT_NC : RMT_N; }

// Lanes per reactor.  G consecutive lanes serve one reactor; node k belongs to lane k % G of node group
// k / G.  Per node group the physics (rates, Jacobian blocks, LU) of the G nodes is evaluated in parallel,
// one node per lane; what is sequential in the node index — the Ergun pressure march and the block forward
// substitution — is handed from lane to lane with shuffles, in node order, so the arithmetic (and every
// result bit) is the same for every G.
static_assert(RMT_N2_G >= 0 && RMT_N2_G <= 32 && (RMT_N2_G & (RMT_N2_G - 1)) == 0, "RMT_N2_G: 0 or a power of two <= 32");
static_assert(RMT_BLOCK % 32 == 0, "block size");
#if !RMT_N2_WF

// mixture molar mass [kg/mol] of a node state — the same arithmetic as rmt_point
__device__ __forceinline__ double n2_mw(const double (&u)[RMT_N], const Hot& h)
{
    double C[RMT_NC], S = 0.0;
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) { C[i] = rmt_clamp_eps(u[i])*h.Cmax; S += C[i]; }
    const double invS = rmt_rcp(S);
    double mw = 0.0;
#pragma unroll
    for (int i = 0; i < RMT_NC; ++i) mw += (C[i]*invS)*RMT_cMW[i];
    return mw*1e-3;
}

// Ergun march across the lanes of a group (:3970-3979): Pin = pressure at the group's first node; returns the
// pressure at this lane's node, Pout = pressure at the first node of the next group.
__device__ __forceinline__ double n2_pressure_chain(const double Pin, const double MWm, const double T, const Hot& h,
                                                    const double dz, const int g, const unsigned gmask, double& Pout)
{
    constexpr int G = RMT_N2_G;
    const double us = h.us0;
    const double invRT = rmt_rcp(RMT_R_CONST*T);
    double P = Pin, Pn = Pin;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const double rho = (P*MWm)*invRT;
        const double E = -1*(h.ergA*us + h.ergC*rho*(us*us));
        Pn = fma(E, dz, P);
        if (j + 1 < G) {
            const double up = __shfl_up_sync(gmask, Pn, 1, G);
            if (g == j + 1) P = up;
        }
    }
    Pout = G > 1 ? __shfl_sync(gmask, Pn, G - 1, G) : Pn;
    return P;
}

// One node group of a sweep: pressures (M9: and velocities) at the nodes, then the node functions.  Pg (vg) enter as
// the values at the group's first node and leave as those of the next group's.
template <bool JAC, class AS>
__device__ __forceinline__ void dyn_eval(const double (&u)[RMT_N], const double (&ub)[RMT_N], const bool inlet,
                                         double& Pg, double& vg, const double dz, const double invdz, const Hot& h,
                                         const int g, const unsigned gmask, double (&f)[RMT_N], NodeJac& nj, AS&& asink)
{
    double E;
#if defined(RMT_MODEL_M9)
    static_assert(RMT_N2_G == 1, "M9: the velocity march runs through the kinetics, one lane per reactor");
    double V;
    m9_node<JAC>(u, ub, inlet, Pg, vg, invdz, h, f, E, V, nj);
    Pg = E*dz + Pg;                                                                // pbReactor.py:2546
    vg = V*dz + vg;                                                                // pbReactor.py:2612
    (void)g; (void)gmask; (void)asink;
#else
#if RMT_ISO
    const double Tn = 0.0*h.Tf + h.Tf;
#else
    const double Tn = u[RMT_ITN]*h.Tf + h.Tf;
#endif
    double Pnext;
    const double P = n2_pressure_chain(Pg, n2_mw(u, h), Tn, h, dz, g, gmask, Pnext);
    n2_node<JAC>(u, ub, inlet, P, invdz, h, f, E, nj, asink);
    Pg = Pnext;
    (void)vg;
#endif
}

// lockstep of the block: a barrier before every stage sweep of a node group (1) or only once per node group (0)
#ifndef RMT_N2_STAGE_SYNC
#define RMT_N2_STAGE_SYNC 0
#endif
#ifndef RMT_N2_MINBLOCKS
#define RMT_N2_MINBLOCKS 1
#endif
// stage-vector loads of the stage sweeps: 1 = loop over the earlier stages, 0 = predicated loads of all of them
#ifndef RMT_N2_KLOOP
#define RMT_N2_KLOOP 0
#endif
// hand-over vector of the substitution: 1 = gathered through shared memory (one store, wide loads), 0 = by shuffles
#ifndef RMT_N2_TVSMEM
#define RMT_N2_TVSMEM 1
#endif
extern "C" __global__ void __launch_bounds__(RMT_BLOCK, RMT_N2_MINBLOCKS) rmt_n2_solve(const SolveArgsN2 a)
{
    constexpr int G = RMT_N2_G;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int g = threadIdx.x & (G - 1);                       // lane within the reactor's group
    const unsigned gmask = (G == 32) ? FULL : (((1u << G) - 1u) << (lane - g));
    const i64 threads = (i64)gridDim.x*blockDim.x;
    const i64 tid = (i64)blockIdx.x*blockDim.x + threadIdx.x;
    const int zNo = a.zNo, NG = (zNo + G - 1)/G;
    // Work layout per block: state [node group][y_n | y_{n+1}][thread] followed by scratch [row][thread].  Every row
    // is a compile-time offset from one base pointer, so an access costs no address arithmetic (a [row][node][thread]
    // layout over the whole launch spent a quarter of the kernel's instructions on 64-bit index multiplies).  The
    // state rows exist per node group; everything else (stage vectors, inverse blocks, couplings) is scratch of the
    // group being processed and is re-used from group to group (group-major sweep order), which keeps it resident
    // in L2: per thread NG*2n + (W_ROWS - 2n) doubles.
    double* const wy = a.work + (i64)blockIdx.x*(((i64)NG*W_K + (W_ROWS - W_K))*RMT_BLOCK) + threadIdx.x;
    double* const wsx = wy + (i64)NG*W_K*RMT_BLOCK;
#define WY(row, kg) wy[((i64)(kg)*W_K + (row))*RMT_BLOCK]
#define WS(row) wsx[((row) - W_K)*RMT_BLOCK]
#if !defined(RMT_MODEL_M9)
    const double dz = 1.0/(zNo - 1), invdz = 1.0/dz;
#endif
    const double ISAFE = 1.0/a.ctrl[0], FAC1 = a.ctrl[1], FAC2 = 1.0/a.ctrl[2], KAPPA = a.ctrl[3], BETA = a.ctrl[4];
    const double inv_nz = 1.0/((double)RMT_N*zNo);             // 1/(unknowns per reactor): rms norms
    // per reactor and sweep (s stage sweeps + the Jacobian sweep): upwind state [n], K of the node before [n], marched
    // pressure, linearised pressure, marched velocity, linearised velocity (M9)
    constexpr int N2_CARRY = 2*RMT_N + 4;
#if RMT_N2_G == 1
    double n2_carry[(RMT_ROS_S + 1)*N2_CARRY];                                     // one lane per reactor: thread-private
#endif
#if !defined(RMT_MODEL_M9)
    // What the block forward substitution reads of a node — the augmented inverse block ((n + 1) x n: W_kk^{-1} and
    // the pressure row dz e_k^T W_kk^{-1}), the couplings L_k and g_k, the stage right-hand side (replaced by K_k) and
    // the factor 1 + dz ep_k — lives in shared memory, [row][thread] with the row stride padded to RMT_BLOCK + 1: in
    // the substitution lane r reads row r of ANOTHER lane's column, and with the padded stride those reads fall into
    // different banks (n = 7, 8 lanes per reactor: no conflict at all); the owner's own accesses are conflict-free
    // as before.  Size: N2_SH_ROWS*(RMT_BLOCK + 1) doubles, passed by the host as dynamic shared memory.
    extern __shared__ double n2_sh[];
    constexpr int SH_LD = RMT_BLOCK + 1;
    double* const shc = n2_sh + threadIdx.x;
#define SH(row) shc[(row)*SH_LD]
#if RMT_N2_G > 1
    // behind it: what the sweeps carry from one node group to the next, one record per reactor of the block
    double* const n2_carry = n2_sh + N2_SH_ROWS*SH_LD;        // [(RMT_BLOCK/G)][RMT_ROS_S + 1][N2_CARRY]
#if RMT_N2_TVSMEM
    // the substitution's hand-over vector, 8 doubles per reactor, 16-byte aligned (the carry records before it hold an even
    // number of doubles per reactor and N2_SH_ROWS*SH_LD is padded to even below)
    double* const n2_tv = n2_sh + ((N2_SH_ROWS*SH_LD + (RMT_BLOCK/G)*((RMT_ROS_S + 1)*N2_CARRY) + 1) & ~1);
#endif
#endif
#endif
    i64 inst = -1;
    bool exhausted = false;
    Hot h = {};
    double t = 0.0, hstep = 0.0, hacc = 0.0, erracc = 1e-2, tend = 0.0;
    int nacc = 0, nrej = 0, nanrej = 0, slab = 0, cur = 0;     // cur: which of Y0/Y1 holds y_n
    bool last_rejected = false, fresh = false;

    while (true) {
        // ---- groups without a reactor pull one from the queue (one atomic per warp) ----
        const bool need = (inst < 0) && !exhausted;              // uniform within a group
        const unsigned m = __ballot_sync(FULL, need && g == 0);
        if (m) {
            unsigned long long base = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(m));
            base = __shfl_sync(FULL, base, leader);
            if (need) {
                const i64 cand = (i64)base + __popc(m & ((1u << (lane - g)) - 1));
                if (cand >= a.B) exhausted = true;
                else {
                    inst = cand;
                    rmt_load_hot(a.consts, a.B, inst, h);
                    for (int kg = 0; kg < NG; ++kg) {                    // IV: feed composition at every node, T-hat = 0 (:3483-3497)
#pragma unroll
                        for (int v = 0; v < RMT_NC; ++v) WY(W_Y0 + v, kg) = h.iv[v];
#if defined(RMT_MODEL_M9)
                        WY(W_Y0 + RMT_ITN, kg) = h.Tf;                   // pbReactor.py:2099-2100
#elif !RMT_ISO
                        WY(W_Y0 + RMT_ITN, kg) = 0.0;
#endif
                    }
                    t = 0.0; nacc = nrej = nanrej = 0; slab = 0; cur = 0; last_rejected = false; fresh = true;
                    hacc = 0.0; erracc = 1e-2;
                    tend = a.period/a.tNo;
                }
            }
        }
        // Lockstep: all warps of the block walk the same node loops (same zNo, same stage count), so a block
        // barrier per node group keeps them on the same instructions and lets them share instruction-cache
        // lines (the sweep body is far larger than the I-cache).  Groups without a reactor run as ghosts on
        // their own work slots; every write to out / status / stats is predicated on `live`.  Lanes whose node
        // index is past the outlet (zNo not a multiple of G) integrate padding nodes that nothing reads.
        if (__syncthreads_and(inst < 0)) break;
        const bool live = inst >= 0;
        const int YN = cur ? W_Y1 : W_Y0, YP = cur ? W_Y0 : W_Y1;
#if defined(RMT_MODEL_M9)
        const double dz = h.zf/(zNo - 1), invdz = 1.0/dz;       // dimensional grid, pbReactor.py:2076
#endif

        if (fresh) {
            // starting step from ||y0|| / ||f(y0)|| (Hairer-Wanner II.4, first guess), scaled like N1
            double d0 = 0.0, d1 = 0.0, Pg = h.Pf, vg = h.us0;
            double carry[RMT_N] = {0}, ub[RMT_N], u[RMT_N], fo[RMT_N];
            NodeJac nj;
            for (int kg = 0; kg < NG; ++kg) {
#pragma unroll
                for (int v = 0; v < RMT_N; ++v) {
                    u[v] = WY(YN + v, kg);
                    const double up = G > 1 ? __shfl_up_sync(gmask, u[v], 1, G) : 0.0;
                    ub[v] = g == 0 ? carry[v] : up;
                    carry[v] = G > 1 ? __shfl_sync(gmask, u[v], G - 1, G) : u[v];
                }
                dyn_eval<false>(u, ub, kg == 0 && g == 0, Pg, vg, dz, invdz, h, g, gmask, fo, nj, NoSink());
                double n0 = 0.0, n1 = 0.0;
#pragma unroll
                for (int v = 0; v < RMT_N; ++v) {
                    const double isc = rmt_rcp(KAPPA*(a.atol + a.rtol*fabs(u[v])));
                    n0 += (u[v]*isc)*(u[v]*isc); n1 += (fo[v]*isc)*(fo[v]*isc);
                }
                if (kg*G + g >= zNo) { n0 = 0.0; n1 = 0.0; }
#pragma unroll
                for (int j = 0; j < G; ++j) {                            // node order, the same sum for every G
                    d0 += G > 1 ? __shfl_sync(gmask, n0, j, G) : n0;
                    d1 += G > 1 ? __shfl_sync(gmask, n1, j, G) : n1;
                }
            }
            d0 = rmt_sqrt(d0*inv_nz); d1 = rmt_sqrt(d1*inv_nz);
            const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01*d0*rmt_rcp(d1);
            hstep = fmin(a.ctrl[5]*100.0*h0, tend);
            fresh = false;
        }
        const double hlim = tend - t;
        const bool clipped = hstep*1.01 >= hlim;
        const double hh = clipped ? hlim : hstep;
        const double invh = rmt_rcp(hh), dg = invh*(1.0/RMT_ROS_GAMMA);

        // ---- one pass over the node groups.  For each group: the Jacobian sweep (f(y_n), Jacobian blocks, inverse
        // diagonal blocks — all nodes of the group in parallel) and then all stage sweeps, back to back, so that the
        // group's work rows (inverse blocks, couplings, stage vectors) are re-read while they are still in L2 / L1
        // instead of being streamed from HBM once per stage.  What a sweep carries from one group to the next (state
        // and K of the group's last node, marched pressure, linearised pressure) is kept per stage in shared memory.
        double errsum = 0.0;
        bool bad = false;
#if RMT_N2_G > 1
        double* const cs = n2_carry + (threadIdx.x/G)*((RMT_ROS_S + 1)*N2_CARRY);
#else
        double* const cs = n2_carry;
#endif
        for (int t = 0; t <= RMT_ROS_S; ++t) {                 // slot RMT_ROS_S: the Jacobian sweep
            double* c = cs + t*N2_CARRY;
#pragma unroll
            for (int v = 0; v < 2*RMT_N; ++v) c[v] = 0.0;      // upwind state / K of the node before the inlet
            c[2*RMT_N] = h.Pf; c[2*RMT_N + 1] = 0.0; c[2*RMT_N + 2] = h.us0; c[2*RMT_N + 3] = 0.0;
        }
        for (int kg = 0; kg < NG; ++kg) {
        {
            double* const c0 = cs + RMT_ROS_S*N2_CARRY;
            double carry[RMT_N], ub[RMT_N], u[RMT_N], fo[RMT_N];
            NodeJac nj;
            {
                __syncthreads();
                double Pg = c0[2*RMT_N], vg = c0[2*RMT_N + 2];
#pragma unroll
                for (int v = 0; v < RMT_N; ++v) carry[v] = c0[v];
#pragma unroll
                for (int v = 0; v < RMT_N; ++v) {
                    u[v] = WY(YN + v, kg);
                    const double up = G > 1 ? __shfl_up_sync(gmask, u[v], 1, G) : 0.0;
                    ub[v] = g == 0 ? carry[v] : up;
                    carry[v] = G > 1 ? __shfl_sync(gmask, u[v], G - 1, G) : u[v];
                }
#if defined(RMT_MODEL_M9)
                dyn_eval<true>(u, ub, kg == 0 && g == 0, Pg, vg, dz, invdz, h, g, gmask, fo, nj, NoSink());
                // W_kk = I/(h*gamma) - A, LU with partial pivoting in registers
                int perm[RMT_N];
#pragma unroll
                for (int r = 0; r < RMT_N; ++r) {
                    perm[r] = r;
#pragma unroll
                    for (int c = 0; c < RMT_N; ++c) nj.A[r][c] = (r == c ? dg : 0.0) - nj.A[r][c];
                }
#pragma unroll
                for (int c = 0; c < RMT_N; ++c) {
                    double best = fabs(nj.A[c][c]);
                    int bi = c;
#pragma unroll
                    for (int r = c + 1; r < RMT_N; ++r) { const double v = fabs(nj.A[r][c]); if (v > best) { best = v; bi = r; } }
#pragma unroll
                    for (int r = c + 1; r < RMT_N; ++r)
                        if (r == bi) {
#pragma unroll
                            for (int q = 0; q < RMT_N; ++q) { const double tv = nj.A[c][q]; nj.A[c][q] = nj.A[r][q]; nj.A[r][q] = tv; }
                            const int tp = perm[c]; perm[c] = perm[r]; perm[r] = tp;
                        }
                    const double piv = rmt_rcp(nj.A[c][c]);
                    nj.A[c][c] = piv;
#pragma unroll
                    for (int r = c + 1; r < RMT_N; ++r) {
                        const double l = nj.A[r][c]*piv;
                        nj.A[r][c] = l;
#pragma unroll
                        for (int q = c + 1; q < RMT_N; ++q) nj.A[r][q] -= l*nj.A[c][q];
                    }
                }
                // The diagonal blocks are applied as explicit inverses (columns by LU solves of the unit vectors,
                // once per step): K_k = W_kk^{-1} (rhs_k + L_k*K_{k-1} + g_k dP_k) is then a matrix-vector product —
                // seven independent dot products instead of two dependent triangular sweeps, which is what the
                // lane-to-lane hand-over of the stage sweeps waits for.
#pragma unroll 1
                for (int c = 0; c < RMT_N; ++c) {
                    double xs[RMT_N];
                    n2_lu_solve_unit(nj.A, perm, c, xs);
#pragma unroll
                    for (int q = 0; q < RMT_N; ++q) WS(W_LU + q*RMT_N + c) = xs[q];
                }
#else
                // N2: W_kk = I/(h*gamma) - A is written into shared memory as it is computed (the rows of the inverse
                // block, free at this point), factorised there — partial pivoting through per-row pointers held in
                // registers, every access [row pointer + immediate] like the steady-state integrator — and inverted
                // column by column into the work rows W_LU (global, own column, coalesced); the inverse then replaces the
                // factors.  The block never lives in registers (round 1: 98 registers + local-memory row swaps, 1.2e9
                // local loads per launch).
                struct WSink {
                    double* shc; double dg;
                    __device__ __forceinline__ void operator()(int r, int cc, double v) const {
                        shc[(S_AUG + r*RMT_N + cc)*(RMT_BLOCK + 1)] = (r == cc ? dg : 0.0) - v;
                    }
                };
                dyn_eval<true>(u, ub, kg == 0 && g == 0, Pg, vg, dz, invdz, h, g, gmask, fo, nj, WSink{shc, dg});
                double* rowp[RMT_N];
                int perm[RMT_N];
#pragma unroll
                for (int r = 0; r < RMT_N; ++r) { rowp[r] = shc + (S_AUG + r*RMT_N)*SH_LD; perm[r] = r; }
#define LUW(r, q) rowp[r][(q)*SH_LD]
#pragma unroll
                for (int c = 0; c < RMT_N; ++c) {
                    double best = fabs(LUW(c, c));
                    int bi = c;
#pragma unroll
                    for (int r = c + 1; r < RMT_N; ++r) { const double v = fabs(LUW(r, c)); if (v > best) { best = v; bi = r; } }
#pragma unroll
                    for (int r = c + 1; r < RMT_N; ++r)
                        if (r == bi) {
                            double* tr = rowp[c]; rowp[c] = rowp[r]; rowp[r] = tr;
                            const int tp = perm[c]; perm[c] = perm[r]; perm[r] = tp;
                        }
                    const double piv = rmt_rcp(LUW(c, c));
                    LUW(c, c) = piv;                               // reciprocal pivot
                    double urow[RMT_N];
#pragma unroll
                    for (int q = c + 1; q < RMT_N; ++q) urow[q] = LUW(c, q);
#pragma unroll
                    for (int r = c + 1; r < RMT_N; ++r) {
                        const double l = LUW(r, c)*piv;
                        LUW(r, c) = l;
#pragma unroll
                        for (int q = c + 1; q < RMT_N; ++q) LUW(r, q) -= l*urow[q];
                    }
                }
                // explicit inverse (see the M9 branch for why), column by column; the pressure row dz e_k^T W_kk^{-1}
                double wrow[RMT_N];
#pragma unroll
                for (int c = 0; c < RMT_N; ++c) {
                    double xs[RMT_N];
#pragma unroll
                    for (int r = 0; r < RMT_N; ++r) {
                        double v = perm[r] == c ? 1.0 : 0.0;
#pragma unroll
                        for (int q = 0; q < r; ++q) v -= LUW(r, q)*xs[q];
                        xs[r] = v;
                    }
#pragma unroll
                    for (int r = RMT_N - 1; r >= 0; --r) {
                        double v = xs[r];
#pragma unroll
                        for (int q = r + 1; q < RMT_N; ++q) v -= LUW(r, q)*xs[q];
                        xs[r] = v*LUW(r, r);
                    }
                    double we = 0.0;
#pragma unroll
                    for (int q = 0; q < RMT_N; ++q) { WS(W_LU + q*RMT_N + c) = xs[q]; we = fma(nj.e[q], xs[q], we); }
                    wrow[c] = dz*we;
                }
#undef LUW
#pragma unroll
                for (int q = 0; q < RMT_N*RMT_N; ++q) SH(S_AUG + q) = WS(W_LU + q);
#pragma unroll
                for (int c = 0; c < RMT_N; ++c) SH(S_AUG + RMT_N*RMT_N + c) = wrow[c];     // row n: d(dP_{k+1})/d(tv_k)
#endif
#pragma unroll
                for (int r = 0; r < RMT_N; ++r) {
                    WS(W_K + r) = fo[r];                   // stage-1 right-hand side
#if defined(RMT_MODEL_M9)
                    WS(W_L + r) = nj.L[r]; WS(W_G + r) = nj.g[r]; WS(W_E + r) = nj.e[r];
                    WS(W_GV + r) = nj.gv[r]; WS(W_LT + r) = nj.Lt[r]; WS(W_EV + r) = nj.eV[r];
#else
                    SH(S_L + r) = nj.L[r]; SH(S_G + r) = nj.g[r];
#endif
                }
#if defined(RMT_MODEL_M9)
                WS(W_EP) = nj.ep;
#else
                SH(S_DPF) = fma(dz, nj.ep, 1.0);           // d(dP_{k+1})/d(dP_k) besides the path through K_k
#endif
#if defined(RMT_MODEL_M9)
                WS(W_S4) = nj.ev; WS(W_S4 + 1) = nj.eVb; WS(W_S4 + 2) = nj.eVP; WS(W_S4 + 3) = nj.eVv;
#endif
#pragma unroll
                for (int v = 0; v < RMT_N; ++v) c0[v] = carry[v];
                c0[2*RMT_N] = Pg; c0[2*RMT_N + 2] = vg;
            }
        }

        // ---- stage sweeps of this group ----
#pragma unroll 1
        for (int s = 0; s < RMT_ROS_S; ++s) {
            double* const c = cs + s*N2_CARRY;
            double carry[RMT_N], kcarry[RMT_N];
            const bool lastStage = s == RMT_ROS_S - 1;
            {
#if RMT_N2_STAGE_SYNC
                __syncthreads();
#else
                __syncwarp();
#endif
                double Pg = c[2*RMT_N], dPg = c[2*RMT_N + 1], vg = c[2*RMT_N + 2];
#if defined(RMT_MODEL_M9)
                double dvg = c[2*RMT_N + 3];                   // linearised velocity march
#endif
#pragma unroll
                for (int v = 0; v < RMT_N; ++v) { carry[v] = c[v]; kcarry[v] = c[RMT_N + v]; }
                double rhs[RMT_N];
                if (s == 0) {
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) rhs[v] = WS(W_K + v);
                } else {
                    double u[RMT_N], ub[RMT_N], vc[RMT_N];
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) { u[v] = WY(YN + v, kg); vc[v] = 0.0; }
#if RMT_N2_KLOOP
                    // the earlier stages' vectors, one per iteration (s is block-uniform: no divergence); the loads of an
                    // iteration are issued together
                    {
                        const double* kj = &WS(W_K);
                        const double* arow = &RMT_cROS_A[s][0];
                        const double* crow = &RMT_cROS_C[s][0];
#pragma unroll 1
                        for (int j = 0; j < s; ++j) {
                            const double aj = arow[j], cj = crow[j]*invh;
                            double kv[RMT_N];
#pragma unroll
                            for (int v = 0; v < RMT_N; ++v) kv[v] = kj[v*RMT_BLOCK];
#pragma unroll
                            for (int v = 0; v < RMT_N; ++v) { u[v] += aj*kv[v]; vc[v] += cj*kv[v]; }
                            kj += RMT_N*RMT_BLOCK;
                        }
                    }
#else
                    // (compile-time trip count with predicated loads: the stage vectors of all earlier stages are
                    // fetched back to back instead of one stage per loop iteration)
#pragma unroll
                    for (int j = 0; j < RMT_ROS_S - 1; ++j) {
                        const bool on = j < s;
                        const double aj = on ? RMT_cROS_A[s][j] : 0.0, cj = on ? RMT_cROS_C[s][j]*invh : 0.0;
#pragma unroll
                        for (int v = 0; v < RMT_N; ++v) {
                            const double kv = on ? WS(W_K + j*RMT_N + v) : 0.0;
                            u[v] += aj*kv; vc[v] += cj*kv;
                        }
                    }
#endif
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) {
                        const double up = G > 1 ? __shfl_up_sync(gmask, u[v], 1, G) : 0.0;
                        ub[v] = g == 0 ? carry[v] : up;
                        carry[v] = G > 1 ? __shfl_sync(gmask, u[v], G - 1, G) : u[v];
                    }
                    NodeJac njd;
                    dyn_eval<false>(u, ub, kg == 0 && g == 0, Pg, vg, dz, invdz, h, g, gmask, rhs, njd, NoSink());
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) rhs[v] += vc[v];
                }
#if !defined(RMT_MODEL_M9)
                // ---- block forward substitution of this group, rows spread over the lanes ----
                // K_k = W_kk^{-1} tv_k,  tv_k = rhs_k + L_k*K_{k-1} + g_k dP_k,  dP_{k+1} = (1 + dz ep_k) dP_k + (dz e_k^T W_kk^{-1}) tv_k.
                // The nodes of the group are visited in order (that is inherent to the upwind coupling), but the n + 1
                // rows of a node's update (n of K_k, one of dP_{k+1}: row n of the augmented inverse block) are
                // independent dot products with the same vector tv_k: lane r of the group takes rows r, r + G, ... of
                // EVERY node — it reads them from the owner's work column — so a hand-over costs each lane one dot
                // product instead of the whole matrix-vector product (round 1: all G lanes computed every node's
                // product and kept one, 47 % of the kernel's instructions).  tv_k is assembled from the lanes' own
                // entries with n shuffles; the finished K_k goes back into the owner's K_s slot.  Same arithmetic, in
                // the same order, for every G.
                constexpr int NROW = (RMT_N + 1 + G - 1)/G;               // rows per lane
                double x[RMT_N];
                {
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) SH(S_R + v) = rhs[v];
                    __syncwarp();
                    double kprev[NROW], dP = dPg;
#pragma unroll
                    for (int q = 0; q < NROW; ++q) { const int r = g + q*G; kprev[q] = r < RMT_N ? c[RMT_N + r] : 0.0; }
#pragma unroll 1
                    for (int j = 0; j < G; ++j) {
                        double* const sj = shc + (j - g);                  // column of the lane that owns node j
                        double tvo[NROW];
#pragma unroll
                        for (int q = 0; q < NROW; ++q) {
                            const int r = g + q*G;
                            tvo[q] = 0.0;
                            if (r < RMT_N)
                                tvo[q] = fma(sj[(S_L + r)*SH_LD], kprev[q], fma(sj[(S_G + r)*SH_LD], dP, sj[(S_R + r)*SH_LD]));
                        }
                        double tv[RMT_N];
#if RMT_N2_TVSMEM && RMT_N2_G >= 8
                        if (NROW == 1) {
                            // every lane stores its entry into the reactor's 8-double record and reads the whole vector back
                            // with four 128-bit loads (all lanes of a reactor read the same addresses: broadcast)
                            double* const tvs = n2_tv + (threadIdx.x/G)*8;
                            if (g < 8) tvs[g] = g < RMT_N ? tvo[0] : 0.0;
                            __syncwarp();
                            const double2* t2 = reinterpret_cast<const double2*>(tvs);
#pragma unroll
                            for (int cc = 0; cc < RMT_N; cc += 2) {
                                const double2 w = t2[cc/2];
                                tv[cc] = w.x;
                                if (cc + 1 < RMT_N) tv[cc + 1] = w.y;
                            }
                            __syncwarp();
                        } else
#endif
#pragma unroll
                        for (int cc = 0; cc < RMT_N; ++cc) tv[cc] = G > 1 ? __shfl_sync(gmask, tvo[cc/G], cc % G, G) : tvo[cc/G];
                        double dPn = 0.0;
#pragma unroll
                        for (int q = 0; q < NROW; ++q) {
                            const int r = g + q*G;
                            if (r <= RMT_N) {
                                const double* wr = sj + (S_AUG + r*RMT_N)*SH_LD;
                                // two partial sums (even / odd columns): half the dependent chain of the hand-over
                                double acc = wr[0]*tv[0], acc1 = RMT_N > 1 ? wr[SH_LD]*tv[RMT_N > 1 ? 1 : 0] : 0.0;
#pragma unroll
                                for (int cc = 2; cc < RMT_N; cc += 2) acc = fma(wr[cc*SH_LD], tv[cc], acc);
#pragma unroll
                                for (int cc = 3; cc < RMT_N; cc += 2) acc1 = fma(wr[cc*SH_LD], tv[cc], acc1);
                                acc += acc1;
                                if (r < RMT_N) { kprev[q] = acc; sj[(S_R + r)*SH_LD] = acc; }
                                else dPn = fma(dP, sj[S_DPF*SH_LD], acc);
                            }
                        }
                        dP = G > 1 ? __shfl_sync(gmask, dPn, RMT_N % G, G) : dPn;
                    }
                    __syncwarp();
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) x[v] = SH(S_R + v);
                    // hand-over to the next group: K of the group's last node (held row-wise), linearised pressure
#pragma unroll
                    for (int q = 0; q < NROW; ++q) { const int r = g + q*G; if (r < RMT_N) c[RMT_N + r] = kprev[q]; }
                    dPg = dP;
                }
#else
                // x0 = W_kk^{-1} rhs_k: all nodes of the group in parallel
                double Wi[RMT_N][RMT_N], Lk[RMT_N], gk[RMT_N], x0[RMT_N];
#pragma unroll
                for (int r = 0; r < RMT_N; ++r) {
                    Lk[r] = WS(W_L + r); gk[r] = WS(W_G + r);
#pragma unroll
                    for (int c = 0; c < RMT_N; ++c) Wi[r][c] = WS(W_LU + r*RMT_N + c);
                }
#pragma unroll
                for (int r = 0; r < RMT_N; ++r) {
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < RMT_N; ++c) acc += Wi[r][c]*rhs[c];
                    x0[r] = acc;
                }
                // Block forward substitution, node by node = lane by lane: lane j adds what the node before it
                // contributes (upwind block and the pressure column), K_k = x0 + W_kk^{-1}(L_k*K_{k-1} + g_k dP_k),
                // and hands K_k and the linearised pressure dP_{k+1} = dP_k + dz*(e_k . K_k + ep_k*dP_k) to lane j + 1.
                double x[RMT_N], kp[RMT_N], dP = dPg, dPn = dPg;
#pragma unroll
                for (int v = 0; v < RMT_N; ++v) { kp[v] = kcarry[v]; x[v] = 0.0; }
#pragma unroll 1
                for (int j = 0; j < G; ++j) {
                    double xx[RMT_N], tv[RMT_N];
#if defined(RMT_MODEL_M9)
                    // M9: besides the upwind block and the pressure column, the velocity column and the T_{k-1} column
#pragma unroll
                    for (int c = 0; c < RMT_N; ++c)
                        tv[c] = fma(Lk[c], kp[c], gk[c]*dP) + (WS(W_GV + c)*dvg + WS(W_LT + c)*kp[RMT_ITN]);
#else
#pragma unroll
                    for (int c = 0; c < RMT_N; ++c) tv[c] = fma(Lk[c], kp[c], gk[c]*dP);     // explicit: the same bits for every G
#endif
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) {
                        double acc = x0[v];
#pragma unroll
                        for (int c = 0; c < RMT_N; ++c) acc += Wi[v][c]*tv[c];
                        xx[v] = acc;
                    }
                    double ek = 0.0;
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) ek += WS(W_E + v)*xx[v];
#if defined(RMT_MODEL_M9)
                    // dP_{k+1} = dP_k + dz*(e_k.K_k + (dE/dv) dv_k);  dv_{k+1} = dv_k + dz*(eV_k.K_k + (dV/dT_{k-1}) K_{k-1,T}
                    //            + (dV/dP) dP_k + (dV/dv) dv_k)
                    double evk = 0.0;
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) evk += WS(W_EV + v)*xx[v];
                    const double dnx = fma(dz, ek + WS(W_S4)*dvg, dP);
                    dvg = fma(dz, evk + WS(W_S4 + 1)*kp[RMT_ITN] + WS(W_S4 + 2)*dP + WS(W_S4 + 3)*dvg, dvg);
#else
                    const double dnx = fma(dz, fma(WS(W_EP), dP, ek), dP);
#endif
                    if (g == j) {
#pragma unroll
                        for (int v = 0; v < RMT_N; ++v) x[v] = xx[v];
                        dPn = dnx;
                    }
                    if (G > 1) {
                        // lane j's result goes to everybody; only lane j + 1 uses it (in the next iteration)
#pragma unroll
                        for (int v = 0; v < RMT_N; ++v) kp[v] = __shfl_sync(gmask, xx[v], j, G);
                        dP = __shfl_sync(gmask, dnx, j, G);
                    }
                }
#pragma unroll
                for (int v = 0; v < RMT_N; ++v) kcarry[v] = G > 1 ? __shfl_sync(gmask, x[v], G - 1, G) : x[v];
                dPg = G > 1 ? __shfl_sync(gmask, dPn, G - 1, G) : dPn;
#endif   // N2 (rows over lanes) / M9 (one lane per reactor)
                if (!lastStage) {
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) WS(W_K + s*RMT_N + v) = x[v];
                } else {
                    // y_{n+1} = y_n + sum_j m_j K_j ; err = sum_j e_j K_j
                    double ne = 0.0;
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) {
                        const double yo = WY(YN + v, kg);
                        double yn = yo;
                        double ev = 0.0;
#pragma unroll
                        for (int j = 0; j < RMT_ROS_S - 1; ++j) {         // the last stage: s == RMT_ROS_S - 1
                            const double kj = WS(W_K + j*RMT_N + v);
                            yn += RMT_cROS_M[j]*kj; ev += RMT_cROS_E[j]*kj;
                        }
                        yn += RMT_cROS_M[s]*x[v]; ev += RMT_cROS_E[s]*x[v];
                        WY(YP + v, kg) = yn;
                        const double en = ev*rmt_rcp(KAPPA*(a.atol + a.rtol*fmax(fabs(yo), fabs(yn))));
                        ne += en*en;
                        bad = bad || (kg*G + g < zNo && !(fabs(yn) <= 1.7e308));
                    }
                    if (kg*G + g >= zNo) ne = 0.0;
#pragma unroll
                    for (int j = 0; j < G; ++j) errsum += G > 1 ? __shfl_sync(gmask, ne, j, G) : ne;   // node order
                }
#pragma unroll
                for (int v = 0; v < RMT_N; ++v) {
                    c[v] = carry[v];
#if defined(RMT_MODEL_M9)
                    c[RMT_N + v] = kcarry[v];
#endif
                }
                c[2*RMT_N] = Pg; c[2*RMT_N + 1] = dPg; c[2*RMT_N + 2] = vg;
#if defined(RMT_MODEL_M9)
                c[2*RMT_N + 3] = dvg;
#endif
            }
        }
        }   // node groups
        bad = (__ballot_sync(FULL, bad) & gmask) != 0u;
        double err = rmt_sqrt(errsum*inv_nz);
        if (bad || !(err == err)) err = 1e30;

        // ---- controller (same as N1; identical in every lane of the group) ----
        const double errc = fmax(err, 1e-10);
        double fac;
        if (BETA > 0.0 && nacc > 0) fac = rmt_powc(errc, 1.0/(RMT_ROS_ORDER) - 0.75*BETA)*rmt_powc(erracc, -BETA)*ISAFE;
        else fac = rmt_root_order(errc, 1.0/(RMT_ROS_ORDER))*ISAFE;
        fac = fmax(FAC2, fmin(FAC1, fac));
        double hnew = hh*rmt_rcp(fac);
        int fin = -1;
        if (err <= 1.0) {
            if (nacc > 0 && BETA <= 0.0) {
                double facgus = (hacc*invh)*rmt_root_order(err*err*rmt_rcp(erracc), 1.0/(RMT_ROS_ORDER))*ISAFE;
                facgus = fmax(FAC2, fmin(FAC1, facgus));
                fac = fmax(fac, facgus);
                hnew = hh*rmt_rcp(fac);
            }
            hacc = hh; erracc = fmax(1e-2, err);
            ++nacc; nanrej = 0;
            cur ^= 1;
            t = clipped ? tend : t + hh;
            if (last_rejected) hnew = fmin(hnew, hh);
            last_rejected = false;
            hstep = clipped ? fmax(hnew, hstep) : hnew;
            if (t >= tend) {
                // end of a slab: un-scale and store (sortResult5, solResultAnalysis.py:252-301; :3630-3661)
                const int YC = cur ? W_Y1 : W_Y0;
                const int rows = n2_out_rows(a.out_mode);
                for (int kg = 0; kg < NG; ++kg) {
                    const int k = kg*G + g;
                    if (!live || k >= zNo) continue;
                    double v[RMT_N];
#pragma unroll
                    for (int q = 0; q < RMT_N; ++q) v[q] = WY(YC + q, kg);
                    double* o = a.out + (((i64)slab*rows)*zNo + k)*a.B + inst;
                    const i64 rs = (i64)zNo*a.B;
                    if (a.out_mode != 1) {
#pragma unroll
                        for (int q = 0; q < RMT_N; ++q) o[q*rs] = v[q];
                        o += RMT_N*rs;
                    }
                    if (a.out_mode != 0) {
                        double S = 0.0, C[RMT_NC];
#if defined(RMT_MODEL_M9)
#pragma unroll
                        for (int q = 0; q < RMT_NC; ++q) { C[q] = v[q]; S += C[q]; }           // pbReactor.py:2189-2196
#else
#pragma unroll
                        for (int q = 0; q < RMT_NC; ++q) { C[q] = v[q]*h.Cmax; S += C[q]; }
#endif
                        if (a.out_mode == 2) {
#pragma unroll
                            for (int q = 0; q < RMT_NC; ++q) o[q*rs] = C[q];
                            o += RMT_NC*rs;
                        }
#pragma unroll
                        for (int q = 0; q < RMT_NC; ++q) o[q*rs] = C[q]/S;
#if defined(RMT_MODEL_M9)
                        o[RMT_ITN*rs] = v[RMT_ITN];
#elif !RMT_ISO
                        o[RMT_ITN*rs] = v[RMT_ITN]*h.Tf + h.Tf;
#endif
                    }
                }
                ++slab;
                if (slab >= a.tNo) fin = 0;
                else tend = (slab + 1 == a.tNo) ? a.period : a.period*(slab + 1)/a.tNo;
            }
            if (fin < 0 && nacc + nrej >= a.max_steps) fin = 1;
        } else {
            ++nrej;
            if (err >= 1e29) { ++nanrej; hnew = hh*0.1; }
            last_rejected = true;
            hstep = hnew;
            if (nacc + nrej >= a.max_steps) fin = 1;
            else if (hstep < 1e-14*fmax(a.period, 1e-300)) fin = 2;
            else if (nanrej > 30) fin = 3;
        }
        if (fin >= 0) {
            if (live && g == 0) {
                a.status[inst] = fin;
                if (a.stats) {
                    a.stats[inst] = nacc; a.stats[a.B + inst] = nrej;
                    a.stats[2*a.B + inst] = (nacc + nrej)*(RMT_ROS_S - 1); a.stats[3*a.B + inst] = nacc + nrej;
                }
            }
            if (live && fin != 0) {
                const int rows = n2_out_rows(a.out_mode);
                const double qnan = __longlong_as_double(0x7ff8000000000000LL);
                for (int sl = slab; sl < a.tNo; ++sl)
                    for (int r = 0; r < rows; ++r)
                        for (int k = g; k < zNo; k += G) a.out[(((i64)sl*rows + r)*zNo + k)*a.B + inst] = qnan;
            }
            if (live) inst = -1;
        }
    }
#undef WY
#undef WS
#undef SH
}
#endif  // !RMT_N2_WF

#if RMT_N2_WF
// ---------------------------------------------------------------------------------
// N2 integrator, stage pipeline.  Same method, same arithmetic per node as above; different mapping.
//
// The forward substitution over the nodes is sequential (upwind block + linearised pressure), and so are the marches
// of the stage arguments: spreading the NODES of a reactor over lanes makes every one of those chains a lane-to-lane
// hand-over (shuffles, shared-memory transposes, explicit inverse blocks to keep the hand-over short — 2/3 of the
// instructions of the lanes kernel, profiles/r02_n2_lanes_line_profile.txt).  Here the parallelism is over the
// STAGES instead.  A block serves 32*WF_RW reactors, one per lane of WF_RW "reactor warps", with WF_ROLES warps
// per reactor warp, one role each —
//   role 0   "JA": f(y_n) and the Jacobian blocks of node k, written as W_kk = I/(h gamma) - A into a ring slot
//   role r   (r >= 1) the WF_SPR stages (r-1)*WF_SPR + 1 ... r*WF_SPR of node k, one after the other: stage argument,
//            f, right-hand side, LU solve.  Role 1 first factorises W_kk in the ring slot (partial pivoting; its
//            first stage takes f(y_n) from JA); the last role forms y_{n+1} and the error norm.
// — all walking the nodes in flow direction, skewed by one node per role: at block step tau role r works on node
// tau - r.  Every chain (upwind stage state, K of the node before, marched and linearised pressure) is then private
// to ONE thread and lives in its registers; a solve is a plain LU solve by the thread that needs it.  What crosses
// roles goes through memory one block step later: the factorised block, its pivot order and the coupling vectors
// through a shared-memory ring of WF_DEPTH node slots per reactor, the stage vectors K_j[k] through a ring of global
// work rows (L1/L2 resident), y_n / y_{n+1} through the two state rows.  One block barrier per step; no shuffles,
// no redundant arithmetic.  The warps of one role (WF_RW of them) run the same instructions at the same time and
// share the fetched instruction lines.
//
// Control: the per-reactor integration state (t, h, counters, controller memory) lives in the registers of the LAST
// role's thread of that reactor's lane, which is the one that knows the error norm; it publishes what the other roles
// need for the next attempt (live / fresh / which state row is y_n / step size / new instance id / pending output)
// in a small shared-memory record.  A fresh reactor's first attempt is the norm pass for the starting step (JA only).
// Lanes pick up a new reactor from the global queue as soon as theirs is done.
// ---------------------------------------------------------------------------------
// M9 (the dimensional twin, pbReactor.py:2296-2660) runs in the same pipeline: its node function marches the superficial
// velocity as well, so a role's chain carries (P, v) and the linearised pair (dP, dv), and the ring slot holds the velocity
// and T_{k-1} couplings (NodeJac.gv / Lt / eV / ev / eVb / eVP / eVv) besides g, e and the upwind diagonal, which depends
// on the node's velocity and is stored instead of being rebuilt from a mask.  Because every chain is private to a thread,
// the velocity march needs no lane hand-over here — the reason the lanes kernel serves M9 with one lane per reactor.
#ifndef WF_SPR
#define WF_SPR 2                                   // stages per role
#endif
#define WF_ROLES (1 + RMT_ROS_S/WF_SPR)
#define WF_RW (RMT_BLOCK/(32*WF_ROLES))            // reactor warps per block
#define WF_NR (32*WF_RW)                           // reactors per block
#define WF_DEPTH (WF_ROLES <= 4 ? 4 : 8)           // ring slots per reactor: > WF_ROLES - 1, a power of two
static_assert(RMT_ROS_S % WF_SPR == 0, "stage pipeline: stages per role must divide the stage count");
static_assert(RMT_BLOCK == 32*WF_ROLES*WF_RW && WF_RW >= 1, "stage pipeline: block = 32 reactors x roles x reactor warps");
static_assert(WF_DEPTH >= WF_ROLES && (WF_DEPTH & (WF_DEPTH - 1)) == 0, "ring depth");
static_assert(!RMT_ROS_REUSE, "stage pipeline: every stage after the first evaluates f at its own argument");
static_assert(RMT_N <= 8, "pivot order (3 bits per row, bits 0..23) and the upwind mask (bits 24..31) share one 32-bit word");
// rows of a ring slot (per reactor): W_kk -> LU (n x n), g_k, dz e_k, 1 + dz ep_k;
// M9: + L_k, gv_k, Lt_k, dz eV_k, dz ev_k, dz eVb_k, dz eVP_k, 1 + dz eVv_k
#if defined(RMT_MODEL_M9)
enum { E_W = 0, E_G = RMT_N*RMT_N, E_E = E_G + RMT_N, E_EPF = E_E + RMT_N, E_L = E_EPF + 1, E_GV = E_L + RMT_N, E_LT = E_GV + RMT_N,
       E_EV = E_LT + RMT_N, E_S0 = E_EV + RMT_N, E_ROWS = E_S0 + 4 };
#else
enum { E_W = 0, E_G = RMT_N*RMT_N, E_E = E_G + RMT_N, E_EPF = E_E + RMT_N, E_ROWS = E_EPF + 1 };
#endif
// control record (doubles) per reactor lane
enum { C_HH = 0, C_D0, C_D1, C_ROWS };
enum { F_LIVE = 1, F_FRESH = 2, F_CUR = 4, F_OUT = 8, F_NEW = 16, F_OUTBUF = 32 };
// ring of stage vectors K_j[k], j < S - 1: [RW][DEPTH][S-1][n][32] doubles — in shared memory when the block's record
// still fits (WF_KSMEM), else in the block's global work rows
#define WF_KRING_DOUBLES (WF_RW*WF_DEPTH*(RMT_ROS_S - 1)*RMT_N*32)
#define WF_SMEM_BASE_DOUBLES (WF_RW*WF_DEPTH*E_ROWS*32 + WF_RW*2*RMT_N*32 + C_ROWS*WF_NR + WF_NR)
#define WF_SMEM_INTS (WF_RW*WF_DEPTH*32 + 2*WF_NR)
#ifndef WF_KSMEM
#define WF_KSMEM ((8*(WF_SMEM_BASE_DOUBLES + WF_KRING_DOUBLES) + 4*WF_SMEM_INTS) <= 226*1024)
#endif
// shared memory, doubles: ring [RW][DEPTH][E_ROWS][32] | f(y_n) ring [RW][2][n][32] | ctl [C_ROWS][NR] | instance ids [NR]
// | (K ring); then ints: pivot/mask words [RW][DEPTH][32], flags [NR], out_slab [NR]
#define WF_SMEM_DOUBLES (WF_SMEM_BASE_DOUBLES + (WF_KSMEM ? WF_KRING_DOUBLES : 0))
#define WF_SMEM_BYTES (8*WF_SMEM_DOUBLES + 4*WF_SMEM_INTS)
// global work rows per block (doubles): y [RW][2][zNo][n][32] | (K ring)
#define WF_WORK_PER_NODE (WF_RW*2*RMT_N*32)
#define WF_WORK_FIXED (WF_KSMEM ? 0 : WF_KRING_DOUBLES)

extern "C" __global__ void rmt_n2_meta(int* out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    out[0] = WF_SMEM_BYTES; out[1] = WF_NR; out[2] = WF_WORK_PER_NODE; out[3] = WF_WORK_FIXED; out[4] = WF_ROLES; out[5] = WF_SPR;
}

// solve (P W = L U) x = b with the factors of a ring slot (row i of the factors is physical row (pw >> 3 i) & 7,
// reciprocal pivots on the diagonal).  The permuted right-hand side goes through a thread-private column (local memory,
// L1 resident) — the pivot order differs from lane to lane, so it cannot be a register index.
__device__ __forceinline__ void wf_solve(const double* __restrict__ slot, const unsigned pw, const double (&b)[RMT_N], double (&x)[RMT_N])
{
    double scr[RMT_N];
#pragma unroll
    for (int v = 0; v < RMT_N; ++v) scr[v] = b[v];
    const double* rowp[RMT_N];
#pragma unroll
    for (int i = 0; i < RMT_N; ++i) {
        const int pi = (pw >> (3*i)) & 7;
        rowp[i] = slot + (E_W + pi*RMT_N)*32;
        x[i] = scr[pi];
    }
#pragma unroll
    for (int i = 0; i < RMT_N; ++i) {
        double v = x[i];
#pragma unroll
        for (int j = 0; j < i; ++j) v -= rowp[i][j*32]*x[j];
        x[i] = v;
    }
#pragma unroll
    for (int i = RMT_N - 1; i >= 0; --i) {
        double v = x[i];
#pragma unroll
        for (int j = i + 1; j < RMT_N; ++j) v -= rowp[i][j*32]*x[j];
        x[i] = v*rowp[i][i*32];
    }
}

// y_{n+1} equals the last stage's argument plus its K, and the error estimator is that K (Rodas4, Rodas3)?
// (measured: the shortcut saves the last role ~0.3 k instructions per node and still makes the kernel 8 % SLOWER — 0.088 vs
// 0.081 s per full round of 9 472 reactors x 200 nodes — so the generic sums stay the default)
#ifndef WF_SA
#define WF_SA 0
#endif
__device__ constexpr bool wf_stiffly_accurate()
{
    constexpr int S = RMT_ROS_S;
    if (!WF_SA) return false;
    if (S < 2) return false;
    for (int j = 0; j < S - 1; ++j)
        if (RMT_ROS_M[j] != RMT_ROS_A[S - 1][j] || RMT_ROS_E[j] != 0.0) return false;
    return RMT_ROS_M[S - 1] == 1.0 && RMT_ROS_E[S - 1] == 1.0;
}

#ifndef WF_MINBLOCKS
#define WF_MINBLOCKS 1
#endif
extern "C" __global__ void __launch_bounds__(RMT_BLOCK, WF_MINBLOCKS) rmt_n2_solve(const SolveArgsN2 a)
{
    extern __shared__ double wf_sh[];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int role = wid/WF_RW, rw = wid - role*WF_RW;
    const int rl = rw*32 + lane;                                           // reactor slot of the block
    const int zNo = a.zNo;
    double* const ring = wf_sh + rw*(WF_DEPTH*E_ROWS*32) + lane;           // [(slot*E_ROWS + row)*32]
    double* const fring = wf_sh + WF_RW*WF_DEPTH*E_ROWS*32 + rw*(2*RMT_N*32) + lane;      // [((k & 1)*n + v)*32]
    double* const ctl = wf_sh + WF_RW*WF_DEPTH*E_ROWS*32 + WF_RW*2*RMT_N*32 + rl;        // [row*WF_NR]
    i64* const c_inst = reinterpret_cast<i64*>(wf_sh + WF_RW*WF_DEPTH*E_ROWS*32 + WF_RW*2*RMT_N*32 + C_ROWS*WF_NR) + rl;
    int* const ibase = reinterpret_cast<int*>(wf_sh + WF_SMEM_DOUBLES);
    unsigned* const c_perm = reinterpret_cast<unsigned*>(ibase) + rw*(WF_DEPTH*32) + lane;   // [slot*32]
    int* const c_flags = ibase + WF_RW*WF_DEPTH*32 + rl;
    int* const c_oslab = ibase + WF_RW*WF_DEPTH*32 + WF_NR + rl;
    double* const wy = a.work + (i64)blockIdx.x*((i64)WF_WORK_PER_NODE*zNo + WF_WORK_FIXED) + (i64)rw*((i64)2*zNo*RMT_N*32) + lane;
    double* const wk = (WF_KSMEM ? wf_sh + WF_SMEM_BASE_DOUBLES
                                 : a.work + (i64)blockIdx.x*((i64)WF_WORK_PER_NODE*zNo + WF_WORK_FIXED) + (i64)WF_WORK_PER_NODE*zNo)
                       + (i64)rw*(WF_DEPTH*(RMT_ROS_S - 1)*RMT_N*32) + lane;
#define YROW(buf, k, v) wy[(((i64)(buf)*zNo + (k))*RMT_N + (v))*32]
#define KSLOT(k) (wk + ((k) & (WF_DEPTH - 1))*((RMT_ROS_S - 1)*RMT_N*32))     // K_j[k][v] = KSLOT(k)[(j*n + v)*32]
#define SLOT(k) (ring + ((k) & (WF_DEPTH - 1))*(E_ROWS*32))
#if !defined(RMT_MODEL_M9)
    const double dz = 1.0/(zNo - 1), invdz = 1.0/dz;
#endif
    const double ISAFE = 1.0/a.ctrl[0], FAC1 = a.ctrl[1], FAC2 = 1.0/a.ctrl[2], KAPPA = a.ctrl[3], BETA = a.ctrl[4];
    const double inv_nz = 1.0/((double)RMT_N*zNo);
    const bool controller = role == WF_ROLES - 1;

    // every thread of a lane: the reactor it serves
    i64 inst = -1;
    Hot h = {};
    // controller thread only: the integration state of the lane's reactor
    bool live = false, exhausted = false, fresh = false, last_rejected = false, clipped = false;
    double t = 0.0, hstep = 0.0, hacc = 0.0, erracc = 1e-2, tend = 0.0, hh = 1.0;
    int nacc = 0, nrej = 0, nanrej = 0, slab = 0, cur = 0;
    double errsum = 0.0;
    bool bad = false;

    while (true) {
        // ---- [A] controller: result of the attempt just made, refill, step size of the next attempt ----
        if (controller) {
            int flags = 0;
            if (live) {
                int fin = -1;
                if (fresh) {
                    // starting step from ||y0|| / ||f(y0)|| (Hairer-Wanner II.4, first guess), scaled like N1
                    const double d0 = rmt_sqrt(ctl[C_D0*WF_NR]*inv_nz), d1 = rmt_sqrt(ctl[C_D1*WF_NR]*inv_nz);
                    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01*d0*rmt_rcp(d1);
                    hstep = fmin(a.ctrl[5]*100.0*h0, tend);
                    fresh = false;
                } else {
                    const double invh = rmt_rcp(hh);
                    double err = rmt_sqrt(errsum*inv_nz);
                    if (bad || !(err == err)) err = 1e30;
                    const double errc = fmax(err, 1e-10);
                    double fac;
                    if (BETA > 0.0 && nacc > 0) fac = rmt_powc(errc, 1.0/(RMT_ROS_ORDER) - 0.75*BETA)*rmt_powc(erracc, -BETA)*ISAFE;
                    else fac = rmt_root_order(errc, 1.0/(RMT_ROS_ORDER))*ISAFE;
                    fac = fmax(FAC2, fmin(FAC1, fac));
                    double hnew = hh*rmt_rcp(fac);
                    if (err <= 1.0) {
                        if (nacc > 0 && BETA <= 0.0) {
                            double facgus = (hacc*invh)*rmt_root_order(err*err*rmt_rcp(erracc), 1.0/(RMT_ROS_ORDER))*ISAFE;
                            facgus = fmax(FAC2, fmin(FAC1, facgus));
                            fac = fmax(fac, facgus);
                            hnew = hh*rmt_rcp(fac);
                        }
                        hacc = hh; erracc = fmax(1e-2, err);
                        ++nacc; nanrej = 0;
                        cur ^= 1;
                        t = clipped ? tend : t + hh;
                        if (last_rejected) hnew = fmin(hnew, hh);
                        last_rejected = false;
                        hstep = clipped ? fmax(hnew, hstep) : hnew;
                        if (t >= tend) {
                            // end of a slab: every thread of the lane un-scales and stores its share in phase [C]
                            flags |= F_OUT | (cur ? F_OUTBUF : 0);
                            c_oslab[0] = slab;
                            ++slab;
                            if (slab >= a.tNo) fin = 0;
                            else tend = (slab + 1 == a.tNo) ? a.period : a.period*(slab + 1)/a.tNo;
                        }
                        if (fin < 0 && nacc + nrej >= a.max_steps) fin = 1;
                    } else {
                        ++nrej;
                        if (err >= 1e29) { ++nanrej; hnew = hh*0.1; }
                        last_rejected = true;
                        hstep = hnew;
                        if (nacc + nrej >= a.max_steps) fin = 1;
                        else if (hstep < 1e-14*fmax(a.period, 1e-300)) fin = 2;
                        else if (nanrej > 30) fin = 3;
                    }
                }
                if (fin >= 0) {
                    a.status[inst] = fin;
                    if (a.stats) {
                        a.stats[inst] = nacc; a.stats[a.B + inst] = nrej;
                        a.stats[2*a.B + inst] = (nacc + nrej)*(RMT_ROS_S - 1); a.stats[3*a.B + inst] = nacc + nrej;
                    }
                    if (fin != 0) {
                        const int rows = n2_out_rows(a.out_mode);
                        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
                        for (int sl = slab; sl < a.tNo; ++sl)
                            for (int r = 0; r < rows; ++r)
                                for (int k = 0; k < zNo; ++k) a.out[(((i64)sl*rows + r)*zNo + k)*a.B + inst] = qnan;
                    }
                    live = false;
                }
            }
            // lanes without a reactor pull one from the queue (one atomic per warp)
            const bool need = !live && !exhausted;
            const unsigned m = __ballot_sync(FULL, need);
            if (m) {
                unsigned long long base = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(m));
                base = __shfl_sync(FULL, base, leader);
                if (need) {
                    const i64 cand = (i64)base + __popc(m & ((1u << lane) - 1));
                    if (cand >= a.B) exhausted = true;
                    else {
                        c_inst[0] = cand;                                  // (every thread of the lane, this one included, adopts it in [C])
                        flags |= F_NEW;
                        live = true; fresh = true;
                        t = 0.0; nacc = nrej = nanrej = 0; slab = 0; cur = 0; last_rejected = false;
                        hacc = 0.0; erracc = 1e-2;
                        tend = a.period/a.tNo;
                    }
                }
            }
            if (live && !fresh) {
                const double hlim = tend - t;
                clipped = hstep*1.01 >= hlim;
                hh = clipped ? hlim : hstep;
            } else hh = 1.0;
            ctl[C_HH*WF_NR] = hh;
            c_flags[0] = flags | (live ? F_LIVE : 0) | (fresh ? F_FRESH : 0) | (cur ? F_CUR : 0);
            errsum = 0.0; bad = false;
        }
        __syncthreads();
        // ---- [C] every thread: slab output of the reactor it served, then the new reactor ----
        const int flags = c_flags[0];
        if (flags & F_OUT) {
            // un-scale and store (sortResult5, solResultAnalysis.py:252-301; :3630-3661): nodes role, role + roles, ...
            const int ob = (flags & F_OUTBUF) ? 1 : 0, oslab = c_oslab[0];
            const int rows = n2_out_rows(a.out_mode);
            const i64 rs = (i64)zNo*a.B;
            for (int k = role; k < zNo; k += WF_ROLES) {
                double v[RMT_N];
#pragma unroll
                for (int q = 0; q < RMT_N; ++q) v[q] = YROW(ob, k, q);
                double* o = a.out + (((i64)oslab*rows)*zNo + k)*a.B + inst;
                if (a.out_mode != 1) {
#pragma unroll
                    for (int q = 0; q < RMT_N; ++q) o[q*rs] = v[q];
                    o += RMT_N*rs;
                }
                if (a.out_mode != 0) {
                    double S = 0.0, C[RMT_NC];
#pragma unroll
#if defined(RMT_MODEL_M9)
                    for (int q = 0; q < RMT_NC; ++q) { C[q] = v[q]; S += C[q]; }               // pbReactor.py:2189-2196
#else
                    for (int q = 0; q < RMT_NC; ++q) { C[q] = v[q]*h.Cmax; S += C[q]; }
#endif
                    if (a.out_mode == 2) {
#pragma unroll
                        for (int q = 0; q < RMT_NC; ++q) o[q*rs] = C[q];
                        o += RMT_NC*rs;
                    }
#pragma unroll
                    for (int q = 0; q < RMT_NC; ++q) o[q*rs] = C[q]/S;
#if defined(RMT_MODEL_M9)
                    o[RMT_ITN*rs] = v[RMT_ITN];
#elif !RMT_ISO
                    o[RMT_ITN*rs] = v[RMT_ITN]*h.Tf + h.Tf;
#endif
                }
            }
        }
        if (flags & F_NEW) {
            inst = c_inst[0];
            rmt_load_hot(a.consts, a.B, inst, h);
            for (int k = role; k < zNo; k += WF_ROLES) {                   // IV: feed composition at every node, T-hat = 0 (:3483-3497)
#pragma unroll
                for (int v = 0; v < RMT_NC; ++v) YROW(0, k, v) = h.iv[v];
#if defined(RMT_MODEL_M9)
                YROW(0, k, RMT_ITN) = h.Tf;                                // pbReactor.py:2099-2100
#elif !RMT_ISO
                YROW(0, k, RMT_ITN) = 0.0;
#endif
            }
        }
        const bool on = (flags & F_LIVE) != 0;
        if (__syncthreads_and(!on)) break;                                 // also: the initial state is visible to JA
        const bool isfresh = (flags & F_FRESH) != 0;
        const int yn = (flags & F_CUR) ? 1 : 0;
        const double hcur = ctl[C_HH*WF_NR];
        const double invh = rmt_rcp(hcur), dg = invh*(1.0/RMT_ROS_GAMMA);
#if defined(RMT_MODEL_M9)
        const double dz = h.zf/(zNo - 1), invdz = 1.0/dz;                  // dimensional grid, pbReactor.py:2076
#else
        const double Lc = h.F1*invdz;                                      // upwind coupling of a species row
#if !RMT_ISO
        const double Lt = h.invZv*invdz;                                   // ... of the temperature row
#endif
#endif

        // ---- [E] one attempt: zNo + roles - 1 block steps; role r works on node tau - r ----
        // chain state of the role's WF_SPR stages (JA uses set 0 for its upwind state and pressure)
        double ub[WF_SPR][RMT_N], kprev[WF_SPR][RMT_N], P[WF_SPR], dP[WF_SPR], d0 = 0.0, d1 = 0.0;
#if defined(RMT_MODEL_M9)
        double vel[WF_SPR], dv[WF_SPR];                                    // marched / linearised superficial velocity
#pragma unroll
        for (int q = 0; q < WF_SPR; ++q) { vel[q] = h.us0; dv[q] = 0.0; }  // v_z[0] = SuGaVe0, pbReactor.py:2433
#endif
#pragma unroll
        for (int q = 0; q < WF_SPR; ++q) {
            P[q] = h.Pf; dP[q] = 0.0;
#pragma unroll
            for (int v = 0; v < RMT_N; ++v) { ub[q][v] = 0.0; kprev[q][v] = 0.0; }
        }
        const int nsteps = zNo + WF_ROLES - 1;
        for (int tau = 0; tau < nsteps; ++tau) {
            const int k = tau - role;
            if (on && k >= 0 && k < zNo) {
                double* const slot = SLOT(k);
                if (role == 0) {
                    // JA: f(y_n) and the Jacobian blocks of node k; W_kk = I/(h gamma) - A goes straight into the slot
                    struct WSink {
                        double* slot; double dg;
                        __device__ __forceinline__ void operator()(int r, int cc, double v) const {
                            slot[(E_W + r*RMT_N + cc)*32] = (r == cc ? dg : 0.0) - v;
                        }
                    };
                    double u[RMT_N], fo[RMT_N], E;
                    NodeJac nj;
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) u[v] = YROW(yn, k, v);
#if defined(RMT_MODEL_M9)
                    double V;
                    m9_node<true>(u, ub[0], k == 0, P[0], vel[0], invdz, h, fo, E, V, nj);
#pragma unroll
                    for (int r = 0; r < RMT_N; ++r)
#pragma unroll
                        for (int cc = 0; cc < RMT_N; ++cc) slot[(E_W + r*RMT_N + cc)*32] = (r == cc ? dg : 0.0) - nj.A[r][cc];
                    (void)sizeof(WSink);
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) {
                        slot[(E_G + v)*32] = nj.g[v]; slot[(E_E + v)*32] = dz*nj.e[v]; slot[(E_L + v)*32] = nj.L[v];
                        slot[(E_GV + v)*32] = nj.gv[v]; slot[(E_LT + v)*32] = nj.Lt[v]; slot[(E_EV + v)*32] = dz*nj.eV[v];
                        fring[((k & 1)*RMT_N + v)*32] = fo[v];
                    }
                    slot[E_EPF*32] = fma(dz, nj.ep, 1.0);
                    slot[(E_S0 + 0)*32] = dz*nj.ev; slot[(E_S0 + 1)*32] = dz*nj.eVb;
                    slot[(E_S0 + 2)*32] = dz*nj.eVP; slot[(E_S0 + 3)*32] = fma(dz, nj.eVv, 1.0);
                    c_perm[(k & (WF_DEPTH - 1))*32] = 0u;
                    vel[0] = V*dz + vel[0];                                // pbReactor.py:2612
#else
                    n2_node<true>(u, ub[0], k == 0, P[0], invdz, h, fo, E, nj, WSink{slot, dg});
                    unsigned lmask = 0u;                                   // rows whose upwind coupling L_k is switched on (:3897-3904 clamp)
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) {
                        slot[(E_G + v)*32] = nj.g[v]; slot[(E_E + v)*32] = dz*nj.e[v];
                        fring[((k & 1)*RMT_N + v)*32] = fo[v];
                        lmask |= (nj.L[v] != 0.0 ? 1u : 0u) << v;
                    }
                    slot[E_EPF*32] = fma(dz, nj.ep, 1.0);
                    c_perm[(k & (WF_DEPTH - 1))*32] = lmask << 24;         // role 1 adds the pivot order below it
#endif
                    if (isfresh) {
                        double n0 = 0.0, n1 = 0.0;
#pragma unroll
                        for (int v = 0; v < RMT_N; ++v) {
                            const double isc = rmt_rcp(KAPPA*(a.atol + a.rtol*fabs(u[v])));
                            n0 += (u[v]*isc)*(u[v]*isc); n1 += (fo[v]*isc)*(fo[v]*isc);
                        }
                        d0 += n0; d1 += n1;
                        if (k == zNo - 1) { ctl[C_D0*WF_NR] = d0; ctl[C_D1*WF_NR] = d1; }
                    }
                    P[0] = fma(E, dz, P[0]);                               // :3979 (dimensionless dz, kept)
#pragma unroll
                    for (int v = 0; v < RMT_N; ++v) ub[0][v] = u[v];
                } else if (!isfresh) {
                    unsigned pw = c_perm[(k & (WF_DEPTH - 1))*32];
                    if (role == 1) {
                        // LU of W_kk with partial pivoting in the slot (per-row pointers in registers)
                        double* rowp[RMT_N];
                        int perm[RMT_N];
#pragma unroll
                        for (int r = 0; r < RMT_N; ++r) { rowp[r] = slot + (E_W + r*RMT_N)*32; perm[r] = r; }
#define LUW(r, q) rowp[r][(q)*32]
#pragma unroll
                        for (int c = 0; c < RMT_N; ++c) {
                            double best = fabs(LUW(c, c));
                            int bi = c;
#pragma unroll
                            for (int r = c + 1; r < RMT_N; ++r) { const double v = fabs(LUW(r, c)); if (v > best) { best = v; bi = r; } }
#pragma unroll
                            for (int r = c + 1; r < RMT_N; ++r)
                                if (r == bi) {
                                    double* tr = rowp[c]; rowp[c] = rowp[r]; rowp[r] = tr;
                                    const int tp = perm[c]; perm[c] = perm[r]; perm[r] = tp;
                                }
                            const double piv = rmt_rcp(LUW(c, c));
                            LUW(c, c) = piv;                               // reciprocal pivot
                            double urow[RMT_N];
#pragma unroll
                            for (int q = c + 1; q < RMT_N; ++q) urow[q] = LUW(c, q);
#pragma unroll
                            for (int r = c + 1; r < RMT_N; ++r) {
                                const double l = LUW(r, c)*piv;
                                LUW(r, c) = l;
#pragma unroll
                                for (int q = c + 1; q < RMT_N; ++q) LUW(r, q) -= l*urow[q];
                            }
                        }
#undef LUW
#pragma unroll
                        for (int r = 0; r < RMT_N; ++r) pw |= (unsigned)perm[r] << (3*r);
                        c_perm[(k & (WF_DEPTH - 1))*32] = pw;
                    }
                    // the role's stages of node k, one after the other (one copy of the code: the chain sets rotate)
#pragma unroll 1
                    for (int q = 0; q < WF_SPR; ++q) {
                        const int s = (role - 1)*WF_SPR + q;
                        double rhs[RMT_N], x[RMT_N];
                        if (s == 0) {
                            // stage 1: right-hand side f(y_n)
#pragma unroll
                            for (int v = 0; v < RMT_N; ++v) rhs[v] = fring[((k & 1)*RMT_N + v)*32];
                        } else {
                            // argument y_n + sum_j a_sj K_j, f, right-hand side f + sum_j (c_sj / h) K_j
                            double u[RMT_N], vc[RMT_N], E;
#pragma unroll
                            for (int v = 0; v < RMT_N; ++v) { u[v] = YROW(yn, k, v); vc[v] = 0.0; }
                            const double* kj = KSLOT(k);
                            const double* arow = &RMT_cROS_A[s][0];
                            const double* crow = &RMT_cROS_C[s][0];
#pragma unroll 1
                            for (int j = 0; j < s; ++j) {
                                const double aj = arow[j], cj = crow[j]*invh;
#pragma unroll
                                for (int v = 0; v < RMT_N; ++v) {
                                    const double kv = kj[v*32];
                                    u[v] += aj*kv; vc[v] += cj*kv;
                                }
                                kj += RMT_N*32;
                            }
                            NodeJac njd;
#if defined(RMT_MODEL_M9)
                            double V;
                            m9_node<false>(u, ub[0], k == 0, P[0], vel[0], invdz, h, rhs, E, V, njd);
                            vel[0] = V*dz + vel[0];
#else
                            n2_node<false>(u, ub[0], k == 0, P[0], invdz, h, rhs, E, njd, NoSink());
#endif
                            P[0] = fma(E, dz, P[0]);
#pragma unroll
                            for (int v = 0; v < RMT_N; ++v) { ub[0][v] = u[v]; rhs[v] += vc[v]; }
                        }
                        double tv[RMT_N];
#if defined(RMT_MODEL_M9)
                        // besides the upwind block and the pressure column: the velocity column and the T_{k-1} column
                        const double ktb = kprev[0][RMT_ITN];
#pragma unroll
                        for (int v = 0; v < RMT_N; ++v)
                            tv[v] = rhs[v] + (fma(slot[(E_L + v)*32], kprev[0][v], slot[(E_G + v)*32]*dP[0])
                                              + (slot[(E_GV + v)*32]*dv[0] + slot[(E_LT + v)*32]*ktb));
                        wf_solve(slot, pw, tv, x);
                        double ek = 0.0, evk = 0.0;
#pragma unroll
                        for (int v = 0; v < RMT_N; ++v) {
                            kprev[0][v] = x[v];
                            ek = fma(slot[(E_E + v)*32], x[v], ek); evk = fma(slot[(E_EV + v)*32], x[v], evk);
                        }
                        // dP_{k+1} = dP_k + dz (e_k . K_k + (dE/dv) dv_k);  dv_{k+1} = dv_k + dz (eV_k . K_k + (dV/dT_{k-1}) K_{k-1,T}
                        //            + (dV/dP) dP_k + (dV/dv) dv_k)
                        const double dPn = fma(dP[0], slot[E_EPF*32], ek + slot[(E_S0 + 0)*32]*dv[0]);
                        dv[0] = fma(dv[0], slot[(E_S0 + 3)*32], evk + slot[(E_S0 + 1)*32]*ktb + slot[(E_S0 + 2)*32]*dP[0]);
                        dP[0] = dPn;
#else
#pragma unroll
                        for (int v = 0; v < RMT_N; ++v) {
#if !RMT_ISO
                            const double lv = ((pw >> (24 + v)) & 1u) ? (v == RMT_ITN ? Lt : Lc) : 0.0;
#else
                            const double lv = ((pw >> (24 + v)) & 1u) ? Lc : 0.0;
#endif
                            tv[v] = fma(lv, kprev[0][v], fma(slot[(E_G + v)*32], dP[0], rhs[v]));
                        }
                        wf_solve(slot, pw, tv, x);
                        // K_k of this stage: kept for the next node's upwind term, stored for the later stages
                        double ek = 0.0;
#pragma unroll
                        for (int v = 0; v < RMT_N; ++v) { kprev[0][v] = x[v]; ek = fma(slot[(E_E + v)*32], x[v], ek); }
                        dP[0] = fma(dP[0], slot[E_EPF*32], ek);            // dP_{k+1} = (1 + dz ep_k) dP_k + dz e_k . K_k
#endif
                        if (s < RMT_ROS_S - 1) {
                            double* ks = KSLOT(k) + s*(RMT_N*32);
#pragma unroll
                            for (int v = 0; v < RMT_N; ++v) ks[v*32] = x[v];
                        } else {
                            // y_{n+1} = y_n + sum_j m_j K_j ; err = sum_j e_j K_j.  Stiffly accurate tableaux (Rodas4: the last
                            // stage's argument is y_n + sum_{j<S} m_j K_j and the error estimator is K_S): y_{n+1} = u_S + K_S, err = K_S
                            double ne = 0.0;
                            const double* kr = KSLOT(k);
#pragma unroll
                            for (int v = 0; v < RMT_N; ++v) {
                                const double yold = YROW(yn, k, v);
                                double ynew, ev;
                                if (wf_stiffly_accurate()) { ynew = ub[0][v] + x[v]; ev = x[v]; }
                                else {
                                    ynew = yold; ev = 0.0;
#pragma unroll
                                    for (int j = 0; j < RMT_ROS_S - 1; ++j) {
                                        const double kj = kr[(j*RMT_N + v)*32];
                                        ynew += RMT_cROS_M[j]*kj; ev += RMT_cROS_E[j]*kj;
                                    }
                                    ynew += RMT_cROS_M[RMT_ROS_S - 1]*x[v]; ev += RMT_cROS_E[RMT_ROS_S - 1]*x[v];
                                }
                                YROW(yn ^ 1, k, v) = ynew;
                                const double en = ev*rmt_rcp(KAPPA*(a.atol + a.rtol*fmax(fabs(yold), fabs(ynew))));
                                ne += en*en;
                                bad = bad || !(fabs(ynew) <= 1.7e308);
                            }
                            errsum += ne;                                  // node order
                        }
                        if (WF_SPR > 1) {
                            // rotate the chain sets: set 0 is always the one of the stage being worked on
#pragma unroll
                            for (int v = 0; v < RMT_N; ++v) {
                                const double tu = ub[0][v], tk = kprev[0][v];
#pragma unroll
                                for (int w = 0; w + 1 < WF_SPR; ++w) { ub[w][v] = ub[w + 1][v]; kprev[w][v] = kprev[w + 1][v]; }
                                ub[WF_SPR - 1][v] = tu; kprev[WF_SPR - 1][v] = tk;
                            }
                            const double tp = P[0], td = dP[0];
#pragma unroll
                            for (int w = 0; w + 1 < WF_SPR; ++w) { P[w] = P[w + 1]; dP[w] = dP[w + 1]; }
                            P[WF_SPR - 1] = tp; dP[WF_SPR - 1] = td;
#if defined(RMT_MODEL_M9)
                            const double tvl = vel[0], tdv = dv[0];
#pragma unroll
                            for (int w = 0; w + 1 < WF_SPR; ++w) { vel[w] = vel[w + 1]; dv[w] = dv[w + 1]; }
                            vel[WF_SPR - 1] = tvl; dv[WF_SPR - 1] = tdv;
#endif
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
#undef YROW
#undef KSLOT
#undef SLOT
}
#endif  // RMT_N2_WF
#endif  // RMT_DYNAMIC

// ---------------------------------------------------------------------------------
// deterministic objective reduction: per-block (sum, min, argmin) partials, then one
// block folds the partials.  The cross-GPU step is an all-reduce of these 3 numbers.
// ---------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256)
rmt_reduce_partials(const double* __restrict__ v, const i64 n, const i64 index_offset,
                    double* __restrict__ psum, double* __restrict__ pmin, i64* __restrict__ parg)
{
    __shared__ double ssum[256], smin[256];
    __shared__ i64 sarg[256];
    double s = 0.0, mn = __longlong_as_double(0x7ff0000000000000LL);
    i64 am = -1;
    for (i64 i = (i64)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x*blockDim.x) {
        const double x = v[i];
        s += x;
        if (x < mn) { mn = x; am = i + index_offset; }
    }
    ssum[threadIdx.x] = s; smin[threadIdx.x] = mn; sarg[threadIdx.x] = am;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            ssum[threadIdx.x] += ssum[threadIdx.x + w];
            const double o = smin[threadIdx.x + w];
            const i64 oa = sarg[threadIdx.x + w];
            if (o < smin[threadIdx.x] || (o == smin[threadIdx.x] && oa >= 0 && (sarg[threadIdx.x] < 0 || oa < sarg[threadIdx.x]))) {
                smin[threadIdx.x] = o; sarg[threadIdx.x] = oa;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { psum[blockIdx.x] = ssum[0]; pmin[blockIdx.x] = smin[0]; parg[blockIdx.x] = sarg[0]; }
}

// accuracy probe of the branch-free math above: out[0..4][n] = exp, log, sqrt, exp10, 1/x
extern "C" __global__ void rmt_math_probe(const double* __restrict__ x, const int n, double* __restrict__ out)
{
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    out[i] = rmt_exp(v); out[n + i] = rmt_log(v); out[2*n + i] = rmt_sqrt(v);
    out[3*n + i] = rmt_exp10(v); out[4*n + i] = rmt_rcp(v);
}

// ---------------------------------------------------------------------------------
// FP64 pipe peak: dependent-chain-free DFMA loop, used to measure the roofline
// denominator on the box (MEASURED_PEAKS.json has no FP64 figure).
// ---------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256) rmt_dfma_peak(double* __restrict__ out, const int iters, const double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(i64)blockIdx.x*blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
