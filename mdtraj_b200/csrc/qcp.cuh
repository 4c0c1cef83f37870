// qcp.cuh -- per-frame QCP solve in registers (double), shared by every kernel.
//
// Replaces msdFromMandG + DirectSolve (mdtraj/rmsd/src/theobald_rmsd.cpp:217-334,
// :183-193).  Same polynomial, same root, same quaternion-from-cofactors rotation
// and the same identity fallback; what differs is by design:
//   * the reference forms K, C2, C1, C0 in float32 and the final G_a+G_b-2*lambda
//     cancellation in float32 (:275); here everything after the streamed float32
//     partial sums is double, so the result sits closer to the float64 truth than
//     the reference does;
//   * the root comes from Newton-Raphson on the quartic from an upper bound of the
//     largest root (monotone from above), not from the closed-form Ferrari route --
//     no acos/cos/pow on the device.
#pragma once
#include <cuda_runtime.h>

namespace b200 {

struct QcpInput {
    double M[9];  // M[3*i+j] = sum_k a_k[i] * b_k[j], both frames centred
    double Ga, Gb;
    double inv_n;  // 1 / number of atoms, computed once by the caller (a float64 division is ~30 dependent instructions)
};

// Closed form of the largest root, for the inputs Newton cannot be trusted on.
//
// The four roots of the QCP quartic are s1+s2+s3', s1-s2-s3', -s1+s2-s3', -s1-s2+s3' with s1 >= s2 >= s3 >= 0 the
// singular values of M and s3' = sign(det M) * s3.  When s2 + s3' ~ 0 (atoms on a line -- any two-atom selection --
// or an improper-rotation-like pair) the two largest roots (nearly) coincide: Newton then converges linearly, the
// float32 phase is pure rounding noise below a distance of ~1e-4 * lambda and a noisy step can throw the iterate
// under the second root, from where the iteration converges to the wrong one.  The squares mu_i = s_i^2 are the
// eigenvalues of M^T M, i.e. the roots of  mu^3 - ss mu^2 + nc mu - det^2  (ss = |M|_F^2, nc = |cof M|_F^2):
// mu1 comes from the trigonometric form (it is always well separated from 0), mu2 and mu3 from the two symmetric
// functions that have no cancellation (mu2 + mu3 = (nc - det^2/mu1)/mu1, mu2 mu3 = det^2/mu1).  The reference takes
// a closed-form route for every frame (DirectSolve, theobald_rmsd.cpp:183-193) but from float32 coefficients, and is
// itself off by up to 2.5e-2 nm on such inputs; here the closed form is the rare slow path, in float64.
#ifndef QCP_SLOW_PATH_HOOK
#define QCP_SLOW_PATH_HOOK()
#endif
static __device__ __noinline__ double qcp_lambda_closed(const double* M)
{
    QCP_SLOW_PATH_HOOK();
    const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    const double c10 = M[2] * M[7] - M[1] * M[8], c11 = M[0] * M[8] - M[2] * M[6], c12 = M[1] * M[6] - M[0] * M[7];
    const double c20 = M[1] * M[5] - M[2] * M[4], c21 = M[2] * M[3] - M[0] * M[5], c22 = M[0] * M[4] - M[1] * M[3];
    double ss = 0.0;
    for (int i = 0; i < 9; ++i) ss += M[i] * M[i];
    if (!(ss > 0.0)) return 0.0;
    // work with M / |M|_F: every intermediate is O(1) whatever the units
    const double inv = 1.0 / ss;
    const double nc = (c00 * c00 + c01 * c01 + c02 * c02 + c10 * c10 + c11 * c11 + c12 * c12 + c20 * c20 + c21 * c21 +
                       c22 * c22) * inv * inv;
    const double det = (M[0] * c00 + M[1] * c01 + M[2] * c02) * inv / sqrt(ss);
    const double det2 = det * det;
    double p = (1.0 - 3.0 * nc) * (1.0 / 9.0);
    double mu1 = 1.0 / 3.0;
    if (p > 0.0) {
        const double sp = sqrt(p);
        double r = (2.0 - 9.0 * nc + 27.0 * det2) * (1.0 / 54.0) / (p * sp);
        r = fmin(1.0, fmax(-1.0, r));
        mu1 += 2.0 * sp * cos(acos(r) * (1.0 / 3.0));
    }
    const double pq = det2 / mu1;                   // mu2 * mu3
    const double sm = fmax((nc - pq) / mu1, 0.0);   // mu2 + mu3
    const double mu2 = 0.5 * (sm + sqrt(fmax(sm * sm - 4.0 * pq, 0.0)));
    const double mu3 = mu2 > 0.0 ? pq / mu2 : 0.0;
    double lam = sqrt(mu1) + sqrt(mu2) + (det < 0.0 ? -sqrt(mu3) : sqrt(mu3));
    // polish on the (scaled) quartic where it is well conditioned; a step that is not a small correction is discarded
    const double C2 = -2.0, C1 = -8.0 * det, C0 = 1.0 - 4.0 * nc;
    for (int it = 0; it < 2; ++it) {
        const double x2 = lam * lam;
        const double b = (x2 + C2) * lam;
        const double a = b + C1;
        const double den = 2.0 * x2 * lam + b + a;
        const double d = (a * lam + C0) / den;
        if (den > 1e-3 && fabs(d) < 1e-6) lam -= d;
    }
    return lam * sqrt(ss);
}

// Largest root of  t^4 + C2 t^2 + C1 t + C0  (all four roots are real: K is symmetric).
//
// Start: lam0 = min((G_a+G_b)/2, sqrt(3)*||M||_F).  Both are upper bounds of lambda_max (the first is the
// reference's own starting value, theobald_rmsd.cpp:245; the second follows from sum(lambda_i) = 0 and
// lambda_max = s1 + s2 +- s3 <= sqrt(3)*||M||_F for the singular values s of M), and lambda_max >= s1 >=
// ||M||_F/sqrt(3), so the start is never more than 3x above the root -- for dissimilar (e.g. iid random)
// frames (G_a+G_b)/2 alone can be 10-30x above it and Newton would crawl down by 3/4 per step.  Newton from the
// right of the largest root of a real-rooted polynomial is monotone, so the iteration cannot leave the basin.
// The polynomial is scaled by an exact power of two (t = lambda * 2^-e in [1,2) at the start: no division, no
// rounding); the first iterations run in float32 (Laguerre steps: 4-cycle FMA + MUFU square root and reciprocal), the
// last ones are Newton steps in float64 with a float32 reciprocal instead of a DDIV chain.
//
// *trusted is the certificate that the iterate sits on the LARGEST root, and on a simple one: the last step was a
// negligible correction and P'(x) > 1e-5 x^3, P''(x) > 0 (for a real-rooted quartic P'' > 0 puts x right of every inflection point, where P' is
// increasing; P' > 0 there puts x right of every critical point, where P has exactly one root).  Callers send
// uncertified inputs (double or nearly double largest root, see qcp_lambda_closed) to the closed form.
__device__ __forceinline__ double qcp_newton(double C2, double C1, double C0, double lam_upper, double frob2, bool* trusted)
{
    float ub = fminf(sqrtf(3.0f * (float)frob2) * 1.000001f, (float)lam_upper * 1.000001f);
    *trusted = true;
    if (!(ub > 1e-30f)) return 0.0;
    const int e = ((__float_as_int(ub) >> 23) & 0xff) - 127;
    const double s1 = __longlong_as_double((long long)(1023 - e) << 52);
    const double s2 = s1 * s1;
    const double c2 = C2 * s2, c1 = C1 * s2 * s1, c0 = C0 * s2 * s2;
    float t = ub * (float)s1;
    {
        const float f2 = (float)c2, f1 = (float)c1, f0 = (float)c0;
        // Laguerre's iteration (degree 4; see qcp_msd_shift): monotone from above like Newton's for a real-rooted
        // polynomial, but the first step from a bound 3x above the root already lands next to it and convergence is
        // cubic -- 3 steps where Newton took 6 on average and 12 at worst on dissimilar frames.  A step below 1e-3
        // relative leaves ~1e-9; a step <= 0 is rounding noise.  The float64 phase below finishes and certifies.
#pragma unroll 1
        for (int it = 0; it < 12; ++it) {
            const float t2 = t * t;
            const float b = (t2 + f2) * t;
            const float a = b + f1;
            const float den = fmaf(2.0f * t2, t, b + a);              // P'
            const float val = fmaf(a, t, f0);                          // P
            const float hc = fmaf(6.0f, t2, f2);                       // P'' / 2
            const float dd = fmaf(3.0f, sqrtf(fmaxf(fmaf(-2.6666667f * val, hc, den * den), 0.0f)), den);
            if (!(fabsf(dd) > 1e-30f)) break;
            const float delta = __fdividef(4.0f * val, dd);
            t -= delta;
            if (!(delta > 1e-3f * fabsf(t))) break;
        }
        if (!(t > 0.0f) || !(t <= 2.0f)) t = ub * (float)s1;  // numerical accident: restart the float64 phase from the bound
    }
    double x = (double)t;
    bool ok = false;
#pragma unroll 1
    for (int it = 0; it < 12; ++it) {
        const double x2 = x * x;
        const double b = (x2 + c2) * x;
        const double a = b + c1;
        const double den = 2.0 * x2 * x + b + a;
        if (!(den > 1e-300)) break;
        // a float32 reciprocal is enough: a 1e-7 relative error in the step only perturbs the quadratically
        // converging iterate by 1e-7 * |delta|
        const double delta = (a * x + c0) * (double)__frcp_rn((float)den);
        if (!(delta == delta)) break;
        x -= delta;
        if (fabs(delta) <= 1e-10 * fabs(x)) {  // quadratic: the step just taken leaves an error ~delta^2
            // P''(x) > 0, and P'(x) not small against x^3: P'(lambda_1) = (lambda_1 - lambda_2)(..)(..), so a largest
            // root closer than ~1e-5 lambda to the second one (two-atom selections: the pair is double up to rounding,
            // ~1e-8 apart) is left to the closed form -- there the float64 noise of P over P' is no longer negligible
            // against the 1e-8 lambda the RMSD of near-identical frames needs
            ok = 6.0 * x * x > -c2 && den > 1e-5 * x2 * x;
            break;
        }
    }
    *trusted = ok;
    return x * __longlong_as_double((long long)(1023 + e) << 52);
}

// Returns the clamped msd.  If rot != nullptr also writes the row-major rotation
// (applied as row-vector * R, rotation_generic.h:40-42) and returns whether the
// reference's degenerate-quaternion branch (theobald_rmsd.cpp:299-302) was taken.
__device__ __forceinline__ double qcp_solve(const QcpInput& in, float* rot, bool* degenerate)
{
    const double Sxx = in.M[0], Sxy = in.M[1], Sxz = in.M[2];
    const double Syx = in.M[3], Syy = in.M[4], Syz = in.M[5];
    const double Szx = in.M[6], Szy = in.M[7], Szz = in.M[8];

    double k00 = Sxx + Syy + Szz;
    const double k01 = Szy - Syz, k02 = Sxz - Szx, k03 = Syx - Sxy;
    double k11 = Sxx - Syy - Szz;
    const double k12 = Syx + Sxy, k13 = Sxz + Szx;
    double k22 = -Sxx + Syy - Szz;
    const double k23 = Szy + Syz;
    double k33 = -Sxx - Syy + Szz;

    double ss = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) ss += in.M[i] * in.M[i];
    const double C2 = -2.0 * ss;
    const double detM = Sxx * (Syy * Szz - Syz * Szy) + Syx * (Szy * Sxz - Szz * Sxy) + Szx * (Sxy * Syz - Sxz * Syy);
    const double C1 = -8.0 * detM;

    // det(K) by 2x2 minors of the (rows 0,1) x (rows 2,3) Laplace expansion
    const double a01 = k00 * k11 - k01 * k01, a02 = k00 * k12 - k02 * k01, a03 = k00 * k13 - k03 * k01;
    const double a12 = k01 * k12 - k02 * k11, a13 = k01 * k13 - k03 * k11, a23 = k02 * k13 - k03 * k12;
    const double b01 = k02 * k13 - k12 * k03, b02 = k02 * k23 - k22 * k03, b03 = k02 * k33 - k23 * k03;
    const double b12 = k12 * k23 - k22 * k13, b13 = k12 * k33 - k23 * k13, b23 = k22 * k33 - k23 * k23;
    const double C0 = a01 * b23 - a02 * b13 + a03 * b12 + a12 * b03 - a13 * b02 + a23 * b01;

    bool trusted;
    double lam = qcp_newton(C2, C1, C0, 0.5 * (in.Ga + in.Gb), ss, &trusted);
    if (!trusted) {  // rare; the copy keeps `in` itself out of local memory (its address would escape otherwise)
        double m[9] = {Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz};
        lam = qcp_lambda_closed(m);
    }
    double msd = (in.Ga + in.Gb - 2.0 * lam) * in.inv_n;
    if (!(msd > 0.0)) msd = 0.0;

    if (rot != nullptr) {
        k00 -= lam; k11 -= lam; k22 -= lam; k33 -= lam;
        const double m2233 = k22 * k33 - k23 * k23, m1233 = k12 * k33 - k13 * k23, m1223 = k12 * k23 - k13 * k22;
        const double m0223 = k02 * k23 - k03 * k22, m0233 = k02 * k33 - k03 * k23, m0213 = k02 * k13 - k03 * k12;
        double qa = k11 * m2233 - k12 * m1233 + k13 * m1223;
        double qx = -k01 * m2233 + k12 * m0233 - k13 * m0223;
        double qy = k01 * m1233 - k11 * m0233 + k13 * m0213;
        double qz = -k01 * m1223 + k11 * m0223 - k12 * m0213;
        const double n2 = qa * qa + qx * qx + qy * qy + qz * qz;
        const bool degen = n2 < 1e-11;  // same absolute threshold as the reference
        if (degenerate) *degenerate = degen;
        if (degen) {
            rot[0] = rot[4] = rot[8] = 1.0f;
            rot[1] = rot[2] = rot[3] = rot[5] = rot[6] = rot[7] = 0.0f;
        } else {
            double inv = (double)rsqrtf((float)n2);          // float32 seed, two Newton refinements in float64
            inv = inv * (1.5 - 0.5 * n2 * inv * inv);
            inv = inv * (1.5 - 0.5 * n2 * inv * inv);
            qa *= inv; qx *= inv; qy *= inv; qz *= inv;
            const double aa = qa * qa, xx = qx * qx, yy = qy * qy, zz = qz * qz;
            const double xy = qx * qy, az = qa * qz, zx = qz * qx, ay = qa * qy, yz = qy * qz, ax = qa * qx;
            rot[0] = (float)(aa + xx - yy - zz); rot[1] = (float)(2.0 * (xy - az)); rot[2] = (float)(2.0 * (zx + ay));
            rot[3] = (float)(2.0 * (xy + az)); rot[4] = (float)(aa - xx + yy - zz); rot[5] = (float)(2.0 * (yz - ax));
            rot[6] = (float)(2.0 * (zx - ay)); rot[7] = (float)(2.0 * (yz + ax)); rot[8] = (float)(aa - xx - yy + zz);
        }
    } else if (degenerate) {
        *degenerate = false;
    }
    return msd;
}


// ---------------------------------------------------------------------------------------------
// Throughput variant for the all-pairs epilogue (no rotation): branch-free so that two independent
// solves interleave in one instruction stream.  Same polynomial and root as qcp_solve:
//   * coefficients in float64 from the invariants of M alone -- with ss = |M|_F^2 and cof M the cofactor matrix,
//     C2 = -2 ss, C1 = -8 det M, C0 = det K = ss^2 - 4 |cof M|_F^2 (expand the product of the four roots
//     s1+s2+s3', s1-s2-s3', ...: (s1^2+s2^2+s3^2)^2 - 4 (s1^2 s2^2 + s2^2 s3^2 + s3^2 s1^2)); 43 float64
//     operations instead of 64 for the ten K entries and the Laplace minors, and det M falls out of row 0 of cof M;
//   * scaling by an exact power of two instead of a division;
//   * float32 Newton steps from the upper bound (warp-uniform exit), then two steps with the polynomial and its
//     derivative in float64 and the quotient in float32 (one MUFU reciprocal);
//   * the same largest-root certificate as qcp_newton, returned in trusted[]: pairs that fail it (collinear atoms,
//     two-atom selections) must be redone by the caller with qcp_rmsd_closed -- kept out of this function so that M
//     is dead after the coefficients and the hot path's register footprint does not pay for the rare one;
//   * float32 square root of the float64 msd.
// ---------------------------------------------------------------------------------------------
// single-MUFU reciprocal and square root (1-2 ulp) for the throughput solvers: the IEEE-rounded __frcp_rn / sqrtf
// expand to a MUFU plus a ~15-instruction fix-up with a slow-path branch each, three times per pair.  The host build
// (tests/host_qcp) has no such instruction and takes the exact operations.
__device__ __forceinline__ float rcp_approx(float x)
{
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
__device__ __forceinline__ float sqrt_approx(float x)
{
#ifdef __CUDA_ARCH__
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}

// the slow path of the two solvers below: RMSD of one pair through the closed form
__device__ __forceinline__ float qcp_rmsd_closed(const float* M, float Ga, float Gb, float inv_n)
{
    double m[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = (double)M[i];
    const double msd = 2.0 * (0.5 * ((double)Ga + (double)Gb) - qcp_lambda_closed(m)) * (double)inv_n;
    return sqrtf(fmaxf((float)msd, 0.f));
}

template <int NP>
__device__ __forceinline__ void qcp_msd_fast(const float (&M)[NP][9], const float (&Ga)[NP], const float (&Gb)[NP],
                                             const bool (&active)[NP], float inv_n, float (&rmsd)[NP],
                                             bool (&trusted)[NP])
{
    double c2[NP], c1[NP], c0[NP], e0[NP], scale_back[NP];
    float t[NP], f2[NP], f1[NP], f0[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const double m0 = M[p][0], m1 = M[p][1], m2 = M[p][2], m3 = M[p][3], m4 = M[p][4], m5 = M[p][5], m6 = M[p][6],
                     m7 = M[p][7], m8 = M[p][8];
        const double k00 = m4 * m8 - m5 * m7, k01 = m5 * m6 - m3 * m8, k02 = m3 * m7 - m4 * m6;
        const double k10 = m2 * m7 - m1 * m8, k11 = m0 * m8 - m2 * m6, k12 = m1 * m6 - m0 * m7;
        const double k20 = m1 * m5 - m2 * m4, k21 = m2 * m3 - m0 * m5, k22 = m0 * m4 - m1 * m3;
        const double ss = m0 * m0 + m1 * m1 + m2 * m2 + m3 * m3 + m4 * m4 + m5 * m5 + m6 * m6 + m7 * m7 + m8 * m8;
        const double nc = k00 * k00 + k01 * k01 + k02 * k02 + k10 * k10 + k11 * k11 + k12 * k12 + k20 * k20 + k21 * k21 +
                          k22 * k22;
        const double detM = m0 * k00 + m1 * k01 + m2 * k02;
        const double C0 = ss * ss - 4.0 * nc, C2 = -2.0 * ss, C1 = -8.0 * detM;
        e0[p] = 0.5 * ((double)Ga[p] + (double)Gb[p]);
        // upper bound of the largest root, float32 is enough (nudged up so rounding cannot undershoot)
        float ub = fminf(sqrt_approx(3.0f * (float)ss) * 1.000001f, 0.5f * (Ga[p] + Gb[p]) * 1.000001f);
        ub = fmaxf(ub, 1e-30f);
        // exact power-of-two scale s = 2^-e with ub*s in [1,2)
        const int e = ((__float_as_int(ub) >> 23) & 0xff) - 127;
        const double s1 = __longlong_as_double((long long)(1023 - e) << 52);
        const double s2 = s1 * s1;
        scale_back[p] = __longlong_as_double((long long)(1023 + e) << 52);
        c2[p] = C2 * s2; c1[p] = C1 * s2 * s1; c0[p] = C0 * s2 * s2;
        f2[p] = (float)c2[p]; f1[p] = (float)c1[p]; f0[p] = (float)c0[p];
        t[p] = ub * __int_as_float((127 - e) << 23);
    }
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {  // two steps per trip: half the votes and branches on the dependency chain
        bool conv = true;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float t2 = t[p] * t[p];
                const float b = (t2 + f2[p]) * t[p];
                const float a = b + f1[p];
                const float den = fmaf(2.0f * t2, t[p], b + a);
                const float num = fmaf(a, t[p], f0[p]);
                const float d = (fabsf(den) > 1e-30f) ? __fdividef(num, den) : 0.0f;
                t[p] -= d;
                if (half == 1) conv = conv && (!active[p] || fabsf(d) <= 4e-6f * t[p]);  // idle lanes must not hold the warp back
            }
        }
        if (__all_sync(0xffffffffu, conv)) break;  // warp-uniform exit: no divergence inside the loop
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        double x = (double)t[p];
        bool ok = true;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const double x2 = x * x;
            const double b = (x2 + c2[p]) * x;
            const double a = b + c1[p];
            const double den = 2.0 * x2 * x + b + a;
            const double num = a * x + c0[p];
            // the step in float32: its ~1e-7 relative error perturbs the quadratically converging iterate by
            // 1e-7 |d| <= 1e-12 x, and the division stays off the float64 dependency chain
            const double d = (double)((float)num * rcp_approx((float)den));
            x -= d;
            // largest-root certificate, on the last step only: a vanishing or negative P'(x), a NaN or a step that
            // is not a negligible correction all fail it, and the pair is redone through the closed form
            if (it == 1) ok = den > 0.0 && fabs(d) <= 1e-8 * x && 6.0 * x * x > -c2[p];
        }
        trusted[p] = ok || !active[p];
        const double lam = x * scale_back[p];
        double msd = 2.0 * (e0[p] - lam) * (double)inv_n;
        rmsd[p] = sqrt_approx(fmaxf((float)msd, 0.0f));
    }
}

// error-free transformation a + b = s + e (Knuth), 6 float32 operations, no ordering of |a|, |b| needed
__device__ __forceinline__ void two_sum(float a, float b, float& s, float& e)
{
    s = __fadd_rn(a, b);
    const float bb = __fadd_rn(s, -a);
    e = __fadd_rn(__fadd_rn(a, -__fadd_rn(s, -bb)), __fadd_rn(b, -bb));
}

// ---------------------------------------------------------------------------------------------
// All-pairs epilogue solver, round 2: float32 throughout, and MORE accurate than the float64 polish above on the
// inputs the all-pairs operands produce -- because it never forms the cancelling quantity.
//
// RMSD^2 = 2 (S - lambda)/N with S = (G_a + G_b)/2 and lambda the largest eigenvalue of the 4x4 key matrix K(M).
// Frames reach the epilogue pre-aligned (every frame is rotated onto its nearest reference structure, the references
// onto each other: allpairs_refs.cu), so for the pairs whose RMSD is small against their size -- the ones where
// S - lambda cancels -- the residual rotation is small and lambda = T + delta with T = tr M and delta second order in
// it.  With K' = K - T I = [[0, a^T], [a, B]]  (a = the antisymmetric part of M, B = -2 (T I - sym M) restricted
// suitably: B11 = -2 (Syy + Szz), B12 = Sxy + Syx, ...) the characteristic polynomial in delta is
//
//     delta^4 + 4T delta^3 + (c(B) - |a|^2) delta^2 - (det B + a^T B a + 4T |a|^2) delta - a^T adj(B) a,
//
// whose coefficients are sums of products WITHOUT the catastrophic cancellation of P(lambda) near lambda ~ S, and
//
//     S - lambda = (S - T) - delta:   S - T is evaluated exactly (error-free float32 sums), delta is small and its
//                                     float32 relative error is harmless.
//
// Numpy emulation of exactly this arithmetic: pre-aligned MD-like pairs 3e-8 nm from the
// float64 eigenvalue (the float64-polished route: 1e-6 class, float32 route of the reference: 1e-5), iid pairs 2.5e-7.
// Laguerre's iteration from the upper bound of delta (monotone, as Newton's is for lambda); 2-3 steps on pre-aligned pairs.
//
// trusted[] = false where float32 is not enough: pairs that are similar but NOT in a common orientation (delta large
// against S - lambda: the estimated root error 4e-7 * sum|terms| / P'(delta) exceeds the tolerance), and (nearly) double
// largest roots.  The caller redoes those through qcp_msd_fast / the closed form.
// ---------------------------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ void qcp_msd_shift(const float (&M)[NP][9], const float (&Ga)[NP], const float (&Gb)[NP],
                                              const bool (&active)[NP], float n_atoms, float (&rmsd)[NP],
                                              bool (&trusted)[NP])
{
    float p3[NP], p2[NP], p1[NP], p0[NP], d[NP], e0[NP], s1v[NP], cn0[NP], cn1[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) ss = fmaf(M[p][i], M[p][i], ss);
        float ub = fminf(sqrt_approx(3.0f * ss) * 1.000001f, 0.5f * (Ga[p] + Gb[p]) * 1.000001f);
        ub = fmaxf(ub, 1e-30f);
        const int e = ((__float_as_int(ub) >> 23) & 0xff) - 127;
        const float s1 = __int_as_float((127 - e) << 23);  // exact power of two: ub * s1 in [1, 2)
        s1v[p] = s1;
        const float m0 = M[p][0] * s1, m1 = M[p][1] * s1, m2 = M[p][2] * s1, m3 = M[p][3] * s1, m4 = M[p][4] * s1,
                    m5 = M[p][5] * s1, m6 = M[p][6] * s1, m7 = M[p][7] * s1, m8 = M[p][8] * s1;
        const float T = (m0 + m4) + m8;
        const float a1 = m7 - m5, a2 = m2 - m6, a3 = m3 - m1;
        const float B11 = -2.0f * (m4 + m8), B22 = -2.0f * (m0 + m8), B33 = -2.0f * (m0 + m4);
        const float B12 = m1 + m3, B13 = m2 + m6, B23 = m5 + m7;
        const float A11 = B22 * B33 - B23 * B23, A22 = B11 * B33 - B13 * B13, A33 = B11 * B22 - B12 * B12;
        const float A12 = B13 * B23 - B12 * B33, A13 = B12 * B23 - B13 * B22, A23 = B12 * B13 - B11 * B23;
        const float cB = A11 + A22 + A33;
        const float detB = B11 * A11 + B12 * A12 + B13 * A13;
        const float aa = a1 * a1 + a2 * a2 + a3 * a3;
        const float aBa = a1 * (B11 * a1 + B12 * a2 + B13 * a3) + a2 * (B12 * a1 + B22 * a2 + B23 * a3) +
                          a3 * (B13 * a1 + B23 * a2 + B33 * a3);
        const float aAa = a1 * (A11 * a1 + A12 * a2 + A13 * a3) + a2 * (A12 * a1 + A22 * a2 + A23 * a3) +
                          a3 * (A13 * a1 + A23 * a2 + A33 * a3);
        // rounding noise of the coefficients themselves, which cancel internally for near-singular B (two-atom and
        // collinear selections): |d p0| <~ eps |B|^2 |a|^2, |d p1| <~ eps |B|^3
        const float bb = B11 * B11 + B22 * B22 + B33 * B33 + 2.0f * (B12 * B12 + B13 * B13 + B23 * B23);
        cn0[p] = bb * aa;
        cn1[p] = bb * sqrt_approx(bb);
        p3[p] = 4.0f * T;
        p2[p] = cB - aa;
        p1[p] = -(detB + aBa + p3[p] * aa);
        p0[p] = -aAa;
        // S - T without rounding: (G_a/2 + G_b/2) - ((m0 + m4) + m8), every partial sum carried as (high, low)
        float h1, l1, h2, l2, h3, l3, h4, l4;
        two_sum(0.5f * Ga[p] * s1, 0.5f * Gb[p] * s1, h1, l1);
        two_sum(m0, m4, h2, l2);
        two_sum(h2, m8, h3, l3);
        two_sum(h1, -h3, h4, l4);
        e0[p] = h4 + (((l1 - l2) - l3) + l4);
        d[p] = fmaxf(ub * s1 - T, 0.0f) + 4e-6f;  // upper bound of delta
    }
    // Laguerre's iteration (degree 4) from the upper bound: for a polynomial whose roots are all real -- the key matrix
    // is symmetric -- the iterates fall monotonically onto the largest root like Newton's, but the first step from far
    // away already lands next to it and convergence is cubic:
    //     x <- x - 4 P / (P' + sqrt(9 P'^2 - 24 P P''/2 ... )) = x - 4 P / (P' + 3 sqrt(P'^2 - (8/3) P (P''/2))).
    // float32 emulation on 20 000 pairs (300 atoms): iid frames 2.7 steps on average, 4.1 for the slowest of the 64 pairs
    // a warp carries (Newton: 6.4, and 12 for the slowest: every lane waits for it); pre-aligned MD-like pairs 2.
    // A step of <= 1e-3 relative leaves an error of ~1e-9; a step <= 0 means P <= 0 at the iterate: rounding noise,
    // nothing left to gain (the certificate below evaluates the polynomial once more at the final iterate either way).
    auto laguerre = [&](int p) -> bool {
        const float x = d[p];
        const float b3 = x + p3[p];
        const float b2 = fmaf(b3, x, p2[p]);
        const float b1 = fmaf(b2, x, p1[p]);
        const float val = fmaf(b1, x, p0[p]);
        const float c3 = b3 + x;
        const float c2 = fmaf(c3, x, b2);
        const float den = fmaf(c2, x, b1);           // P'
        const float hc = fmaf(c3 + x, x, c2);        // P'' / 2
        const float disc = fmaxf(fmaf(-2.6666667f * val, hc, den * den), 0.0f);
        const float dd = fmaf(3.0f, sqrt_approx(disc), den);
        const float step = (fabsf(dd) > 1e-30f) ? 4.0f * val * rcp_approx(dd) : 0.0f;
        d[p] = x - step;
        return step <= fmaf(1e-3f, fabsf(d[p]), 1e-9f);
    };
#pragma unroll
    for (int p = 0; p < NP; ++p) laguerre(p);
#pragma unroll 1
    for (int it = 0; it < 12; ++it) {
        bool conv = true;
#pragma unroll
        for (int p = 0; p < NP; ++p) conv = (laguerre(p) || !active[p]) && conv;  // idle lanes must not hold the warp back
        if (__all_sync(0xffffffffu, conv)) break;  // warp-uniform exit: no divergence inside the loop
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        // one more evaluation at the last iterate: it feeds the certificate below, and its Newton correction is applied
        // for free (the loop stops at a step of 1e-3 relative; where the largest roots are close -- dissimilar frames --
        // the cubic tail of that step was still up to ~1e-6 of delta, 6e-6 nm on a 2.4 nm RMSD)
        const float x = d[p], ax = fabsf(x);
        const float b3 = x + p3[p];
        const float b2 = fmaf(b3, x, p2[p]);
        const float b1 = fmaf(b2, x, p1[p]);
        const float val = fmaf(b1, x, p0[p]);
        const float den = fmaf(fmaf(b3 + x, x, b2), x, b1);
        const float xp = den > 1e-30f ? x - val * rcp_approx(den) : x;
        const float es = fmaxf(e0[p] - xp, 0.0f);                    // (S - lambda), scaled
        const float msd = 2.0f * es / (s1v[p] * n_atoms);
        rmsd[p] = sqrt_approx(msd);
        const float mag = fmaf(fmaf(fmaf(ax + fabsf(p3[p]), ax, fabsf(p2[p])), ax, fabsf(p1[p])), ax, fabsf(p0[p]));
        // float32 noise of P at the iterate: evaluation (4e-7 * sum |terms|) plus the coefficients' own rounding
        const float noise = fmaf(4e-7f, mag, 2e-7f * fmaf(cn1[p], ax, cn0[p]));
        // acceptable error of the RMSD: 4e-6 nm + 5e-5 relative on the ESTIMATE (the estimate is ~4x conservative; the
        // parity tolerance is 1e-5 nm or 1e-4 relative); d(rmsd) = d(delta) / (N s1 rmsd); plus the 3e-7 noise floor of M
        const float tol = fmaf(n_atoms * s1v[p] * rmsd[p], fmaf(5e-5f, rmsd[p], 4e-6f), 3e-7f);
        const float half_curv = fmaf(6.0f * x, x + 0.5f * p3[p], p2[p]);   // P''(x) / 2
        // converged: the next Newton step would be negligible, or P is at its noise floor; P' > 0 and P'' > 0: right of
        // every other root; the root-error estimate noise / P' within tolerance; and the root simple at float32
        // resolution -- next to a (nearly) double root the error is sqrt(2 noise / P'') instead, and the linear estimate
        // only holds while P'^2 >> noise * P''
        const bool ok = den > 0.0f && fabsf(val) <= fmaxf(den * fmaf(1e-6f, ax, 4e-9f), 2.0f * noise) && half_curv > 0.0f &&
                        noise <= den * tol && den * den >= 64.0f * noise * half_curv;
        trusted[p] = ok || !active[p];
    }
}

// All-float32 variant (development / comparison): coefficients, Newton and the final cancellation in float32,
// i.e. the precision class of the reference's own msdFromMandG (theobald_rmsd.cpp:217-277).  ok[] as in qcp_msd_fast.
template <int NP>
__device__ __forceinline__ void qcp_msd_f32(const float (&M)[NP][9], const float (&Ga)[NP], const float (&Gb)[NP],
                                            const bool (&active)[NP], float inv_n, float (&rmsd)[NP],
                                            bool (&ok)[NP])
{
    float c2[NP], c1[NP], c0[NP], e0[NP], t[NP], sc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        e0[p] = 0.5f * (Ga[p] + Gb[p]);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) ss = fmaf(M[p][i], M[p][i], ss);
        float ub = fminf(sqrt_approx(3.0f * ss) * 1.000001f, e0[p] * 1.000001f);
        ub = fmaxf(ub, 1e-30f);
        const int e = ((__float_as_int(ub) >> 23) & 0xff) - 127;
        const float s1 = __int_as_float((127 - e) << 23);
        sc[p] = __int_as_float((127 + e) << 23);
        float m[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) m[i] = M[p][i] * s1;  // exact scaling: work with M/2^e
        const float k00 = m[4] * m[8] - m[5] * m[7], k01 = m[5] * m[6] - m[3] * m[8], k02 = m[3] * m[7] - m[4] * m[6];
        const float k10 = m[2] * m[7] - m[1] * m[8], k11 = m[0] * m[8] - m[2] * m[6], k12 = m[1] * m[6] - m[0] * m[7];
        const float k20 = m[1] * m[5] - m[2] * m[4], k21 = m[2] * m[3] - m[0] * m[5], k22 = m[0] * m[4] - m[1] * m[3];
        const float nc = k00 * k00 + k01 * k01 + k02 * k02 + k10 * k10 + k11 * k11 + k12 * k12 + k20 * k20 + k21 * k21 +
                         k22 * k22;
        const float sss = ss * s1 * s1;
        c0[p] = sss * sss - 4.0f * nc;
        c2[p] = -2.0f * sss;
        c1[p] = -8.0f * (m[0] * k00 + m[1] * k01 + m[2] * k02);
        t[p] = ub * s1;
        ok[p] = false;
    }
#pragma unroll 1
    for (int it = 0; it < 20; ++it) {
        bool conv = true;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float t2 = t[p] * t[p];
            const float b = (t2 + c2[p]) * t[p];
            const float a = b + c1[p];
            const float den = fmaf(2.0f * t2, t[p], b + a);
            const float num = fmaf(a, t[p], c0[p]);
            const float d = (fabsf(den) > 1e-30f) ? __fdividef(num, den) : 0.0f;
            t[p] -= d;
            // float32 noise in P is ~1e-6 t^4: trust the root only where P' is not small against t^3
            ok[p] = den > 0.05f * t2 * t[p] && fabsf(d) <= 1e-6f * t[p] && 6.0f * t[p] * t[p] > -c2[p];
            conv = conv && (!active[p] || fabsf(d) <= 2e-7f * t[p]);
        }
        if (__all_sync(0xffffffffu, conv)) break;
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const float msd = 2.0f * (e0[p] - t[p] * sc[p]) * inv_n;
        rmsd[p] = sqrt_approx(fmaxf(msd, 0.f));
        ok[p] = ok[p] || !active[p];
    }
}

}  // namespace b200
