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
    int n_atoms;
};

// Largest root of  t^4 + C2 t^2 + C1 t + C0  (all four roots are real: K is symmetric).
//
// Start: lam0 = min((G_a+G_b)/2, sqrt(3)*||M||_F).  Both are upper bounds of lambda_max (the first is the
// reference's own starting value, theobald_rmsd.cpp:245; the second follows from sum(lambda_i) = 0 and
// lambda_max = s1 + s2 +- s3 <= sqrt(3)*||M||_F for the singular values s of M), and lambda_max >= s1 >=
// ||M||_F/sqrt(3), so the start is never more than 3x above the root -- for dissimilar (e.g. iid random)
// frames (G_a+G_b)/2 alone can be 10-30x above it and Newton would crawl down by 3/4 per step.  Newton from the
// right of the largest root of a real-rooted polynomial is monotone, so the iteration cannot leave the basin.
// The polynomial is scaled by an exact power of two (t = lambda * 2^-e in [1,2) at the start: no division, no
// rounding); the first iterations run in float32 (4-cycle FMA + MUFU reciprocal), the last ones in float64 with
// a float32-seeded, once-refined reciprocal instead of a DDIV chain.
__device__ __forceinline__ double qcp_newton(double C2, double C1, double C0, double lam_upper, double frob2)
{
    float ub = fminf(sqrtf(3.0f * (float)frob2) * 1.000001f, (float)lam_upper * 1.000001f);
    if (!(ub > 1e-30f)) return 0.0;
    const int e = ((__float_as_int(ub) >> 23) & 0xff) - 127;
    const double s1 = __longlong_as_double((long long)(1023 - e) << 52);
    const double s2 = s1 * s1;
    const double c2 = C2 * s2, c1 = C1 * s2 * s1, c0 = C0 * s2 * s2;
    float t = ub * (float)s1;
    {
        const float f2 = (float)c2, f1 = (float)c1, f0 = (float)c0;
#pragma unroll 1
        for (int it = 0; it < 24; ++it) {
            const float t2 = t * t;
            const float b = (t2 + f2) * t;
            const float a = b + f1;
            const float den = fmaf(2.0f * t2, t, b + a);
            if (!(fabsf(den) > 1e-30f)) break;
            const float delta = __fdividef(fmaf(a, t, f0), den);
            t -= delta;
            if (!(fabsf(delta) > 4e-6f * fabsf(t))) break;
        }
        if (!(t > 0.0f) || !(t <= 2.0f)) t = ub * (float)s1;  // numerical accident: restart the float64 phase from the bound
    }
    double x = (double)t;
#pragma unroll 1
    for (int it = 0; it < 40; ++it) {
        const double x2 = x * x;
        const double b = (x2 + c2) * x;
        const double a = b + c1;
        const double den = 2.0 * x2 * x + b + a;
        if (!(fabs(den) > 1e-300)) break;
        // a float32 reciprocal is enough: a 1e-7 relative error in the step only perturbs the quadratically
        // converging iterate by 1e-7 * |delta|
        const double delta = (a * x + c0) * (double)__frcp_rn((float)den);
        if (!(delta == delta)) break;
        x -= delta;
        if (fabs(delta) <= 1e-10 * fabs(x)) break;  // quadratic: the step just taken leaves an error ~delta^2
    }
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

    const double lam = qcp_newton(C2, C1, C0, 0.5 * (in.Ga + in.Gb), ss);
    double msd = (in.Ga + in.Gb - 2.0 * lam) / in.n_atoms;
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
//   * coefficients C2, C1, C0 in float64 (the Laplace minors have plenty of ILP);
//   * scaling by an exact power of two instead of a division;
//   * a fixed number of float32 Newton steps from the upper bound, then two float64 steps whose
//     divisions are a float32 reciprocal refined by one Newton step (3 DFMA instead of a DDIV chain);
//   * float32 square root of the float64 msd.
// ---------------------------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ void qcp_msd_fast(const float (&M)[NP][9], const float (&Ga)[NP], const float (&Gb)[NP],
                                             const bool (&active)[NP], float inv_n, float (&rmsd)[NP])
{
    double c2[NP], c1[NP], c0[NP], e0[NP], scale_back[NP];
    float t[NP], f2[NP], f1[NP], f0[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const double Sxx = M[p][0], Sxy = M[p][1], Sxz = M[p][2];
        const double Syx = M[p][3], Syy = M[p][4], Syz = M[p][5];
        const double Szx = M[p][6], Szy = M[p][7], Szz = M[p][8];
        const double k00 = Sxx + Syy + Szz, k11 = Sxx - Syy - Szz, k22 = -Sxx + Syy - Szz, k33 = -Sxx - Syy + Szz;
        const double k01 = Szy - Syz, k02 = Sxz - Szx, k03 = Syx - Sxy;
        const double k12 = Syx + Sxy, k13 = Sxz + Szx, k23 = Szy + Syz;
        const double ss = Sxx * Sxx + Sxy * Sxy + Sxz * Sxz + Syx * Syx + Syy * Syy + Syz * Syz + Szx * Szx + Szy * Szy +
                          Szz * Szz;
        const double detM = Sxx * (Syy * Szz - Syz * Szy) + Syx * (Szy * Sxz - Szz * Sxy) + Szx * (Sxy * Syz - Sxz * Syy);
        const double a01 = k00 * k11 - k01 * k01, a02 = k00 * k12 - k02 * k01, a03 = k00 * k13 - k03 * k01;
        const double a12 = k01 * k12 - k02 * k11, a13 = k01 * k13 - k03 * k11, a23 = k02 * k13 - k03 * k12;
        const double b01 = k02 * k13 - k12 * k03, b02 = k02 * k23 - k22 * k03, b03 = k02 * k33 - k23 * k03;
        const double b12 = k12 * k23 - k22 * k13, b13 = k12 * k33 - k23 * k13, b23 = k22 * k33 - k23 * k23;
        const double C0 = a01 * b23 - a02 * b13 + a03 * b12 + a12 * b03 - a13 * b02 + a23 * b01;
        const double C2 = -2.0 * ss, C1 = -8.0 * detM;
        e0[p] = 0.5 * ((double)Ga[p] + (double)Gb[p]);
        // upper bound of the largest root, float32 is enough (nudged up so rounding cannot undershoot)
        float ub = fminf(sqrtf(3.0f * (float)ss) * 1.000001f, (float)e0[p] * 1.000001f);
        ub = fmaxf(ub, 1e-30f);
        // exact power-of-two scale s = 2^-e with ub*s in [1,2)
        const int e = ((__float_as_int(ub) >> 23) & 0xff) - 127;
        const double s1 = __longlong_as_double((long long)(1023 - e) << 52);
        const double s2 = s1 * s1;
        scale_back[p] = __longlong_as_double((long long)(1023 + e) << 52);
        c2[p] = C2 * s2; c1[p] = C1 * s2 * s1; c0[p] = C0 * s2 * s2;
        f2[p] = (float)c2[p]; f1[p] = (float)c1[p]; f0[p] = (float)c0[p];
        t[p] = ub * (float)s1;
    }
#pragma unroll 1
    for (int it = 0; it < 16; ++it) {
        bool conv = true;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float t2 = t[p] * t[p];
            const float b = (t2 + f2[p]) * t[p];
            const float a = b + f1[p];
            const float den = fmaf(2.0f * t2, t[p], b + a);
            const float num = fmaf(a, t[p], f0[p]);
            const float d = (fabsf(den) > 1e-30f) ? __fdividef(num, den) : 0.0f;
            t[p] -= d;
            conv = conv && (!active[p] || fabsf(d) <= 4e-6f * t[p]);  // idle lanes must not hold the warp back
        }
        if (__all_sync(0xffffffffu, conv)) break;  // warp-uniform exit: no divergence inside the loop
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        double x = (double)t[p];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const double x2 = x * x;
            const double b = (x2 + c2[p]) * x;
            const double a = b + c1[p];
            const double den = 2.0 * x2 * x + b + a;
            const double num = a * x + c0[p];
            double r = (double)__frcp_rn((float)den);
            r = r * (2.0 - den * r);
            const double d = num * r;
            x -= (fabs(den) > 1e-300 && d == d) ? d : 0.0;
        }
        const double lam = x * scale_back[p];
        double msd = 2.0 * (e0[p] - lam) * (double)inv_n;
        msd = msd > 0.0 ? msd : 0.0;
        rmsd[p] = sqrtf((float)msd);
    }
}

// All-float32 variant (development / comparison): coefficients, Newton and the final cancellation in float32,
// i.e. the precision class of the reference's own msdFromMandG (theobald_rmsd.cpp:217-277).
template <int NP>
__device__ __forceinline__ void qcp_msd_f32(const float (&M)[NP][9], const float (&Ga)[NP], const float (&Gb)[NP],
                                            const bool (&active)[NP], float inv_n, float (&rmsd)[NP])
{
    float c2[NP], c1[NP], c0[NP], e0[NP], t[NP], sc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        e0[p] = 0.5f * (Ga[p] + Gb[p]);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) ss = fmaf(M[p][i], M[p][i], ss);
        float ub = fminf(sqrtf(3.0f * ss) * 1.000001f, e0[p] * 1.000001f);
        ub = fmaxf(ub, 1e-30f);
        const int e = ((__float_as_int(ub) >> 23) & 0xff) - 127;
        const float s1 = __int_as_float((127 - e) << 23);
        sc[p] = __int_as_float((127 + e) << 23);
        float m[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) m[i] = M[p][i] * s1;  // exact scaling: work with M/2^e
        const float Sxx = m[0], Sxy = m[1], Sxz = m[2], Syx = m[3], Syy = m[4], Syz = m[5], Szx = m[6], Szy = m[7], Szz = m[8];
        const float k00 = Sxx + Syy + Szz, k11 = Sxx - Syy - Szz, k22 = -Sxx + Syy - Szz, k33 = -Sxx - Syy + Szz;
        const float k01 = Szy - Syz, k02 = Sxz - Szx, k03 = Syx - Sxy, k12 = Syx + Sxy, k13 = Sxz + Szx, k23 = Szy + Syz;
        const float detM = Sxx * (Syy * Szz - Syz * Szy) + Syx * (Szy * Sxz - Szz * Sxy) + Szx * (Sxy * Syz - Sxz * Syy);
        const float a01 = k00 * k11 - k01 * k01, a02 = k00 * k12 - k02 * k01, a03 = k00 * k13 - k03 * k01;
        const float a12 = k01 * k12 - k02 * k11, a13 = k01 * k13 - k03 * k11, a23 = k02 * k13 - k03 * k12;
        const float b01 = k02 * k13 - k12 * k03, b02 = k02 * k23 - k22 * k03, b03 = k02 * k33 - k23 * k03;
        const float b12 = k12 * k23 - k22 * k13, b13 = k12 * k33 - k23 * k13, b23 = k22 * k33 - k23 * k23;
        c0[p] = a01 * b23 - a02 * b13 + a03 * b12 + a12 * b03 - a13 * b02 + a23 * b01;
        c2[p] = -2.0f * ss * s1 * s1;
        c1[p] = -8.0f * detM;
        t[p] = ub * s1;
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
            conv = conv && (!active[p] || fabsf(d) <= 2e-7f * t[p]);
        }
        if (__all_sync(0xffffffffu, conv)) break;
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const float msd = 2.0f * (e0[p] - t[p] * sc[p]) * inv_n;
        rmsd[p] = sqrtf(fmaxf(msd, 0.f));
    }
}

}  // namespace b200
