/*
 * oracle.c -- CPU restatement of the mdtraj RMSD hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  The shipped path
 * (mdtraj_b200/) never imports it and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against (a) oracle/_ref/libmdtraj_rmsd_ref.so, which is the reference's own
 * theobald_rmsd.cpp / center.cpp / rotation.cpp compiled where they lie under
 * /root/reference (see oracle/Makefile), and (b) the committed golden vectors
 * under tests/golden/ that were produced by importing the reference's real
 * md.rmsd / Trajectory.superpose (tests/golden/make_golden.py), including the
 * two known answers printed in the reference's notebooks
 * (examples/clustering.ipynb:73 -> 0.188493 nm, examples/centroids.ipynb:111
 * -> index 83).
 *
 * Each function names the reference file:line whose behaviour it restates.
 * The arithmetic follows the reference's *generic* (non-SIMD) variants: single
 * precision where the reference is single, double where it is double.  The
 * SSE build of the reference sums the 3x3 inner products in four SIMD lanes,
 * so agreement with _ref is to float32 rounding noise, not bit-for-bit; the
 * tests state the tolerance.
 *
 * Plain C99, no dependencies beyond libm.  Serial on purpose.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------
 * Largest real root of  t^4 + c2 t^2 + c1 t + c0 = 0  in double.
 * Restates the closed-form route the reference takes:
 *   DirectSolve                    mdtraj/rmsd/src/theobald_rmsd.cpp:183-193
 *   quartic_equation_solve_exact   mdtraj/rmsd/src/theobald_rmsd.cpp:81-137
 *   solve_cubic_equation           mdtraj/rmsd/src/theobald_rmsd.cpp:139-179
 * (Ferrari's method through the resolvent cubic; the cubic term of the
 * quartic is identically zero for the QCP polynomial.)
 * ---------------------------------------------------------------------- */
static double signed_cbrt(double v) { return v >= 0.0 ? pow(v, 1.0 / 3.0) : -pow(-v, 1.0 / 3.0); }

/* roots of u^3 + p2 u^2 + p1 u + p0; returns how many are distinct-real (1 or 3) */
static int monic_cubic_roots(double p2, double p1, double p0, double root[3])
{
    const double shift = p2 / 3.0;
    const double q = p1 / 3.0 - p2 * p2 / 9.0;
    const double r = (p1 * p2 - 3.0 * p0) / 6.0 - p2 * p2 * p2 / 27.0;
    const double disc = q * q * q + r * r;
    if (disc > 0.0) {
        const double sq = sqrt(disc);
        const double s = signed_cbrt(r + sq) + signed_cbrt(r - sq);
        root[0] = s - shift;
        root[1] = root[2] = -0.5 * s - shift;
        return 1;
    }
    if (disc < 0.0) {
        const double th = acos(r / sqrt(-q * q * q)) / 3.0;
        const double m = sqrt(-q);
        const double c = cos(th), s = sin(th);
        root[0] = 2.0 * m * c - shift;
        root[1] = -m * c - shift - sqrt(3.0) * m * s;
        root[2] = -m * c - shift + sqrt(3.0) * m * s;
        return 3;
    }
    {
        const double s = signed_cbrt(r);
        root[0] = 2.0 * s - shift;
        root[1] = root[2] = -s - shift;
        return 3;
    }
}

ORACLE_API double oracle_qcp_largest_root(double c0, double c1, double c2)
{
    double u[3];
    /* resolvent cubic of the depressed quartic (a3 == 0) */
    const int nreal = monic_cubic_roots(-c2, -4.0 * c0, 4.0 * c0 * c2 - c1 * c1, u);
    const double u1 = (nreal == 1) ? u[0] : (u[0] > u[2] ? u[0] : u[2]);

    const double R2 = u1 - c2;
    const double R = R2 > 0.0 ? sqrt(R2) : 0.0;
    double common, split;
    if (R != 0.0) {
        common = -R2 - 2.0 * c2;
        split = 0.25 * (-8.0 * c1) / R;
    } else {
        common = -2.0 * c2;
        split = 2.0 * sqrt(u1 * u1 - 4.0 * c0);
    }
    const double D2 = common + split, E2 = common - split;
    /* the two "upper" members of each root pair; the other two are never larger */
    const double hiD = 0.5 * R + (D2 >= 0.0 ? 0.5 * sqrt(D2) : 0.0);
    const double loD = 0.5 * R - (D2 >= 0.0 ? 0.5 * sqrt(D2) : 0.0);
    const double hiE = -0.5 * R + (E2 >= 0.0 ? 0.5 * sqrt(E2) : 0.0);
    const double loE = -0.5 * R - (E2 >= 0.0 ? 0.5 * sqrt(E2) : 0.0);
    double best = hiD > loD ? hiD : loD;
    if (hiE > best) best = hiE;
    if (loE > best) best = loE;
    return best;
}

/* ------------------------------------------------------------------------
 * msd (and optionally the rotation) from the 3x3 inner-product matrix and the
 * two traces.  Restates msdFromMandG, mdtraj/rmsd/src/theobald_rmsd.cpp:217-334.
 * M is row-major with M[3*i+j] = sum_k a_k[i] * b_k[j].
 * Returns the clamped float32 msd.  *degenerate is set when the reference
 * would have printed "UNCONVERGED ROTATION MATRIX" and returned identity.
 * ---------------------------------------------------------------------- */
ORACLE_API float oracle_msd_from_M_and_G(const float M[9], float Ga, float Gb, int n_atoms,
                                         int want_rot, float rot[9], int *degenerate)
{
    const float Sxx = M[0], Sxy = M[1], Sxz = M[2];
    const float Syx = M[3], Syy = M[4], Syz = M[5];
    const float Szx = M[6], Szy = M[7], Szz = M[8];

    /* symmetric 4x4 key matrix, upper triangle (theobald_rmsd.cpp:227-236;
     * the reference indexes M[i + 3*j], i.e. element (row j, col i)) */
    float k00 = Sxx + Syy + Szz;
    const float k01 = Szy - Syz;
    const float k02 = Sxz - Szx;
    const float k03 = Syx - Sxy;
    float k11 = Sxx - Syy - Szz;
    const float k12 = Syx + Sxy;
    const float k13 = Sxz + Szx;
    float k22 = -Sxx + Syy - Szz;
    const float k23 = Szy + Syz;
    float k33 = -Sxx - Syy + Szz;

    /* characteristic polynomial coefficients, all float32 (:249-271) */
    float sumsq = 0.0f;
    for (int i = 0; i < 9; ++i) sumsq += M[i] * M[i];
    const float C2 = -2.0f * sumsq;

    const float detM = M[0] * (M[4] * M[8] - M[5] * M[7]) + M[3] * (M[7] * M[2] - M[8] * M[1]) +
                       M[6] * (M[1] * M[5] - M[2] * M[4]);
    const float C1 = -8.0f * detM;

    const float C0 = k01 * k01 * k23 * k23 - k22 * k33 * k01 * k01 + 2 * k33 * k01 * k02 * k12 -
                     2 * k01 * k02 * k13 * k23 - 2 * k01 * k03 * k12 * k23 + 2 * k22 * k01 * k03 * k13 +
                     k02 * k02 * k13 * k13 - k11 * k33 * k02 * k02 - 2 * k02 * k03 * k12 * k13 +
                     2 * k11 * k02 * k03 * k23 + k03 * k03 * k12 * k12 - k11 * k22 * k03 * k03 -
                     k00 * k33 * k12 * k12 + 2 * k00 * k12 * k13 * k23 - k00 * k22 * k13 * k13 -
                     k00 * k11 * k23 * k23 + k00 * k11 * k22 * k33;

    /* closed-form root in double, handed back as float32 (:273, :183-193) */
    const float lam = (float)oracle_qcp_largest_root((double)C0, (double)C1, (double)C2);

    /* float32 cancellation, clamp at zero (:275-277) */
    float msd = (Ga + Gb - 2.0f * lam) / n_atoms;
    if (!(msd > 0.0f)) msd = 0.0f;

    if (degenerate) *degenerate = 0;
    if (want_rot) {
        /* quaternion = cofactors of row 0 of (K - lam I)  (:281-297) */
        k00 -= lam; k11 -= lam; k22 -= lam; k33 -= lam;
        const float m2233 = k22 * k33 - k23 * k23;
        const float m1233 = k12 * k33 - k13 * k23;
        const float m1223 = k12 * k23 - k13 * k22;
        const float m0223 = k02 * k23 - k03 * k22;
        const float m0233 = k02 * k33 - k03 * k23;
        const float m0213 = k02 * k13 - k03 * k12;
        float qa = k11 * m2233 - k12 * m1233 + k13 * m1223;
        float qx = -k01 * m2233 + k12 * m0233 - k13 * m0223;
        float qy = k01 * m1233 - k11 * m0233 + k13 * m0213;
        float qz = -k01 * m1223 + k11 * m0223 - k12 * m0213;
        const float nrm2 = qa * qa + qx * qx + qy * qy + qz * qz;
        if (nrm2 < 1e-11f) {
            /* identity fallback (:299-302) */
            for (int i = 0; i < 9; ++i) rot[i] = (i % 4 == 0) ? 1.0f : 0.0f;
            if (degenerate) *degenerate = 1;
        } else {
            const float nrm = sqrtf(nrm2); /* (float)sqrt((double)float) == sqrtf(float) */
            qa /= nrm; qx /= nrm; qy /= nrm; qz /= nrm;
            const float aa = qa * qa, xx = qx * qx, yy = qy * qy, zz = qz * qz;
            const float xy = qx * qy, az = qa * qz, zx = qz * qx, ay = qa * qy, yz = qy * qz, ax = qa * qx;
            rot[0] = aa + xx - yy - zz; rot[1] = 2 * (xy - az);     rot[2] = 2 * (zx + ay);
            rot[3] = 2 * (xy + az);     rot[4] = aa - xx + yy - zz; rot[5] = 2 * (yz - ax);
            rot[6] = 2 * (zx - ay);     rot[7] = 2 * (yz + ax);     rot[8] = aa - xx - yy + zz;
        }
    }
    return msd;
}

/* ------------------------------------------------------------------------
 * Centre every frame in place and emit its trace.
 * Restates inplace_center_and_trace_atom_major,
 *   mdtraj/rmsd/src/center_generic.h:3-44 (declared mdtraj/rmsd/include/center.h:7):
 * double sums -> float32 mean -> float32 subtraction -> double trace -> float32.
 * traces may be NULL.
 * ---------------------------------------------------------------------- */
ORACLE_API void oracle_center_and_trace(float *xyz, float *traces, int64_t n_frames, int n_atoms)
{
    for (int64_t f = 0; f < n_frames; ++f) {
        float *p = xyz + (size_t)f * n_atoms * 3;
        double s[3] = {0, 0, 0};
        for (int a = 0; a < n_atoms; ++a)
            for (int c = 0; c < 3; ++c) s[c] += p[3 * a + c];
        const float mu[3] = {(float)(s[0] / n_atoms), (float)(s[1] / n_atoms), (float)(s[2] / n_atoms)};
        double tr = 0.0;
        for (int a = 0; a < n_atoms; ++a) {
            float sq = 0.0f;
            for (int c = 0; c < 3; ++c) {
                p[3 * a + c] -= mu[c];
                /* reference adds the three float32 squares (float32 sum) into the double */
            }
            sq = p[3 * a] * p[3 * a] + p[3 * a + 1] * p[3 * a + 1] + p[3 * a + 2] * p[3 * a + 2];
            tr += sq;
        }
        if (traces) traces[f] = (float)tr;
    }
}

/* ------------------------------------------------------------------------
 * msd between two centred frames.  Restates msd_atom_major,
 *   mdtraj/rmsd/src/theobald_rmsd_generic.h:63-119 (SSE: theobald_rmsd_sse.h:184-335),
 * including the same-pointer shortcut (generic :75-81, SSE :256-262).
 * ---------------------------------------------------------------------- */
ORACLE_API float oracle_msd_atom_major(int n_atoms, const float *a, const float *b, float Ga, float Gb,
                                       int want_rot, float rot[9])
{
    if (a == b && Ga == Gb) {
        if (want_rot)
            for (int i = 0; i < 9; ++i) rot[i] = (i % 4 == 0) ? 1.0f : 0.0f;
        return 0.0f;
    }
    float M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < n_atoms; ++k)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) M[3 * i + j] += b[3 * k + j] * a[3 * k + i];
    return oracle_msd_from_M_and_G(M, Ga, Gb, n_atoms, want_rot, rot, NULL);
}

/* axis-major twin: mdtraj/rmsd/src/theobald_rmsd_generic.h:7-60 */
ORACLE_API float oracle_msd_axis_major(int n_atoms, int rowstride, const float *aT, const float *bT, float Ga,
                                       float Gb)
{
    if (aT == bT && Ga == Gb) return 0.0f;
    float M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < n_atoms; ++k)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) M[3 * i + j] += aT[i * rowstride + k] * bT[j * rowstride + k];
    return oracle_msd_from_M_and_G(M, Ga, Gb, n_atoms, 0, NULL, NULL);
}

/* a <- a . R (row vector times row-major R).  rot_atom_major,
 *   mdtraj/rmsd/src/rotation_generic.h:29-44 */
ORACLE_API void oracle_rot_atom_major(int n_atoms, float *a, const float rot[9])
{
    for (int k = 0; k < n_atoms; ++k) {
        const float x = a[3 * k], y = a[3 * k + 1], z = a[3 * k + 2];
        a[3 * k + 0] = x * rot[0] + y * rot[3] + z * rot[6];
        a[3 * k + 1] = x * rot[1] + y * rot[4] + z * rot[7];
        a[3 * k + 2] = x * rot[2] + y * rot[5] + z * rot[8];
    }
}

/* msd between a.R and b, double accumulation.  rot_msd_atom_major,
 *   mdtraj/rmsd/src/rotation_generic.h:47-102 */
ORACLE_API float oracle_rot_msd_atom_major(int n_atoms, const float *a, const float *b, const float rot[9])
{
    double acc = 0.0;
    for (int k = 0; k < n_atoms; ++k) {
        const float x = a[3 * k], y = a[3 * k + 1], z = a[3 * k + 2];
        const float dx = b[3 * k + 0] - (x * rot[0] + y * rot[3] + z * rot[6]);
        const float dy = b[3 * k + 1] - (x * rot[1] + y * rot[4] + z * rot[7]);
        const float dz = b[3 * k + 2] - (x * rot[2] + y * rot[5] + z * rot[8]);
        acc += dx * dx + dy * dy + dz * dz;
    }
    return (float)(acc / (double)n_atoms);
}

/* plain msd without alignment, float32 accumulation.  msd_nosuperpose,
 *   mdtraj/rmsd/_rmsd.pyx:765-793 */
ORACLE_API float oracle_msd_nosuperpose(int n_atoms, const float *cur, const float *ref)
{
    float acc = 0.0f;
    for (int k = 0; k < 3 * n_atoms; ++k) {
        const float d = cur[k] - ref[k];
        acc += d * d;
    }
    return acc / n_atoms;
}

/* ------------------------------------------------------------------------
 * The per-frame driver loops of the Cython boundary, serial.
 * ---------------------------------------------------------------------- */

/* one-vs-many on already centred data with traces: the loop of
 *   mdtraj/rmsd/_rmsd.pyx:217-224 (== getMultipleRMSDs_atom_major :562-615
 * with the argument roles swapped: there a = xyz1[frame], b = xyz2[i]). */
ORACLE_API void oracle_rmsd_one_vs_many(const float *target, const float *target_g, int64_t n_frames, int n_atoms,
                                        const float *ref_frame, float ref_g, float *out)
{
    for (int64_t i = 0; i < n_frames; ++i)
        out[i] = sqrtf(oracle_msd_atom_major(n_atoms, target + (size_t)i * n_atoms * 3, ref_frame, target_g[i],
                                             ref_g, 0, NULL));
}

/* the superpose=False loop, mdtraj/rmsd/_rmsd.pyx:234-241 */
ORACLE_API void oracle_rmsd_nosuperpose(const float *target, int64_t n_frames, int n_atoms, const float *ref_frame,
                                        float *out)
{
    for (int64_t i = 0; i < n_frames; ++i)
        out[i] = sqrtf(oracle_msd_nosuperpose(n_atoms, target + (size_t)i * n_atoms * 3, ref_frame));
}

/* superpose_atom_major, mdtraj/rmsd/_rmsd.pyx:620-674: rotation from the align
 * subset (a = mobile frame, b = target frame), applied to all displaced atoms.
 * rot_out (n_frames x 9) may be NULL. */
ORACLE_API void oracle_superpose_atom_major(const float *align_target_frame, float g_target, const float *align_mobile,
                                            const float *g_mobile, int64_t n_frames, int n_align, float *displace,
                                            int n_displace, float *rot_out)
{
    for (int64_t i = 0; i < n_frames; ++i) {
        float R[9];
        oracle_msd_atom_major(n_align, align_mobile + (size_t)i * n_align * 3, align_target_frame, g_target,
                              g_mobile[i], 1, R);
        oracle_rot_atom_major(n_displace, displace + (size_t)i * n_displace * 3, R);
        if (rot_out)
            for (int c = 0; c < 9; ++c) rot_out[9 * i + c] = R[c];
    }
}
