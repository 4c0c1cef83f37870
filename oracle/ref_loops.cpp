/*
 * ref_loops.cpp -- OpenMP frame loops around the UNMODIFIED reference kernels.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/oracle.c header).
 *
 * The reference's C entry points (center.h:7, theobald_rmsd.h:13-18,
 * rotation.h:7-12) are per-frame; the frame loop that makes them a hot path
 * lives in Cython (mdtraj/rmsd/_rmsd.pyx, `prange(..., nogil=True)`), which
 * cannot be linked without CPython.  This file restates only those loops, as
 * `#pragma omp parallel for` with the same static schedule Cython's prange
 * emits, and is linked with the reference's own theobald_rmsd.cpp, center.cpp
 * and rotation.cpp compiled from /root/reference (oracle/Makefile).  The
 * arithmetic therefore is the reference's own SSE code; only the loop shell is
 * ours.  Output: oracle/_ref/libmdtraj_rmsd_ref.so (git-ignored).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {
/* prototypes as declared by the reference headers */
void inplace_center_and_trace_atom_major(float *coords, float *traces, const int n_frames, const int n_atoms);
float msd_atom_major(const int nrealatoms, const int npaddedatoms, const float *a, const float *b, const float G_a,
                     const float G_b, int computeRot, float rot[9]);
void rot_atom_major(const int n_atoms, float *a, const float rot[9]);
float rot_msd_atom_major(const int n_real_atoms, const int n_padded_atoms, const float *a, const float *b,
                         const float rot[9]);

int refloops_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void refloops_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

/* md.rmsd, superpose=True branch: mdtraj/rmsd/_rmsd.pyx:203-224.
 * target (F,N,3) and ref_frame (N,3) are centred IN PLACE unless
 * use_traces != 0, exactly as the view path of the reference does. */
void refloops_rmsd(float *target, int64_t n_frames, int n_atoms, float *ref_frame, int use_traces,
                   const float *target_traces, float ref_trace, int parallel, float *out)
{
    float *tg = NULL;
    float rg = ref_trace;
    float *owned = NULL;
    if (use_traces) {
        tg = (float *)target_traces;
    } else {
        owned = new float[n_frames > 0 ? n_frames : 1];
        tg = owned;
        inplace_center_and_trace_atom_major(target, tg, (int)n_frames, n_atoms);
        inplace_center_and_trace_atom_major(ref_frame, &rg, 1, n_atoms);
    }
    if (parallel) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n_frames; ++i) {
            float msd = msd_atom_major(n_atoms, n_atoms, target + (size_t)i * n_atoms * 3, ref_frame, tg[i], rg, 0, NULL);
            out[i] = sqrtf(msd);
        }
    } else {
        for (int64_t i = 0; i < n_frames; ++i) {
            float msd = msd_atom_major(n_atoms, n_atoms, target + (size_t)i * n_atoms * 3, ref_frame, tg[i], rg, 0, NULL);
            out[i] = sqrtf(msd);
        }
    }
    delete[] owned;
}

/* superpose_atom_major: mdtraj/rmsd/_rmsd.pyx:620-674.  rot (F,9) is the
 * scratch the Cython allocates at :661; here the caller may keep it. */
void refloops_superpose_atom_major(const float *align_target_frame, float g_target, const float *align_mobile,
                                   const float *g_mobile, int64_t n_frames, int n_align, float *displace,
                                   int n_displace, int parallel, float *rot)
{
#pragma omp parallel for schedule(static) if (parallel)
    for (int64_t i = 0; i < n_frames; ++i) {
        msd_atom_major(n_align, n_align, align_mobile + (size_t)i * n_align * 3, align_target_frame, g_target,
                       g_mobile[i], 1, rot + 9 * i);
        rot_atom_major(n_displace, displace + (size_t)i * n_displace * 3, rot + 9 * i);
    }
}

/* getMultipleAlignDisplaceRMSDs_atom_major: mdtraj/rmsd/_rmsd.pyx:679-759
 * (a = xyz_align1[frame], b = xyz_align2[i]); n_align_padded % 4 == 0 required. */
void refloops_align_displace(const float *align1_frame, float g1, const float *align2, const float *g2,
                             const float *displ1_frame, const float *displ2, int64_t n_frames, int n_align,
                             int n_align_padded, int n_displ, int n_displ_padded, int parallel, float *out, float *rot)
{
#pragma omp parallel for schedule(static) if (parallel)
    for (int64_t i = 0; i < n_frames; ++i) {
        msd_atom_major(n_align, n_align_padded, align1_frame, align2 + (size_t)i * n_align_padded * 3, g1, g2[i], 1,
                       rot + 9 * i);
        float msd = rot_msd_atom_major(n_displ, n_displ_padded, displ1_frame, displ2 + (size_t)i * n_displ_padded * 3,
                                       rot + 9 * i);
        out[i] = sqrtf(msd);
    }
}

/* superpose=False branch: mdtraj/rmsd/_rmsd.pyx:234-241 with the body of
 * msd_nosuperpose (:765-793) -- that function is `cdef` Cython, not C, so the
 * four-line float32 loop is restated here. */
void refloops_rmsd_nosuperpose(const float *target, int64_t n_frames, int n_atoms, const float *ref_frame,
                               int parallel, float *out)
{
#pragma omp parallel for schedule(static) if (parallel)
    for (int64_t i = 0; i < n_frames; ++i) {
        const float *cur = target + (size_t)i * n_atoms * 3;
        float acc = 0;
        for (int k = 0; k < 3 * n_atoms; ++k) {
            float d = cur[k] - ref_frame[k];
            acc += d * d;
        }
        out[i] = sqrtf(acc / n_atoms);
    }
}
} /* extern "C" */
