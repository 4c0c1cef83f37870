// allpairs_layout.cuh -- workspace geometry shared by the SIMT and tensor-core all-pairs paths.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/b200rmsd.h"

namespace b200 {

__host__ __device__ inline int ap_kpad(int n_sel) { return (n_sel + 31) / 32 * 32; }

// Tensor-core operands (allpairs_tc144.cu): row 3f+c = component c of frame f, no padding rows; 160 spare rows cover the
// boxes of the last i-tile (40 frames, four 32-row boxes) and j-tile (48 frames, one 144-row box).  K holds the n_sel
// atoms, then -- starting at the next multiple of 8, i.e. in a K-step of their own -- six "augmentation" columns that
// add a per-frame 3x3 matrix to every block inside the GEMM (see allpairs_tc144_prepare_kernel), padded to 32.
__host__ __device__ inline int64_t ap_tc144_row(int64_t frame, int comp) { return 3 * frame + comp; }
__host__ __device__ inline int64_t ap_tc144_rows_pad(int64_t n_frames) { return (3 * n_frames + 160 + 7) / 8 * 8; }
__host__ __device__ inline int ap_tc144_k0(int n_sel) { return (n_sel + 7) / 8 * 8; }
__host__ __device__ inline int ap_tc144_kpad(int n_sel) { return (ap_tc144_k0(n_sel) + 6 + 31) / 32 * 32; }

inline size_t ap_align256(size_t x) { return (x + 255) / 256 * 256; }

// which kernel serves a problem of this size (deterministic on the host: prepare and rows must agree)
inline bool ap_use_tc(int64_t n_frames)
{
    const char* s = getenv("B200RMSD_ALLPAIRS");
    if (s && s[0] == 's') return false;  // "simt"
    if (s && s[0] == 't') return true;   // "tc"
    return n_frames >= 512;
}

struct ApGeometry {
    bool tc;
    int k_pad;
    int64_t rows_pad;
    size_t traces_off, x_off, total;
    // tensor-core path: A operand (aligned frames), B operand (their differences from the common reference), tf32 hi/lo
    size_t a_hi_off, a_lo_off, b_hi_off, b_lo_off;
    // alignment pass of the prepare step: reference selection, its statistics, per-frame rotation / centroid / rmsd,
    // scratch of the one-vs-many kernel
    size_t ref_off, stats_off, rot_off, cen_off, rmsd_off, scratch_off, scratch_bytes;
};
inline ApGeometry ap_geometry(int64_t n_frames, int n_sel)
{
    ApGeometry g{};
    g.tc = ap_use_tc(n_frames);
    g.traces_off = 256;
    size_t off = 256 + ap_align256((size_t)n_frames * 4);
    if (g.tc) {
        g.k_pad = ap_tc144_kpad(n_sel);
        g.rows_pad = ap_tc144_rows_pad(n_frames);
        const size_t op = ap_align256((size_t)g.rows_pad * g.k_pad * 4);
        g.a_hi_off = off; off += op;
        g.a_lo_off = off; off += op;
        g.b_hi_off = off; off += op;
        g.b_lo_off = off; off += op;
        g.ref_off = off; off += ap_align256((size_t)((n_sel + 3) / 4 * 4) * 3 * 4);
        g.stats_off = off; off += 256;
        g.rot_off = off; off += ap_align256((size_t)n_frames * 9 * 4);
        g.cen_off = off; off += ap_align256((size_t)n_frames * 3 * 8);
        g.rmsd_off = off; off += ap_align256((size_t)n_frames * 4);
        g.scratch_bytes = b200rmsd_scratch_bytes(n_frames, n_sel);
        g.scratch_off = off; off += ap_align256(g.scratch_bytes);
    } else {
        g.k_pad = ap_kpad(n_sel);
        g.x_off = off;
        off += ap_align256((size_t)n_frames * 3 * g.k_pad * 4);
    }
    g.total = off;
    return g;
}

// operands + traces from the frames and their alignment onto the common reference
cudaError_t launch_allpairs_tc144_prepare(const float* xyz, int64_t n_frames, int64_t frame_stride, const int* idx,
                                          int n_sel, int k_pad, const float* ref, const void* ref_stats,
                                          const float* rmsd_to_ref, const float* rot, const double* centroid, float* a_hi,
                                          float* a_lo, float* b_hi, float* b_lo, float* traces, int64_t rows_pad,
                                          int sm_count, cudaStream_t st);
int launch_allpairs_tc144_block(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo,
                                const float* traces, int n_sel, int k_pad, int64_t rows_pad, int64_t row0, int64_t row1,
                                int64_t col0, int64_t col1, float* out, int64_t ld, float* out_t, int64_t ld_t,
                                unsigned flags, int sm_count, cudaStream_t st);

}  // namespace b200
