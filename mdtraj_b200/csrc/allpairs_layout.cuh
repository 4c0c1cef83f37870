// allpairs_layout.cuh -- workspace geometry shared by the SIMT and tensor-core all-pairs paths.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/b200rmsd.h"

namespace b200 {

__host__ __device__ inline int ap_kpad(int n_sel) { return (n_sel + 31) / 32 * 32; }

// Tensor-core operands (allpairs_tc144.cu): row 3f+c = component c of frame f, no padding rows; 160 spare rows cover the
// boxes of the last i-tile (40 frames, four 32-row boxes) and j-tile (48 frames, one 144-row box).  The "atom" matrices
// hold the n_sel atoms along K, padded to 32.  The "augmentation" matrices hold, per reference structure r (at most
// kApMaxRefs, chosen by ap_select_references), eight K columns 8r .. 8r+7 that add the 3x3 matrix X'_i c_r^T to every
// block whose column frame is stored as a difference from c_r (see allpairs_tc144_prepare_kernel); four references
// share one 32-column K block, and a tile only walks the blocks its column frames refer to.
constexpr int kApMaxRefs = 32;
constexpr int kApAugCols = 8 * kApMaxRefs;
__host__ __device__ inline int64_t ap_tc144_row(int64_t frame, int comp) { return 3 * frame + comp; }
__host__ __device__ inline int64_t ap_tc144_rows_pad(int64_t n_frames) { return (3 * n_frames + 160 + 7) / 8 * 8; }
__host__ __device__ inline int64_t ap_tc144_jtiles(int64_t n_frames) { return (n_frames + 47) / 48; }

inline size_t ap_align256(size_t x) { return (x + 255) / 256 * 256; }

// process-wide settings of b200rmsd_allpairs_configure (allpairs.cu)
extern int g_ap_min_tc_frames;  // trajectories with at least this many frames take the tensor-core path (default 512)
extern int g_ap_max_refs;       // upper bound on reference structures (default kApMaxRefs)
extern int g_ap_cta_pair;       // tensor-core kernel on 2-CTA clusters (default 1: 2-15 % faster)

// which kernel serves a problem of this size (deterministic on the host: prepare and rows must agree)
inline bool ap_use_tc(int64_t n_frames) { return n_frames >= g_ap_min_tc_frames; }

// first 256 bytes of the workspace (device memory; written by the prepare step)
struct ApHeader {
    int n_refs;            // reference structures in use (tensor-core path; 0 on the SIMT path)
    int n_far;             // frames farther than half a radius of gyration from every reference: stored as they are
    float cover_radius;    // largest RMSD of a frame to its nearest reference
    int ref_frame[kApMaxRefs];
};
static_assert(sizeof(ApHeader) <= 256, "ApHeader must fit the workspace header");

struct ApGeometry {
    bool tc;
    int k_pad;             // K of the atom matrices (multiple of 32)
    int64_t rows_pad;
    size_t traces_off, x_off, total;
    // tensor-core path: A operand (aligned frames), B operand (their differences from the nearest reference), tf32 hi/lo
    size_t a_hi_off, a_lo_off, b_hi_off, b_lo_off;
    // augmentation matrices (rows_pad x kApAugCols): A pieces g1|g2 (hi) and g3 (lo), B unit vectors
    size_t aug_a_hi_off, aug_a_lo_off, aug_b_off;
    size_t tile_aug_off;   // int2 per absolute j-tile: first and last augmentation K block (first > last: none)
    // reference selection: the references and their statistics, per-frame owner / rotation / centroid / rmsd to the
    // owner, outputs of the one-vs-many pass in flight, its scratch, the reduction record
    size_t ref_off, ref_stride, stats_off, owner_off, rot_off, cen_off, rmsd_off, tmp_rot_off, tmp_rmsd_off, status_off,
        scratch_off, scratch_bytes;
};
inline ApGeometry ap_geometry(int64_t n_frames, int n_sel)
{
    ApGeometry g{};
    g.tc = ap_use_tc(n_frames);
    g.traces_off = 256;
    size_t off = 256 + ap_align256((size_t)n_frames * 4);
    g.k_pad = ap_kpad(n_sel);
    if (g.tc) {
        g.rows_pad = ap_tc144_rows_pad(n_frames);
        const size_t op = ap_align256((size_t)g.rows_pad * g.k_pad * 4);
        g.a_hi_off = off; off += op;
        g.a_lo_off = off; off += op;
        g.b_hi_off = off; off += op;
        g.b_lo_off = off; off += op;
        const size_t aug = ap_align256((size_t)g.rows_pad * kApAugCols * 4);
        g.aug_a_hi_off = off; off += aug;
        g.aug_a_lo_off = off; off += aug;
        g.aug_b_off = off; off += aug;
        g.tile_aug_off = off; off += ap_align256((size_t)ap_tc144_jtiles(n_frames) * 8);
        g.ref_stride = ap_align256((size_t)((n_sel + 3) / 4 * 4) * 3 * 4);
        g.ref_off = off; off += g.ref_stride * kApMaxRefs;
        g.stats_off = off; off += ap_align256((size_t)kApMaxRefs * B200RMSD_REFSTATS_BYTES);
        g.owner_off = off; off += ap_align256((size_t)n_frames * 4);
        g.rot_off = off; off += ap_align256((size_t)n_frames * 9 * 4);
        g.cen_off = off; off += ap_align256((size_t)n_frames * 3 * 8);
        g.rmsd_off = off; off += ap_align256((size_t)n_frames * 4);
        g.tmp_rot_off = off; off += ap_align256((size_t)n_frames * 9 * 4);
        g.tmp_rmsd_off = off; off += ap_align256((size_t)n_frames * 4);
        g.status_off = off; off += 256;
        g.scratch_bytes = b200rmsd_scratch_bytes(n_frames, n_sel);
        g.scratch_off = off; off += ap_align256(g.scratch_bytes);
    } else {
        g.x_off = off;
        off += ap_align256((size_t)n_frames * 3 * g.k_pad * 4);
    }
    g.total = off;
    return g;
}

// reference selection + alignment of every frame onto its nearest reference (allpairs_refs.cu); synchronises `st`
int ap_select_references(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const int* idx, int n_sel,
                         const ApGeometry& g, char* base, int max_refs, int sm_count, cudaStream_t st, int* n_refs_out);

// operands + traces from the frames and their alignment onto their references
cudaError_t launch_allpairs_tc144_prepare(const float* xyz, int64_t n_frames, int64_t frame_stride, const int* idx,
                                          int n_sel, const ApGeometry& g, char* base, int n_refs, int sm_count,
                                          cudaStream_t st);
int launch_allpairs_tc144_block(const ApGeometry& g, const char* base, int n_sel, int64_t n_frames, int64_t row0,
                                int64_t row1, int64_t col0, int64_t col1, float* out, int64_t ld, float* out_t,
                                int64_t ld_t, float* out_rot, unsigned flags, int sm_count, cudaStream_t st);

}  // namespace b200
