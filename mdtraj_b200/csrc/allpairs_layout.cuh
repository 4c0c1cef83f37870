// allpairs_layout.cuh -- workspace geometry shared by the SIMT and tensor-core all-pairs paths.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace b200 {

__host__ __device__ inline int ap_kpad(int n_sel) { return (n_sel + 31) / 32 * 32; }

// tensor-core operand rows: 10 frames (30 rows) + 2 zero rows per 32; 128-row tiles hold 40 frames
__host__ __device__ inline int64_t ap_tc_row(int64_t frame, int comp) { return 32 * (frame / 10) + 3 * (frame % 10) + comp; }
__host__ __device__ inline int64_t ap_tc_rows_pad(int64_t n_frames) { return (n_frames + 39) / 40 * 128; }

// dense tensor-core layout (allpairs_tc144.cu): row 3f+c, no padding rows; 160 spare rows cover the boxes of the last
// i-tile (40 frames, four 32-row boxes) and j-tile (48 frames, one 144-row box)
__host__ __device__ inline int64_t ap_tc144_row(int64_t frame, int comp) { return 3 * frame + comp; }
__host__ __device__ inline int64_t ap_tc144_rows_pad(int64_t n_frames) { return (3 * n_frames + 160 + 7) / 8 * 8; }

inline size_t ap_align256(size_t x) { return (x + 255) / 256 * 256; }

// which kernel serves a problem of this size (deterministic on the host: prepare and rows must agree)
inline bool ap_use_tc(int64_t n_frames)
{
    const char* s = getenv("B200RMSD_ALLPAIRS");
    if (s && s[0] == 's') return false;  // "simt"
    if (s && s[0] == 't') return true;   // "tc"
    return n_frames >= 512;
}

// which operand layout the tensor-core path uses (deterministic on the host, like ap_use_tc): the dense 40 x 48-frame
// tiles of allpairs_tc144.cu, or -- B200RMSD_TC_LAYOUT=grouped, kept for A/B timing -- the 40 x 40-frame tiles of
// allpairs_tc.cu
inline bool ap_tc_dense()
{
    const char* s = getenv("B200RMSD_TC_LAYOUT");
    return !(s && s[0] == 'g');
}

struct ApGeometry {
    bool tc;
    bool dense;
    int k_pad;
    int64_t rows_pad;
    size_t traces_off, x_off, hi_off, lo_off, total;
};
inline ApGeometry ap_geometry(int64_t n_frames, int n_sel)
{
    ApGeometry g{};
    g.tc = ap_use_tc(n_frames);
    g.k_pad = ap_kpad(n_sel);
    g.dense = ap_tc_dense();
    g.rows_pad = g.dense ? ap_tc144_rows_pad(n_frames) : ap_tc_rows_pad(n_frames);
    g.traces_off = 256;
    size_t off = 256 + ap_align256((size_t)n_frames * 4);
    if (g.tc) {
        g.hi_off = off;
        off += ap_align256((size_t)g.rows_pad * g.k_pad * 4);
        g.lo_off = off;
        off += ap_align256((size_t)g.rows_pad * g.k_pad * 4);
    } else {
        g.x_off = off;
        off += ap_align256((size_t)n_frames * 3 * g.k_pad * 4);
    }
    g.total = off;
    return g;
}

cudaError_t launch_allpairs_tc_prepare(const float* xyz, int64_t n_frames, int64_t frame_stride, const int* idx, int n_sel,
                                       int k_pad, float* hi, float* lo, float* traces, int64_t rows_pad, int sm_count,
                                       cudaStream_t st);
int launch_allpairs_tc_block(const float* hi, const float* lo, const float* traces, int64_t n_frames, int n_sel, int k_pad,
                             int64_t rows_pad, int64_t row0, int64_t row1, int64_t col0, int64_t col1, float* out,
                             int64_t ld, float* out_t, int64_t ld_t, unsigned flags, int sm_count, cudaStream_t st);

cudaError_t launch_allpairs_tc144_prepare(const float* xyz, int64_t n_frames, int64_t frame_stride, const int* idx,
                                          int n_sel, int k_pad, float* hi, float* lo, float* traces, int64_t rows_pad,
                                          int sm_count, cudaStream_t st);
int launch_allpairs_tc144_block(const float* hi, const float* lo, const float* traces, int64_t n_frames, int n_sel,
                                int k_pad, int64_t rows_pad, int64_t row0, int64_t row1, int64_t col0, int64_t col1,
                                float* out, int64_t ld, float* out_t, int64_t ld_t, unsigned flags, int sm_count,
                                cudaStream_t st);

}  // namespace b200
