// kernels.cuh -- launch-parameter structs and launcher prototypes (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int kWarpsPerCta = 16;               // streaming kernels: 512 threads, 1 CTA / SM
constexpr int kThreadsPerCta = kWarpsPerCta * 32;
constexpr int kBatch = 16;                     // frames solved together (one lane each) per warp
constexpr int kSumStride = 17;                 // padded stride of a 16-value sum record in smem
constexpr int kFlushUnits = 8;                 // 4-atom units a lane accumulates in float32 before the float64 fold
constexpr int kUnitFloats = 12;                // one "unit" = 4 atoms = 48 bytes = 3 x float4
constexpr int kMaxSegUnits = 896;              // reference segment resident in smem: <= 3584 atoms (42 KB), leaves a 3-stage ring

// reference-frame statistics written by prepare_ref_kernel (device memory)
struct RefStats {
    double G;       // trace of the centred reference selection
    double sum[3];  // residual sum of the float32-centred reference (~0)
    double mean[3]; // centroid that was removed (float64)
};

struct OvmParams {
    const float* xyz;       // (F, frame_stride) floats, atom-major, padded to a multiple of 4 atoms
    int64_t n_frames;
    int64_t frame_stride;   // floats between consecutive frames (multiple of 4)
    int n_atoms;            // real atoms taking part (== n_sel when idx != nullptr)
    const int* idx;         // optional atom selection (gather path), length n_atoms
    int frame_atoms;        // atoms per frame in memory (>= every index); 0: same as n_atoms (no selection)
    const float* ref;       // centred reference selection, (n_pad,3) floats, zero padded
    const RefStats* ref_stats;
    const float* traces;    // precentered mode: per-frame traces, else nullptr
    float* out_rmsd;        // (F)
    float* out_rot;         // (F,9) or nullptr
    double* out_centroid;   // (F,3) or nullptr
    unsigned int* degenerate;  // counter of identity-fallback rotations, or nullptr
    // tiling of the TMA path
    int n_seg;              // atom segments (1 => fused epilogue)
    int seg_units;          // units per segment (last may be shorter)
    int total_units;        // ceil(n_atoms/4)
    int chunk_units;        // units per bulk copy
    int stages;             // ring depth per warp
    double* partials;       // (F, n_seg, 16) float64 when n_seg > 1
    double inv_n;           // 1 / n_atoms (set by the launchers)
};

struct ApplyParams {
    float* xyz;             // in/out (F, frame_stride)
    int64_t n_frames;
    int64_t frame_stride;
    int n_atoms;
    const float* rot;       // (F,9)
    const double* centroid; // (F,3) removed before rotation
    const RefStats* ref_stats;  // mean[] added after rotation
};

// frame-resident read-modify-write kernels (frame_resident.cu)
enum { OP_SUPERPOSE = 0, OP_CENTER = 1 };
struct FusedParams {
    float* xyz;             // in/out, padded atom-major
    int64_t n_frames;
    int64_t frame_stride;
    int n_atoms;
    int n_pad;
    const int* idx;         // align selection (OP_SUPERPOSE) or nullptr
    int n_sel;              // == n_atoms when idx == nullptr
    const float* ref;       // centred packed reference selection
    const RefStats* ref_stats;
    float* out_rmsd;        // may be nullptr
    float* out_rot;         // may be nullptr
    float* traces;          // OP_CENTER output, may be nullptr
    unsigned int* degenerate;
    int batch;              // G: slot groups computing concurrently (<= 16)
    int nbuf;               // shared-memory slot buffers in the ring (a multiple of G)
    int fpb;                // frames per slot (> 1 only when frame_stride == 3*n_pad)
    int team_warps;         // warps sharing one frame inside a group: 16/G or 1
    int lanes;              // team_warps == 1: lanes sharing one frame inside a warp (2, 4, 8, 16, 32)
    double inv_n_sel;       // 1 / n_sel (set by launch_frame_resident)
    // stage-pipelined superpose (superpose_pipe_kernel): pipe != 0, then batch = solver warps S, team_warps = streaming
    // warps per frame W, lanes = lanes per frame when W == 1, nbuf = slot buffers, depth = slots between the sums of a
    // slot and its transform
    int pipe;
    int depth;
    int ref_global;         // pipe: the reference stays in global memory (L2-resident, read through L1) because frame buffers
                            // plus a resident copy would not fit shared memory (all atoms selected on ~5,000-atom frames)
};
bool fused_config(FusedParams& p, int op);
bool fused_override(FusedParams& p, int op, int G, int nbuf, int fpb, int lanes);
cudaError_t launch_frame_resident(const FusedParams& p, int op, int sm_count, cudaStream_t st);

cudaError_t launch_ovm_tma(const OvmParams& p, bool precentered, int sm_count, cudaStream_t st);
bool launch_ovm_group(OvmParams& p, bool precentered, int sm_count, cudaStream_t st, cudaError_t* err);
cudaError_t launch_ovm_gather(const OvmParams& p, bool precentered, int sm_count, cudaStream_t st);
cudaError_t launch_prepare_ref(const float* frame, const int* idx, int n_sel, int do_center, float given_trace,
                               float* ref_out, RefStats* stats, cudaStream_t st);
cudaError_t launch_center_trace(float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, float* traces,
                                int sm_count, cudaStream_t st);
cudaError_t launch_nosuperpose(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const int* idx,
                               const float* ref_raw, float* out, int sm_count, cudaStream_t st);
cudaError_t launch_apply_transform(const ApplyParams& p, int sm_count, cudaStream_t st);
size_t ovm_tma_smem_bytes(const OvmParams& p);
int rmsf_chunks(int64_t n_frames);
cudaError_t launch_rmsf(const float* xyz, int64_t n_frames, int64_t frame_stride, const int* idx, int n_sel,
                        const float* rot, const double* centroid, double* partials, float* out, cudaStream_t st);
cudaError_t launch_rot_msd(const float* a, const float* b, int64_t n_frames, int n_atoms, int64_t frame_stride,
                           const float* rot, int transpose, float* rot_out, float* out, int sm_count, cudaStream_t st);

// records a thread-local message for b200rmsd_last_error() and returns `code` (defined in capi.cu)
int set_error(int code, const char* fmt, ...);

}  // namespace b200
