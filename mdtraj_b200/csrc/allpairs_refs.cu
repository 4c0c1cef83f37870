// allpairs_refs.cu -- reference structures for the tensor-core all-pairs operands.
//
// The tcgen05 accumulator truncates (DESIGN.md "tensor-core accumulation"), so the all-pairs GEMM stores every column
// frame as its DIFFERENCE from a reference structure it is close to, and adds the missing X'_i c^T block in one last
// K-step.  The error that is left behaves like 3.4e-6 nm * delta^2 / rmsd_ij, where delta is the RMSD of the column
// frame to its reference: with ONE reference (frame 0, round 1) pairs inside a second basin, or late frames of a
// drifting trajectory, kept the plain-product error class (4e-5 nm).  Here the references are chosen from the data:
//
//   greedy farthest-point traversal in RMSD space -- reference 0 is frame 0, reference r+1 is the frame whose RMSD to
//   its nearest reference so far is largest -- one one-vs-many pass (b200rmsd_rmsd_dev with rotations) per reference,
//   each frame keeping the rotation onto, and the RMSD to, its NEAREST reference (its "owner").
//
// The traversal stops when (a) kApMaxRefs (or the configured maximum) is reached; (b) frames remain that are "far" from
// every reference (further than half the reference's radius of gyration: they are stored as they are, see the prepare
// kernel) but the last three references captured no frames besides themselves -- iid-like data, more references would
// not help; (c) every frame is near a reference and the covering radius is below kCoverStop (0.25 nm: worst error
// 2e-7 nm / rmsd_ij), or has stopped shrinking (thermal noise floor of the ensemble).
// The loop is driven from the host (one 32-byte status copy + stream synchronisation per reference), so a
// single-basin trajectory pays for exactly one pass.
#include <algorithm>
#include <cstring>

#include "../../include/b200rmsd.h"
#include "allpairs_layout.cuh"
#include "common.cuh"
#include "kernels.cuh"

namespace b200 {

constexpr float kCoverStop = 0.25f;  // nm

struct ApSelectStatus {
    unsigned long long argmax;  // (float bits of rmsd-to-owner) << 32 | (0xffffffff - frame): largest rmsd, lowest frame
    int n_far;                  // frames that are not near their owner
    int captured;               // frames that moved to the newest reference and are near it
};

// After the one-vs-many pass against reference r: keep, per frame, the nearest reference seen so far.
__global__ void __launch_bounds__(256)
ap_select_update_kernel(int r, int64_t n_frames, int n_sel, const float* __restrict__ tmp_rmsd,
                        const float* __restrict__ tmp_rot, const RefStats* __restrict__ stats, float* __restrict__ best,
                        int* __restrict__ owner, float* __restrict__ rot, ApSelectStatus* __restrict__ status)
{
    __shared__ unsigned long long s_arg[8];
    __shared__ int s_far[8], s_cap[8];
    unsigned long long arg = 0;
    int far = 0, cap = 0;
    for (int64_t f = (int64_t)blockIdx.x * 256 + threadIdx.x; f < n_frames; f += (int64_t)gridDim.x * 256) {
        const float t = tmp_rmsd[f];
        float b = r == 0 ? t : best[f];
        const int raw = r == 0 ? 0 : owner[f];
        int o = raw < 0 ? -1 - raw : raw;
        const bool moved = r == 0 || t < b;
        if (moved) {
            b = t;
            o = r;
            best[f] = t;
#pragma unroll
            for (int i = 0; i < 9; ++i) rot[f * 9 + i] = tmp_rot[f * 9 + i];
        }
        // near: closer to the reference than half the reference's radius of gyration (N rmsd^2 < G_c / 4)
        const bool near = b * b < 0.25f * (float)stats[o].G / (float)n_sel;
        owner[f] = near ? o : -1 - o;  // far frames keep their nearest reference as -1-o
        far += near ? 0 : 1;
        cap += (moved && near) ? 1 : 0;
        const unsigned long long key = ((unsigned long long)__float_as_uint(fmaxf(b, 0.f)) << 32) |
                                       (unsigned long long)(0xffffffffu - (unsigned)f);
        arg = key > arg ? key : arg;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, arg, o);
        arg = other > arg ? other : arg;
        far += __shfl_xor_sync(0xffffffffu, far, o);
        cap += __shfl_xor_sync(0xffffffffu, cap, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_arg[warp] = arg; s_far[warp] = far; s_cap[warp] = cap; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            arg = s_arg[w] > arg ? s_arg[w] : arg;
            far += s_far[w];
            cap += s_cap[w];
        }
        atomicMax(&status->argmax, arg);
        atomicAdd(&status->n_far, far);
        atomicAdd(&status->captured, cap);
    }
}

// Reference r >= 1 is turned into the orientation of the reference it is nearest to (the rotation its frame got in the
// passes so far), so that all references -- and with them all frames, which are rotated onto their owners -- share one
// orientation up to second-order terms.  The all-pairs epilogue solves for lambda - tr M (qcp_msd_shift), which is only
// small, and only then cheap to get to float32 accuracy, when the two frames of a pair are not rotated against each
// other.  One block; the float32 residual sum of the centred reference is recomputed for the rotated coordinates.
__global__ void __launch_bounds__(256)
ap_rotate_ref_kernel(float* __restrict__ ref, int n_sel, const float* __restrict__ rot9, RefStats* __restrict__ stats)
{
    __shared__ double s_sum[8][3];
    float R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = rot9[i];
    double sx = 0, sy = 0, sz = 0;
    for (int k = threadIdx.x; k < n_sel; k += 256) {
        const float x = ref[3 * k], y = ref[3 * k + 1], z = ref[3 * k + 2];
        const float nx = fmaf(x, R[0], fmaf(y, R[3], z * R[6]));
        const float ny = fmaf(x, R[1], fmaf(y, R[4], z * R[7]));
        const float nz = fmaf(x, R[2], fmaf(y, R[5], z * R[8]));
        ref[3 * k] = nx; ref[3 * k + 1] = ny; ref[3 * k + 2] = nz;
        sx += nx; sy += ny; sz += nz;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_sum[warp][0] = sx; s_sum[warp][1] = sy; s_sum[warp][2] = sz; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { sx += s_sum[w][0]; sy += s_sum[w][1]; sz += s_sum[w][2]; }
        stats->sum[0] = sx; stats->sum[1] = sy; stats->sum[2] = sz;
    }
}

// owner >= 0: stored as a difference from reference `owner`; the blocks of four references a j-tile (48 frames) needs
__global__ void __launch_bounds__(64)
ap_tile_aug_kernel(const int* __restrict__ owner, int64_t n_frames, int2* __restrict__ tile_aug)
{
    const int64_t tj = blockIdx.x;
    int lo = 1 << 30, hi = -1;
    const int64_t f = tj * 48 + threadIdx.x;
    if (threadIdx.x < 48 && f < n_frames) {
        const int o = owner[f];
        if (o >= 0) { lo = o >> 2; hi = o >> 2; }
    }
    __shared__ int s_lo[2], s_hi[2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) tile_aug[tj] = make_int2(min(s_lo[0], s_lo[1]), max(s_hi[0], s_hi[1]));
}

int ap_select_references(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const int* idx, int n_sel,
                         const ApGeometry& g, char* base, int max_refs, int sm_count, cudaStream_t st, int* n_refs_out)
{
    max_refs = std::max(1, std::min(max_refs, kApMaxRefs));
    RefStats* stats = (RefStats*)(base + g.stats_off);
    float* best = (float*)(base + g.rmsd_off);
    int* owner = (int*)(base + g.owner_off);
    float* rot = (float*)(base + g.rot_off);
    double* cen = (double*)(base + g.cen_off);
    float* tmp_rot = (float*)(base + g.tmp_rot_off);
    float* tmp_rmsd = (float*)(base + g.tmp_rmsd_off);
    ApSelectStatus* status = (ApSelectStatus*)(base + g.status_off);
    ApHeader hdr{};
    float refine_hist[kApMaxRefs + 1] = {0};  // covering radius after each reference added while every frame was near
    int n_refine = 0;
    int64_t ref_frame = 0;
    int R = 0, unproductive = 0;
    const int productive_min = (int)std::max<int64_t>(2, n_frames / 1000);
    int64_t update_ctas = std::min<int64_t>((n_frames + 255) / 256, (int64_t)sm_count * 8);
    for (;;) {
        float* ref = (float*)(base + g.ref_off + (size_t)R * g.ref_stride);
        if (int rc = b200rmsd_prepare_reference_dev(xyz + ref_frame * frame_stride, idx, n_sel, 1, 0.f, ref, stats + R, st))
            return rc;
        if (R > 0) {  // into the orientation of its nearest reference so far (rot[] still holds the passes before this one)
            ap_rotate_ref_kernel<<<1, 256, 0, st>>>(ref, n_sel, rot + ref_frame * 9, stats + R);
            if (cudaGetLastError() != cudaSuccess) return set_error(B200RMSD_ECUDA, "allpairs_prepare: reference rotation launch failed");
        }
        if (int rc = b200rmsd_rmsd_dev(xyz, n_frames, n_atoms, frame_stride, idx, n_sel, ref, stats + R, nullptr, 0,
                                       tmp_rmsd, tmp_rot, cen, nullptr, base + g.scratch_off, g.scratch_bytes, st))
            return rc;
        cudaError_t e = cudaMemsetAsync(status, 0, sizeof(ApSelectStatus), st);
        if (e == cudaSuccess) {
            ap_select_update_kernel<<<(unsigned)update_ctas, 256, 0, st>>>(R, n_frames, n_sel, tmp_rmsd, tmp_rot, stats, best,
                                                                          owner, rot, status);
            e = cudaGetLastError();
        }
        ApSelectStatus h{};
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h, status, sizeof(h), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return set_error(B200RMSD_ECUDA, "allpairs_prepare (reference selection): %s", cudaGetErrorString(e));
        hdr.ref_frame[R] = (int)ref_frame;
        ++R;
        const unsigned bits = (unsigned)(h.argmax >> 32);
        float radius;
        memcpy(&radius, &bits, 4);
        hdr.n_far = h.n_far;
        hdr.cover_radius = radius;
        if (R >= max_refs || !(radius > 0.f)) break;
        if (h.n_far > 0) {
            // uncovered frames exist: go on unless the references stop capturing anything (iid-like data)
            if (R >= 2) unproductive = h.captured >= productive_min ? 0 : unproductive + 1;
            if (unproductive >= 3) break;
        } else {
            refine_hist[n_refine++] = radius;
            if (radius <= kCoverStop) break;
            if (n_refine >= 3 && radius > 0.9f * refine_hist[n_refine - 3]) break;  // thermal noise floor of the ensemble
        }
        ref_frame = (int64_t)(0xffffffffu - (unsigned)(h.argmax & 0xffffffffu));
    }
    hdr.n_refs = R;
    cudaError_t e = cudaMemcpyAsync(base, &hdr, sizeof(hdr), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        ap_tile_aug_kernel<<<(unsigned)ap_tc144_jtiles(n_frames), 64, 0, st>>>(owner, n_frames, (int2*)(base + g.tile_aug_off));
        e = cudaGetLastError();
    }
    // hdr lives on this stack frame: the copy above must have read it before we return
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return set_error(B200RMSD_ECUDA, "allpairs_prepare (reference selection): %s", cudaGetErrorString(e));
    *n_refs_out = R;
    return 0;
}

}  // namespace b200
