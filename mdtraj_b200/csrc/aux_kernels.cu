// aux_kernels.cu -- the small streaming kernels either side of the QCP solve.
//
//   prepare_ref_kernel     centre the single reference frame (selection applied) the way
//                          inplace_center_and_trace_atom_major does for n_frames == 1
//                          (center_generic.h:3-44, called at _rmsd.pyx:213), pack it
//                          zero-padded for the streaming kernels.
//   center_trace_kernel    Trajectory.center_coordinates / _center_inplace_atom_major
//                          (_rmsd.pyx:487-491 -> center.h:7): in-place centring + traces,
//                          float64 sums, float32 mean, float32 subtraction, float64 trace.
//   nosuperpose_kernel     msd_nosuperpose loop (_rmsd.pyx:234-241, :765-793).
//   apply_transform_kernel x' = (x - c) . R + c_ref : the rot_atom_major pass of
//                          superpose_atom_major (_rmsd.pyx:668 -> rotation_generic.h:29-44)
//                          fused with the two numpy broadcasts around it
//                          (core/trajectory.py:1140-1144 and :1171).
#include "common.cuh"
#include "kernels.cuh"

namespace b200 {

// block-wide sum of a double, result valid in every thread
template <int THREADS>
__device__ __forceinline__ double block_sum(double x, double* scratch)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    x = warp_sum(x);
    __syncthreads();
    if (lane == 0) scratch[warp] = x;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) t += scratch[w];
    return t;
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prepare_ref_kernel(const float* __restrict__ frame, const int* __restrict__ idx,
                                                          int n_sel, int do_center, float given_trace,
                                                          float* __restrict__ ref_out, RefStats* __restrict__ stats)
{
    __shared__ double scratch[8];
    const int n_pad = (n_sel + 3) & ~3;
    double sx = 0, sy = 0, sz = 0;
    for (int k = threadIdx.x; k < n_sel; k += 256) {
        const int a = idx ? idx[k] : k;
        sx += frame[3 * a]; sy += frame[3 * a + 1]; sz += frame[3 * a + 2];
    }
    sx = block_sum<256>(sx, scratch); sy = block_sum<256>(sy, scratch); sz = block_sum<256>(sz, scratch);
    const double mean_x = sx / n_sel, mean_y = sy / n_sel, mean_z = sz / n_sel;
    const float mx = do_center ? (float)mean_x : 0.f, my = do_center ? (float)mean_y : 0.f,
                mz = do_center ? (float)mean_z : 0.f;
    double tr = 0, rx = 0, ry = 0, rz = 0;
    for (int k = threadIdx.x; k < n_pad; k += 256) {
        float x = 0.f, y = 0.f, z = 0.f;
        if (k < n_sel) {
            const int a = idx ? idx[k] : k;
            x = frame[3 * a] - mx; y = frame[3 * a + 1] - my; z = frame[3 * a + 2] - mz;
            const float sq = x * x + y * y + z * z;
            tr += (double)sq;
            rx += x; ry += y; rz += z;
        }
        ref_out[3 * k] = x; ref_out[3 * k + 1] = y; ref_out[3 * k + 2] = z;
    }
    tr = block_sum<256>(tr, scratch);
    rx = block_sum<256>(rx, scratch); ry = block_sum<256>(ry, scratch); rz = block_sum<256>(rz, scratch);
    if (threadIdx.x == 0) {
        stats->G = do_center ? tr : (double)given_trace;
        stats->sum[0] = do_center ? rx : 0.0; stats->sum[1] = do_center ? ry : 0.0; stats->sum[2] = do_center ? rz : 0.0;
        stats->mean[0] = mean_x; stats->mean[1] = mean_y; stats->mean[2] = mean_z;
    }
}

cudaError_t launch_prepare_ref(const float* frame, const int* idx, int n_sel, int do_center, float given_trace,
                               float* ref_out, RefStats* stats, cudaStream_t st)
{
    prepare_ref_kernel<<<1, 256, 0, st>>>(frame, idx, n_sel, do_center, given_trace, ref_out, stats);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// centring: GROUP threads per frame (32 => warp shuffles only, 256 => one CTA per frame)
// pass 1 reads the frame, pass 2 re-reads it (L1/L2-hot), subtracts, writes, accumulates the trace
// ---------------------------------------------------------------------------
template <int GROUP>
__global__ void __launch_bounds__(256) center_trace_kernel(float* __restrict__ xyz, int64_t n_frames, int n_atoms,
                                                           int64_t frame_stride, float* __restrict__ traces)
{
    __shared__ double scratch[8];
    constexpr int GROUPS_PER_CTA = 256 / GROUP;
    const int g_in_cta = threadIdx.x / GROUP, t = threadIdx.x % GROUP;
    const int64_t n_groups = (int64_t)gridDim.x * GROUPS_PER_CTA;
    const int units = (n_atoms + 3) >> 2;

    for (int64_t f = (int64_t)blockIdx.x * GROUPS_PER_CTA + g_in_cta; f < n_frames; f += n_groups) {
        float4* fr = reinterpret_cast<float4*>(xyz + f * frame_stride);
        double sx = 0, sy = 0, sz = 0;
        for (int u = t; u < units; u += GROUP) {
            const float4 a0 = fr[3 * u], a1 = fr[3 * u + 1], a2 = fr[3 * u + 2];
            // padding atoms are zero in the staged layout, so they do not disturb the sums
            sx += (double)a0.x; sy += (double)a0.y; sz += (double)a0.z;
            sx += (double)a0.w; sy += (double)a1.x; sz += (double)a1.y;
            sx += (double)a1.z; sy += (double)a1.w; sz += (double)a2.x;
            sx += (double)a2.y; sy += (double)a2.z; sz += (double)a2.w;
        }
        if (GROUP == 32) { sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); }
        else { sx = block_sum<256>(sx, scratch); sy = block_sum<256>(sy, scratch); sz = block_sum<256>(sz, scratch); }
        const float mx = (float)(sx / n_atoms), my = (float)(sy / n_atoms), mz = (float)(sz / n_atoms);
        double tr = 0;
        for (int u = t; u < units; u += GROUP) {
            float4 a0 = fr[3 * u], a1 = fr[3 * u + 1], a2 = fr[3 * u + 2];
            const int nvalid = n_atoms - 4 * u;
            a0.x -= mx; a0.y -= my; a0.z -= mz;
            tr += (double)(a0.x * a0.x); tr += (double)(a0.y * a0.y); tr += (double)(a0.z * a0.z);
            if (nvalid > 1) {
                a0.w -= mx; a1.x -= my; a1.y -= mz;
                tr += (double)(a0.w * a0.w); tr += (double)(a1.x * a1.x); tr += (double)(a1.y * a1.y);
            }
            if (nvalid > 2) {
                a1.z -= mx; a1.w -= my; a2.x -= mz;
                tr += (double)(a1.z * a1.z); tr += (double)(a1.w * a1.w); tr += (double)(a2.x * a2.x);
            }
            if (nvalid > 3) {
                a2.y -= mx; a2.z -= my; a2.w -= mz;
                tr += (double)(a2.y * a2.y); tr += (double)(a2.z * a2.z); tr += (double)(a2.w * a2.w);
            }
            fr[3 * u] = a0; fr[3 * u + 1] = a1; fr[3 * u + 2] = a2;
        }
        if (GROUP == 32) tr = warp_sum(tr);
        else tr = block_sum<256>(tr, scratch);
        if (t == 0 && traces) traces[f] = (float)tr;
    }
}

cudaError_t launch_center_trace(float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, float* traces,
                                int sm_count, cudaStream_t st)
{
    if (n_frames <= 0) return cudaSuccess;
    if (n_atoms >= 4096) {
        int64_t ctas = (int64_t)sm_count * 8;
        if (ctas > n_frames) ctas = n_frames;
        center_trace_kernel<256><<<(unsigned)ctas, 256, 0, st>>>(xyz, n_frames, n_atoms, frame_stride, traces);
    } else {
        int64_t ctas = (int64_t)sm_count * 8;
        const int64_t need = (n_frames + 7) / 8;
        if (ctas > need) ctas = need;
        center_trace_kernel<32><<<(unsigned)ctas, 256, 0, st>>>(xyz, n_frames, n_atoms, frame_stride, traces);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nosuperpose_kernel(const float* __restrict__ xyz, int64_t n_frames, int n_atoms,
                                                          int64_t frame_stride, const int* __restrict__ idx,
                                                          const float* __restrict__ ref_raw, float* __restrict__ out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t)gridDim.x * 8;
    for (int64_t f = (int64_t)blockIdx.x * 8 + warp; f < n_frames; f += n_warps) {
        const float* fr = xyz + f * frame_stride;
        float acc = 0.f;
#pragma unroll 4
        for (int k = lane; k < n_atoms; k += 32) {
            const int a = idx ? __ldg(idx + k) : k;
            const float dx = __ldg(fr + 3 * a) - __ldg(ref_raw + 3 * k);
            const float dy = __ldg(fr + 3 * a + 1) - __ldg(ref_raw + 3 * k + 1);
            const float dz = __ldg(fr + 3 * a + 2) - __ldg(ref_raw + 3 * k + 2);
            acc = fmaf(dx, dx, acc); acc = fmaf(dy, dy, acc); acc = fmaf(dz, dz, acc);
        }
        const double tot = warp_sum((double)acc);
        if (lane == 0) out[f] = (float)sqrt(tot / n_atoms);
    }
}

cudaError_t launch_nosuperpose(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride, const int* idx,
                               const float* ref_raw, float* out, int sm_count, cudaStream_t st)
{
    if (n_frames <= 0) return cudaSuccess;
    int64_t ctas = (int64_t)sm_count * 8;
    const int64_t need = (n_frames + 7) / 8;
    if (ctas > need) ctas = need;
    nosuperpose_kernel<<<(unsigned)ctas, 256, 0, st>>>(xyz, n_frames, n_atoms, frame_stride, idx, ref_raw, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// x' = f32(f32(x - c) . R + c_ref); centroid and target offset carried as float32 hi/lo pairs
// ---------------------------------------------------------------------------
__device__ __forceinline__ void xform_atom(float& x, float& y, float& z, const float (&R)[9], const float (&ch)[3],
                                           const float (&cl)[3], const float (&oh)[3], const float (&ol)[3])
{
    const float tx = (x - ch[0]) - cl[0], ty = (y - ch[1]) - cl[1], tz = (z - ch[2]) - cl[2];
    const float rx = tx * R[0] + ty * R[3] + tz * R[6];
    const float ry = tx * R[1] + ty * R[4] + tz * R[7];
    const float rz = tx * R[2] + ty * R[5] + tz * R[8];
    x = (rx + oh[0]) + ol[0]; y = (ry + oh[1]) + ol[1]; z = (rz + oh[2]) + ol[2];
}

template <int GROUP>
__global__ void __launch_bounds__(256) apply_transform_kernel(const ApplyParams p)
{
    constexpr int GROUPS_PER_CTA = 256 / GROUP;
    const int g_in_cta = threadIdx.x / GROUP, t = threadIdx.x % GROUP;
    const int64_t n_groups = (int64_t)gridDim.x * GROUPS_PER_CTA;
    const int units = (p.n_atoms + 3) >> 2;
    float oh[3], ol[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double o = p.ref_stats ? p.ref_stats->mean[i] : 0.0;
        oh[i] = (float)o; ol[i] = (float)(o - (double)oh[i]);
    }
    for (int64_t f = (int64_t)blockIdx.x * GROUPS_PER_CTA + g_in_cta; f < p.n_frames; f += n_groups) {
        float R[9], ch[3], cl[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(p.rot + f * 9 + i);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double c = p.centroid ? p.centroid[f * 3 + i] : 0.0;
            ch[i] = (float)c; cl[i] = (float)(c - (double)ch[i]);
        }
        float4* fr = reinterpret_cast<float4*>(p.xyz + f * p.frame_stride);
#pragma unroll 2
        for (int u = t; u < units; u += GROUP) {
            float4 a0 = ldg_stream(fr + 3 * u), a1 = ldg_stream(fr + 3 * u + 1), a2 = ldg_stream(fr + 3 * u + 2);
            const int nvalid = p.n_atoms - 4 * u;
            xform_atom(a0.x, a0.y, a0.z, R, ch, cl, oh, ol);
            if (nvalid > 1) xform_atom(a0.w, a1.x, a1.y, R, ch, cl, oh, ol);
            if (nvalid > 2) xform_atom(a1.z, a1.w, a2.x, R, ch, cl, oh, ol);
            if (nvalid > 3) xform_atom(a2.y, a2.z, a2.w, R, ch, cl, oh, ol);
            stg_stream(fr + 3 * u, a0); stg_stream(fr + 3 * u + 1, a1); stg_stream(fr + 3 * u + 2, a2);
        }
    }
}

cudaError_t launch_apply_transform(const ApplyParams& p, int sm_count, cudaStream_t st)
{
    if (p.n_frames <= 0) return cudaSuccess;
    int64_t ctas = (int64_t)sm_count * 8;
    if (p.n_atoms >= 2048) {
        if (ctas > p.n_frames) ctas = p.n_frames;
        apply_transform_kernel<256><<<(unsigned)ctas, 256, 0, st>>>(p);
    } else {
        const int64_t need = (p.n_frames + 7) / 8;
        if (ctas > need) ctas = need;
        apply_transform_kernel<32><<<(unsigned)ctas, 256, 0, st>>>(p);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// md.rmsf per-atom statistics (_rmsd.pyx:411-444): y = (x - c) . R per frame, then over frames
// mean(y) and sqrt(mean |y - mean(y)|^2).  One pass: float64 sums of y and |y|^2 per atom over a
// chunk of frames; rmsf_finalize_kernel combines the chunks (the reference accumulates both in
// float32 over all frames, two passes).
// grid = (ceil(n_sel/128), n_chunks)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rmsf_stats_kernel(const float* __restrict__ xyz, int64_t n_frames,
                                                         int64_t frame_stride, const int* __restrict__ idx, int n_sel,
                                                         const float* __restrict__ rot, const double* __restrict__ centroid,
                                                         int64_t frames_per_chunk, double* __restrict__ partials)
{
    const int k = blockIdx.x * 128 + threadIdx.x;
    const int64_t f0 = (int64_t)blockIdx.y * frames_per_chunk;
    const int64_t f1 = min(n_frames, f0 + frames_per_chunk);
    if (k >= n_sel) return;
    const int a = idx ? __ldg(idx + k) : k;
    double sx = 0, sy = 0, sz = 0, s2 = 0;
    for (int64_t f = f0; f < f1; ++f) {
        const float* fr = xyz + f * frame_stride + 3 * (int64_t)a;
        float x = __ldg(fr), y = __ldg(fr + 1), z = __ldg(fr + 2);
        if (centroid) {
            x -= (float)centroid[f * 3]; y -= (float)centroid[f * 3 + 1]; z -= (float)centroid[f * 3 + 2];
        }
        float rx = x, ry = y, rz = z;
        if (rot) {
            const float* R = rot + f * 9;
            rx = x * __ldg(R + 0) + y * __ldg(R + 3) + z * __ldg(R + 6);
            ry = x * __ldg(R + 1) + y * __ldg(R + 4) + z * __ldg(R + 7);
            rz = x * __ldg(R + 2) + y * __ldg(R + 5) + z * __ldg(R + 8);
        }
        sx += (double)rx; sy += (double)ry; sz += (double)rz;
        s2 += (double)rx * rx + (double)ry * ry + (double)rz * rz;
    }
    double* out = partials + ((size_t)blockIdx.y * n_sel + k) * 4;
    out[0] = sx; out[1] = sy; out[2] = sz; out[3] = s2;
}

__global__ void __launch_bounds__(128) rmsf_finalize_kernel(const double* __restrict__ partials, int n_chunks, int n_sel,
                                                            int64_t n_frames, float* __restrict__ out)
{
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k >= n_sel) return;
    double sx = 0, sy = 0, sz = 0, s2 = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const double* p = partials + ((size_t)c * n_sel + k) * 4;
        sx += p[0]; sy += p[1]; sz += p[2]; s2 += p[3];
    }
    const double inv = 1.0 / (double)n_frames;
    const double mx = sx * inv, my = sy * inv, mz = sz * inv;
    double var = s2 * inv - (mx * mx + my * my + mz * mz);
    out[k] = (float)sqrt(var > 0.0 ? var : 0.0);
}

int rmsf_chunks(int64_t n_frames) { return (int)((n_frames + 255) / 256 < 592 ? (n_frames + 255) / 256 : 592); }

cudaError_t launch_rmsf(const float* xyz, int64_t n_frames, int64_t frame_stride, const int* idx, int n_sel,
                        const float* rot, const double* centroid, double* partials, float* out, cudaStream_t st)
{
    if (n_frames <= 0 || n_sel <= 0) return cudaSuccess;
    const int n_chunks = rmsf_chunks(n_frames);
    const int64_t fpc = (n_frames + n_chunks - 1) / n_chunks;
    dim3 grid((unsigned)((n_sel + 127) / 128), (unsigned)n_chunks);
    rmsf_stats_kernel<<<grid, 128, 0, st>>>(xyz, n_frames, frame_stride, idx, n_sel, rot, centroid, fpc, partials);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    rmsf_finalize_kernel<<<(unsigned)((n_sel + 127) / 128), 128, 0, st>>>(partials, n_chunks, n_sel, n_frames, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// rot_msd_atom_major over frames (rotation_generic.h:47-102, loop body _rmsd.pyx:746-749):
// out[i] = sqrt( mean_k | b_i[k] - a[k] . R_i |^2 ), R_i = rot[i] (or its transpose), float32 terms summed in
// float64 like the reference.  One warp per frame.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rot_msd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      int64_t n_frames, int n_atoms, int64_t frame_stride,
                                                      const float* __restrict__ rot, int transpose,
                                                      float* __restrict__ rot_out, float* __restrict__ out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t)gridDim.x * 8;
    for (int64_t f = (int64_t)blockIdx.x * 8 + warp; f < n_frames; f += n_warps) {
        float R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(rot + f * 9 + (transpose ? (i % 3) * 3 + i / 3 : i));
        if (rot_out && lane < 9) rot_out[f * 9 + lane] = __ldg(rot + f * 9 + (transpose ? (lane % 3) * 3 + lane / 3 : lane));
        const float* bf = b + f * frame_stride;
        double acc = 0.0;
        for (int k = lane; k < n_atoms; k += 32) {
            const float ax = __ldg(a + 3 * k), ay = __ldg(a + 3 * k + 1), az = __ldg(a + 3 * k + 2);
            const float dx = __ldg(bf + 3 * k) - (ax * R[0] + ay * R[3] + az * R[6]);
            const float dy = __ldg(bf + 3 * k + 1) - (ax * R[1] + ay * R[4] + az * R[7]);
            const float dz = __ldg(bf + 3 * k + 2) - (ax * R[2] + ay * R[5] + az * R[8]);
            acc += (double)(dx * dx + dy * dy + dz * dz);
        }
        acc = warp_sum(acc);
        if (lane == 0) out[f] = sqrtf((float)(acc / (double)n_atoms));
    }
}

cudaError_t launch_rot_msd(const float* a, const float* b, int64_t n_frames, int n_atoms, int64_t frame_stride,
                           const float* rot, int transpose, float* rot_out, float* out, int sm_count, cudaStream_t st)
{
    if (n_frames <= 0) return cudaSuccess;
    int64_t ctas = (int64_t)sm_count * 8;
    const int64_t need = (n_frames + 7) / 8;
    if (ctas > need) ctas = need;
    rot_msd_kernel<<<(unsigned)ctas, 256, 0, st>>>(a, b, n_frames, n_atoms, frame_stride, rot, transpose, rot_out, out);
    return cudaGetLastError();
}

}  // namespace b200
