// allpairs.cu -- all-pairs RMSD matrix  D[i][j] = rmsd(frame j onto frame i).
//
// The reference has no such function: clustering code loops `md.rmsd(traj, traj, i)` over
// every i (examples/clustering.ipynb:78-81, centroids.ipynb:80-82), i.e. F one-vs-many
// passes, O(F^2 N) work behind O(F) Python calls.  Here it is one dense contraction
// (3F x A).(A x 3F) whose 3x3 blocks go straight into the QCP solve; the inner products
// are never written to HBM.
//
//   allpairs_prepare_kernel  centre every frame (selection applied) exactly like
//                            inplace_center_and_trace_atom_major (center_generic.h:3-44),
//                            store it axis-major (F,3,K) -- the layout of the reference's
//                            msd_axis_major (theobald_rmsd_generic.h:7-60), K-contiguous
//                            so it is a GEMM operand -- and its trace.
//   allpairs_simt_kernel     fp32 FMA tiles (32x32 frame pairs per CTA, 2x2 pairs per
//                            thread), float64 QCP epilogue.  Exact-arithmetic path: used
//                            for small problems and as the on-device accuracy check of the
//                            tensor-core path (allpairs_tc.cu).
#include "../../include/b200rmsd.h"
#include <cstring>

#include "allpairs_layout.cuh"
#include "common.cuh"
#include "kernels.cuh"
#include "qcp.cuh"

namespace b200 {

int g_ap_min_tc_frames = 512;
int g_ap_max_refs = kApMaxRefs;
int g_ap_cta_pair = 1;

// one warp per frame
__global__ void __launch_bounds__(256) allpairs_prepare_kernel(const float* __restrict__ xyz, int64_t n_frames,
                                                               int64_t frame_stride, const int* __restrict__ idx,
                                                               int n_sel, int k_pad, float* __restrict__ X,
                                                               float* __restrict__ traces)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t)gridDim.x * 8;
    for (int64_t f = (int64_t)blockIdx.x * 8 + warp; f < n_frames; f += n_warps) {
        const float* fr = xyz + f * frame_stride;
        double sx = 0, sy = 0, sz = 0;
        for (int k = lane; k < n_sel; k += 32) {
            const int a = idx ? __ldg(idx + k) : k;
            sx += (double)__ldg(fr + 3 * a); sy += (double)__ldg(fr + 3 * a + 1); sz += (double)__ldg(fr + 3 * a + 2);
        }
        sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
        const float mx = (float)(sx / n_sel), my = (float)(sy / n_sel), mz = (float)(sz / n_sel);
        float* out = X + (size_t)f * 3 * k_pad;
        double tr = 0;
        for (int k = lane; k < k_pad; k += 32) {
            float x = 0.f, y = 0.f, z = 0.f;
            if (k < n_sel) {
                const int a = idx ? __ldg(idx + k) : k;
                x = __ldg(fr + 3 * a) - mx; y = __ldg(fr + 3 * a + 1) - my; z = __ldg(fr + 3 * a + 2) - mz;
                tr += (double)(x * x); tr += (double)(y * y); tr += (double)(z * z);
            }
            out[k] = x; out[k_pad + k] = y; out[2 * k_pad + k] = z;
        }
        tr = warp_sum(tr);
        if (lane == 0) traces[f] = (float)tr;
    }
}

// ---------------------------------------------------------------------------
// SIMT tile kernel.  CTA = 256 threads = 16 (ti) x 16 (tj); thread owns i-frames
// {ti, ti+16} and j-frames {tj, tj+16} of a 32x32 tile.
// smem per operand: 32 frames x 3 comps x (32+4) floats (row pad keeps 128-bit loads
// conflict-free: frame stride 108 floats = 12 banks mod 32).
// ---------------------------------------------------------------------------
constexpr int kTile = 32, kKc = 32, kRow = kKc + 4;

__global__ void __launch_bounds__(256) allpairs_simt_kernel(const float* __restrict__ X, const float* __restrict__ traces,
                                                            int64_t n_frames, int n_sel, int k_pad, int64_t row0,
                                                            int64_t row1, int64_t col0, int64_t col1,
                                                            float* __restrict__ out, int64_t ld, float* __restrict__ out_t,
                                                            int64_t ld_t, float* __restrict__ out_rot, unsigned flags)
{
    __shared__ __align__(16) float As[kTile * 3 * kRow];
    __shared__ __align__(16) float Bs[kTile * 3 * kRow];
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    const int64_t i0 = row0 + (int64_t)blockIdx.y * kTile, j0 = col0 + (int64_t)blockIdx.x * kTile;

    float acc[2][2][9];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 9; ++c) acc[a][b][c] = 0.f;

    for (int k0 = 0; k0 < k_pad; k0 += kKc) {
        // 96 rows x 8 float4 per operand, 3 float4 per thread
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int q = tid + r * 256;          // 0..767
            const int row = q >> 3, c4 = q & 7;   // row = frame*3 + comp
            const int fl = row / 3, comp = row - fl * 3;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            const int64_t fi = i0 + fl, fj = j0 + fl;
            if (fi < row1) va = *reinterpret_cast<const float4*>(X + ((size_t)fi * 3 + comp) * k_pad + k0 + c4 * 4);
            if (fj < col1) vb = *reinterpret_cast<const float4*>(X + ((size_t)fj * 3 + comp) * k_pad + k0 + c4 * 4);
            *reinterpret_cast<float4*>(&As[row * kRow + c4 * 4]) = va;
            *reinterpret_cast<float4*>(&Bs[row * kRow + c4 * 4]) = vb;
        }
        __syncthreads();
#pragma unroll 2
        for (int kk = 0; kk < kKc; kk += 4) {
            float4 a[2][3], b[2][3];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    a[s][c] = *reinterpret_cast<const float4*>(&As[((ti + 16 * s) * 3 + c) * kRow + kk]);
                    b[s][c] = *reinterpret_cast<const float4*>(&Bs[((tj + 16 * s) * 3 + c) * kRow + kk]);
                }
            // M[3*p+q] = sum_k a_k[p] * b_k[q] with a = frame j (column), b = frame i (row): D[i][j] = rmsd(j onto i)
#pragma unroll
            for (int si = 0; si < 2; ++si)
#pragma unroll
                for (int sj = 0; sj < 2; ++sj)
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            float m = acc[si][sj][3 * p + q];
                            m = fmaf(b[sj][p].x, a[si][q].x, m);
                            m = fmaf(b[sj][p].y, a[si][q].y, m);
                            m = fmaf(b[sj][p].z, a[si][q].z, m);
                            m = fmaf(b[sj][p].w, a[si][q].w, m);
                            acc[si][sj][3 * p + q] = m;
                        }
        }
        __syncthreads();
    }

    const double inv_n = 1.0 / (double)n_sel;
#pragma unroll
    for (int si = 0; si < 2; ++si)
#pragma unroll
        for (int sj = 0; sj < 2; ++sj) {
            const int64_t i = i0 + ti + 16 * si, j = j0 + tj + 16 * sj;
            if (i >= row1 || j >= col1) continue;
            float r;
            // optional rotation of frame j onto frame i (row vector x R, rotation_generic.h:40-42), 9 floats per pair
            float* rot = out_rot ? out_rot + ((size_t)(i - row0) * (size_t)(col1 - col0) + (size_t)(j - col0)) * 9 : nullptr;
            if (i == j && (flags & 1u)) {
                r = 0.f;  // a frame against itself in the same memory: theobald_rmsd_sse.h:256-262
                if (rot) {
#pragma unroll
                    for (int c = 0; c < 9; ++c) rot[c] = (c % 4 == 0) ? 1.0f : 0.0f;
                }
            } else {
                QcpInput q;
                q.inv_n = inv_n;
                q.Ga = (double)traces[j];
                q.Gb = (double)traces[i];
#pragma unroll
                for (int c = 0; c < 9; ++c) q.M[c] = (double)acc[si][sj][c];
                r = sqrtf((float)qcp_solve(q, rot, nullptr));
            }
            out[(size_t)(i - row0) * ld + j] = r;
            if (out_t) out_t[(size_t)(j - col0) * ld_t + (i - row0)] = r;
        }
}

}  // namespace b200

using namespace b200;

#define fail b200::set_error

static int ap_sm_count()
{
    int dev = 0, sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    return sm;
}

extern "C" {

int b200rmsd_allpairs_configure(int min_tc_frames, int max_refs)
{
    if (min_tc_frames > 0) g_ap_min_tc_frames = min_tc_frames;
    if (max_refs > 0) g_ap_max_refs = max_refs < kApMaxRefs ? max_refs : kApMaxRefs;
    return 0;
}

int b200rmsd_allpairs_set_cta_pair(int cta_pair)
{
    const int was = g_ap_cta_pair;
    g_ap_cta_pair = cta_pair != 0;
    return was;
}

// ---- peer-visible buffers (multi-process all-pairs: transposed blocks are written into the owner's row block over NVLink)
int b200rmsd_peer_alloc(size_t bytes, void** dev_ptr, void* handle)
{
    if (!dev_ptr || !handle || bytes == 0) return fail(B200RMSD_EINVAL, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == B200RMSD_PEER_HANDLE_BYTES, "handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(B200RMSD_ECUDA, "peer_alloc: %s", cudaGetErrorString(e));
    e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle), p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(B200RMSD_ECUDA, "peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    *dev_ptr = p;
    return 0;
}

int b200rmsd_peer_open(const void* handle, void** dev_ptr)
{
    if (!dev_ptr || !handle) return fail(B200RMSD_EINVAL, "peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(B200RMSD_ECUDA, "peer_open: %s", cudaGetErrorString(e));
    *dev_ptr = p;
    return 0;
}

int b200rmsd_peer_close(void* dev_ptr)
{
    const cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    return e == cudaSuccess ? 0 : fail(B200RMSD_ECUDA, "peer_close: %s", cudaGetErrorString(e));
}

int b200rmsd_peer_free(void* dev_ptr)
{
    const cudaError_t e = cudaFree(dev_ptr);
    return e == cudaSuccess ? 0 : fail(B200RMSD_ECUDA, "peer_free: %s", cudaGetErrorString(e));
}

int b200rmsd_allpairs_info_dev(const void* workspace, size_t workspace_bytes, int* n_refs, int* n_far, float* cover_radius,
                               int* ref_frames, int ref_frames_cap, void* stream)
{
    if (!workspace || workspace_bytes < 256) return fail(B200RMSD_EINVAL, "allpairs_info: bad arguments");
    ApHeader h;
    cudaError_t e = cudaMemcpyAsync(&h, workspace, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return fail(B200RMSD_ECUDA, "allpairs_info: %s", cudaGetErrorString(e));
    if (n_refs) *n_refs = h.n_refs;
    if (n_far) *n_far = h.n_far;
    if (cover_radius) *cover_radius = h.cover_radius;
    for (int i = 0; ref_frames && i < ref_frames_cap && i < h.n_refs; ++i) ref_frames[i] = h.ref_frame[i];
    return 0;
}

size_t b200rmsd_allpairs_workspace_bytes(int64_t n_frames, int n_sel)
{
    if (n_frames <= 0 || n_sel <= 0) return 256;
    return ap_geometry(n_frames, n_sel).total;
}

int b200rmsd_allpairs_prepare_dev(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride,
                                  const int32_t* idx, int n_sel, void* workspace, size_t workspace_bytes, void* stream)
{
    if (!xyz || !workspace || n_frames <= 0 || n_atoms <= 0) return fail(B200RMSD_EINVAL, "allpairs_prepare: bad arguments");
    const int ns = idx ? n_sel : n_atoms;
    if (ns <= 0) return fail(B200RMSD_EINVAL, "allpairs_prepare: empty selection");
    const ApGeometry g = ap_geometry(n_frames, ns);
    if (workspace_bytes < g.total) return fail(B200RMSD_EINVAL, "allpairs_prepare: workspace too small (need %zu bytes)", g.total);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return fail(B200RMSD_EINVAL, "allpairs_prepare: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)workspace;
    const int sm = ap_sm_count();
    cudaError_t e;
    if (g.tc) {
        // Choose the reference structures and align every frame onto its nearest one: the RMSD of a pair does not change
        // under a rigid motion of either frame, and a frame differs from a reference of its own basin by fluctuations
        // only -- which is what keeps the tensor-core accumulators small (allpairs_refs.cu, allpairs_tc144_prepare_kernel).
        int n_refs = 0;
        if (int rc = ap_select_references(xyz, n_frames, n_atoms, frame_stride, idx, ns, g, base, g_ap_max_refs, sm, st, &n_refs))
            return rc;
        e = launch_allpairs_tc144_prepare(xyz, n_frames, frame_stride, idx, ns, g, base, n_refs, sm, st);
    } else {
        int64_t ctas = (int64_t)sm * 8;
        const int64_t need = (n_frames + 7) / 8;
        if (ctas > need) ctas = need;
        e = cudaMemsetAsync(base, 0, 256, st);  // ApHeader: no references on this path
        if (e == cudaSuccess) {
            allpairs_prepare_kernel<<<(unsigned)ctas, 256, 0, st>>>(xyz, n_frames, frame_stride, idx, ns, g.k_pad,
                                                                    (float*)(base + g.x_off), (float*)(base + g.traces_off));
            e = cudaGetLastError();
        }
    }
    return e == cudaSuccess ? 0 : fail(B200RMSD_ECUDA, "allpairs_prepare: %s", cudaGetErrorString(e));
}

static int ap_block(const void* workspace, size_t workspace_bytes, int64_t n_frames, int n_sel, int64_t row0, int64_t row1,
                    int64_t col0, int64_t col1, float* out, int64_t ld, float* out_t, int64_t ld_t, float* out_rot,
                    unsigned flags, void* stream)
{
    if (!workspace || !out || n_frames <= 0 || n_sel <= 0 || row0 < 0 || row1 > n_frames || row0 > row1 || col0 < 0 ||
        col1 > n_frames || col0 > col1 || ld < col1 - col0 || (out_t && ld_t < row1 - row0))
        return fail(B200RMSD_EINVAL, "allpairs_block: bad arguments");
    const ApGeometry g = ap_geometry(n_frames, n_sel);
    if (workspace_bytes < g.total) return fail(B200RMSD_EINVAL, "allpairs_block: workspace too small");
    if (row0 == row1 || col0 == col1) return 0;
    const char* base = (const char*)workspace;
    if (g.tc)
        return launch_allpairs_tc144_block(g, base, n_sel, n_frames, row0, row1, col0, col1, out, ld, out_t, ld_t, out_rot,
                                           flags, ap_sm_count(), (cudaStream_t)stream);
    dim3 grid((unsigned)((col1 - col0 + kTile - 1) / kTile), (unsigned)((row1 - row0 + kTile - 1) / kTile));
    if (grid.y > 65535) return fail(B200RMSD_EINVAL, "allpairs_block: at most 65535*32 rows per call");
    allpairs_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)(base + g.x_off),
                                                                 (const float*)(base + g.traces_off), n_frames, n_sel,
                                                                 g.k_pad, row0, row1, col0, col1, out, ld, out_t, ld_t,
                                                                 out_rot, flags);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail(B200RMSD_ECUDA, "allpairs_block: %s", cudaGetErrorString(e));
}

int b200rmsd_allpairs_block_dev(const void* workspace, size_t workspace_bytes, int64_t n_frames, int n_sel,
                                int64_t row0, int64_t row1, int64_t col0, int64_t col1, float* out, int64_t ld,
                                float* out_t, int64_t ld_t, unsigned flags, void* stream)
{
    return ap_block(workspace, workspace_bytes, n_frames, n_sel, row0, row1, col0, col1, out, ld, out_t, ld_t, nullptr,
                    flags, stream);
}

int b200rmsd_allpairs_block_rot_dev(const void* workspace, size_t workspace_bytes, int64_t n_frames, int n_sel,
                                    int64_t row0, int64_t row1, int64_t col0, int64_t col1, float* out, int64_t ld,
                                    float* out_rot, unsigned flags, void* stream)
{
    if (!out_rot) return fail(B200RMSD_EINVAL, "allpairs_block_rot: out_rot is NULL");
    return ap_block(workspace, workspace_bytes, n_frames, n_sel, row0, row1, col0, col1, out, ld, nullptr, 0, out_rot,
                    flags, stream);
}

int b200rmsd_allpairs_rows_dev(const void* workspace, size_t workspace_bytes, int64_t n_frames, int n_sel,
                               int64_t row0, int64_t row1, float* out, int64_t ld, unsigned flags, void* stream)
{
    if (ld < n_frames) return fail(B200RMSD_EINVAL, "allpairs_rows: ld < n_frames");
    return b200rmsd_allpairs_block_dev(workspace, workspace_bytes, n_frames, n_sel, row0, row1, 0, n_frames, out, ld,
                                       nullptr, 0, flags, stream);
}

}  // extern "C"
