// lprmsd.cu -- LP-RMSD: RMSD minimised over rotation/translation AND over the labels of exchangeable atoms
// (md.lprmsd, mdtraj/rmsd/_lprmsd.pyx:71-221; waters, identical ligands).  SURVEY.md section 8(f), last "next" row.
//
// The reference's three steps per frame (_lprmsd.pyx:186-224), one WARP per frame here, everything in shared memory:
//   1. centre the selected atoms (center.h:7 semantics: float64 mean, float32 subtraction, float64 trace); rotate them
//      onto the reference with the rotation that is optimal for the DISTINGUISHABLE atoms only (msd_atom_major with
//      computeRot on the subset, centred on its own centroid; rot_atom_major on the whole selection);
//   2. with that orientation fixed, the assignment problem of every permute group: cost[i][j] = |ref_i - target_j|^2
//      (euclidean_permutation.cpp:29-41: float32 differences and squares, float64 sums), minimum-cost perfect matching.
//      The reference runs Munkres on a dense n_sel x n_sel float64 matrix (O(n^3) with a large constant: 0.7 s per
//      frame at 300 atoms on this container's CPU); groups do not interact (cross-group entries are DBL_MAX), so each
//      group is solved on its own, by the shortest-augmenting-path form of the Hungarian method (Jonker-Volgenant
//      potentials) with the row scan spread over the 32 lanes and the costs recomputed from the coordinates instead of
//      stored -- O(g^2) shared memory traffic per augmentation, no matrix.  Any exact solver returns the same matching
//      whenever the optimum is unique;
//   3. the QCP RMSD of the relabelled selection (msd_atom_major again; with `out_rot` also rot1 . rot2, sgemm33 of
//      rotation.cpp:12-24, which `superpose=True` applies to the whole frame after centring ALL its atoms).
#include <cuda_runtime.h>

#include <cfloat>

#include "../../include/b200rmsd.h"
#include "common.cuh"
#include "kernels.cuh"
#include "qcp.cuh"

namespace b200 {
namespace {

struct LpParams {
    const float* xyz;
    int64_t n_frames, frame_stride;
    const int* idx;          // selection (sorted, unique) or nullptr: all atoms
    int n_sel;
    const float* ref_sel;    // (n_sel,3) reference conformation, selected atoms, NOT centred
    const int* dis;          // positions (inside the selection) of the distinguishable atoms
    int n_dis;
    const int* group_atoms;  // positions (inside the selection) of the permutable atoms, group after group
    const int* group_off;    // n_groups + 1 offsets into group_atoms
    int n_groups;
    int g_max;               // largest group
    float* out_rmsd;
    float* out_rot;          // (F,9) or nullptr
    int* out_map;            // (F,n_sel) or nullptr: mapping[i] = position of the target atom matched to reference atom i
    int warps;
};

// shared-memory layout: CTA-wide reference block, then one block per warp
__host__ __device__ inline size_t lp_align8(size_t x) { return (x + 7) & ~(size_t)7; }
__host__ __device__ inline size_t lp_cta_bytes(int n_sel, int n_dis)
{
    return lp_align8((size_t)n_sel * 12) + lp_align8((size_t)n_dis * 12) + 64;  // ref, ref_dis, {G_ref, G_ref_dis}
}
__host__ __device__ inline size_t lp_warp_bytes(int n_sel, int g_max)
{
    const size_t g = (size_t)g_max + 1;
    return lp_align8((size_t)n_sel * 12) + lp_align8((size_t)n_sel * 4) +  // target coordinates, mapping
           3 * g * 8 + 2 * lp_align8(g * 4) + lp_align8(g);                 // u, v, minv; p, way; used
}

struct ArgMin {
    double v;
    int j;
};
__device__ __forceinline__ ArgMin warp_argmin(double v, int j)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int j2 = __shfl_xor_sync(0xffffffffu, j, o);
        if (v2 < v || (v2 == v && j2 < j)) { v = v2; j = j2; }
    }
    return {v, j};
}

// centre n atoms of c (shared memory, (n,3)) like inplace_center_and_trace_atom_major (center_generic.h:3-44); returns
// the trace to every lane
__device__ __forceinline__ double lp_center(float* c, int n, int lane)
{
    double sx = 0, sy = 0, sz = 0;
    for (int k = lane; k < n; k += 32) { sx += (double)c[3 * k]; sy += (double)c[3 * k + 1]; sz += (double)c[3 * k + 2]; }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    const float mx = (float)(sx / n), my = (float)(sy / n), mz = (float)(sz / n);
    double tr = 0;
    for (int k = lane; k < n; k += 32) {
        const float x = c[3 * k] - mx, y = c[3 * k + 1] - my, z = c[3 * k + 2] - mz;
        c[3 * k] = x; c[3 * k + 1] = y; c[3 * k + 2] = z;
        tr += (double)(x * x); tr += (double)(y * y); tr += (double)(z * z);
    }
    __syncwarp();
    return warp_sum(tr);
}

__global__ void __launch_bounds__(256) lprmsd_kernel(const LpParams p)
{
    extern __shared__ __align__(16) unsigned char lp_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.n_sel, nd = p.n_dis;
    float* ref = reinterpret_cast<float*>(lp_smem);
    float* ref_dis = reinterpret_cast<float*>(lp_smem + lp_align8((size_t)n * 12));
    double* ref_g = reinterpret_cast<double*>(lp_smem + lp_align8((size_t)n * 12) + lp_align8((size_t)nd * 12));
    unsigned char* wbase = lp_smem + lp_cta_bytes(n, nd) + (size_t)warp * lp_warp_bytes(n, p.g_max);
    const size_t g1 = (size_t)p.g_max + 1;
    float* tgt = reinterpret_cast<float*>(wbase);
    int* mapping = reinterpret_cast<int*>(wbase + lp_align8((size_t)n * 12));
    double* u = reinterpret_cast<double*>(wbase + lp_align8((size_t)n * 12) + lp_align8((size_t)n * 4));
    double* v = u + g1;
    double* minv = v + g1;
    int* pcol = reinterpret_cast<int*>(minv + g1);
    int* way = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(pcol) + lp_align8(g1 * 4));
    unsigned char* used = reinterpret_cast<unsigned char*>(way) + lp_align8(g1 * 4);

    // ---- the reference, once per CTA (warp 0): whole selection centred, distinguishable subset centred on its own
    if (warp == 0) {
        for (int k = lane; k < 3 * n; k += 32) ref[k] = __ldg(p.ref_sel + k);
        for (int k = lane; k < nd; k += 32) {
            const int a = __ldg(p.dis + k);
            ref_dis[3 * k] = __ldg(p.ref_sel + 3 * a); ref_dis[3 * k + 1] = __ldg(p.ref_sel + 3 * a + 1);
            ref_dis[3 * k + 2] = __ldg(p.ref_sel + 3 * a + 2);
        }
        __syncwarp();
        const double g_all = lp_center(ref, n, lane);
        const double g_dis = nd > 0 ? lp_center(ref_dis, nd, lane) : 0.0;
        if (lane == 0) { ref_g[0] = g_all; ref_g[1] = g_dis; }
    }
    __syncthreads();
    const double G_ref = ref_g[0], G_ref_dis = ref_g[1];

    for (int64_t f = (int64_t)blockIdx.x * p.warps + warp; f < p.n_frames; f += (int64_t)gridDim.x * p.warps) {
        const float* fr = p.xyz + f * p.frame_stride;
        for (int k = lane; k < n; k += 32) {
            const int a = p.idx ? __ldg(p.idx + k) : k;
            tgt[3 * k] = __ldg(fr + 3 * a); tgt[3 * k + 1] = __ldg(fr + 3 * a + 1); tgt[3 * k + 2] = __ldg(fr + 3 * a + 2);
            mapping[k] = k;
        }
        __syncwarp();
        const double G_t = lp_center(tgt, n, lane);

        // ---- 1. rotation from the distinguishable atoms (_lprmsd.pyx:192-207)
        float rot1[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
        if (nd > 0) {
            double sx = 0, sy = 0, sz = 0;
            for (int k = lane; k < nd; k += 32) {
                const int a = __ldg(p.dis + k);
                sx += (double)tgt[3 * a]; sy += (double)tgt[3 * a + 1]; sz += (double)tgt[3 * a + 2];
            }
            sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
            const float mx = (float)(sx / nd), my = (float)(sy / nd), mz = (float)(sz / nd);
            double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, tr = 0;
            for (int k = lane; k < nd; k += 32) {
                const int a = __ldg(p.dis + k);
                const float x = tgt[3 * a] - mx, y = tgt[3 * a + 1] - my, z = tgt[3 * a + 2] - mz;
                tr += (double)(x * x); tr += (double)(y * y); tr += (double)(z * z);
                const double rx = ref_dis[3 * k], ry = ref_dis[3 * k + 1], rz = ref_dis[3 * k + 2];
                M[0] += x * rx; M[1] += x * ry; M[2] += x * rz;
                M[3] += y * rx; M[4] += y * ry; M[5] += y * rz;
                M[6] += z * rx; M[7] += z * ry; M[8] += z * rz;
            }
            QcpInput q;
#pragma unroll
            for (int i = 0; i < 9; ++i) q.M[i] = warp_sum(M[i]);
            q.Ga = warp_sum(tr);
            q.Gb = G_ref_dis;
            q.inv_n = 1.0 / (double)nd;
            qcp_solve(q, rot1, nullptr);  // every lane: the same inputs, the same rotation
            for (int k = lane; k < n; k += 32) {  // rot_atom_major: row vector x R
                const float x = tgt[3 * k], y = tgt[3 * k + 1], z = tgt[3 * k + 2];
                tgt[3 * k] = fmaf(z, rot1[6], fmaf(y, rot1[3], x * rot1[0]));
                tgt[3 * k + 1] = fmaf(z, rot1[7], fmaf(y, rot1[4], x * rot1[1]));
                tgt[3 * k + 2] = fmaf(z, rot1[8], fmaf(y, rot1[5], x * rot1[2]));
            }
            __syncwarp();
        }

        // ---- 2. assignment inside every permute group (rows = reference atoms, columns = target atoms)
        for (int gi = 0; gi < p.n_groups; ++gi) {
            const int* ga = p.group_atoms + __ldg(p.group_off + gi);
            const int g = __ldg(p.group_off + gi + 1) - __ldg(p.group_off + gi);
            if (g < 2) continue;
            for (int j = lane; j <= g; j += 32) { u[j] = 0.0; v[j] = 0.0; pcol[j] = 0; way[j] = 0; }
            __syncwarp();
            for (int i = 1; i <= g; ++i) {  // add row i (1-based) to the matching
                for (int j = lane; j <= g; j += 32) { minv[j] = DBL_MAX; used[j] = 0; }
                if (lane == 0) pcol[0] = i;
                __syncwarp();
                int j0 = 0;
                while (true) {
                    if (lane == 0) used[j0] = 1;
                    __syncwarp();
                    const int i0 = pcol[j0];
                    const int ra = __ldg(ga + i0 - 1);
                    const float rx = ref[3 * ra], ry = ref[3 * ra + 1], rz = ref[3 * ra + 2];
                    const double ui0 = u[i0];
                    double best = DBL_MAX;
                    int bj = 0x7fffffff;
                    for (int j = 1 + lane; j <= g; j += 32) {
                        if (used[j]) continue;
                        const int ta = __ldg(ga + j - 1);
                        const float dx = rx - tgt[3 * ta], dy = ry - tgt[3 * ta + 1], dz = rz - tgt[3 * ta + 2];
                        const double cost = (double)__fmul_rn(dx, dx) + (double)__fmul_rn(dy, dy) + (double)__fmul_rn(dz, dz);
                        const double cur = cost - ui0 - v[j];
                        double mj = minv[j];
                        if (cur < mj) { mj = cur; minv[j] = cur; way[j] = j0; }
                        if (mj < best) { best = mj; bj = j; }
                    }
                    __syncwarp();  // minv / way of column j were written by lane j - 1, the update below reads them on lane j
                    const ArgMin am = warp_argmin(best, bj);
                    const double delta = am.v;
                    for (int j = lane; j <= g; j += 32) {
                        if (used[j]) { u[pcol[j]] += delta; v[j] -= delta; }
                        else minv[j] -= delta;
                    }
                    __syncwarp();
                    j0 = am.j;
                    if (pcol[j0] == 0) break;
                }
                __syncwarp();     // every lane has read pcol[j0] above
                if (lane == 0) {  // flip the augmenting path
                    while (j0) {
                        const int j1 = way[j0];
                        pcol[j0] = pcol[j1];
                        j0 = j1;
                    }
                }
                __syncwarp();
            }
            for (int j = 1 + lane; j <= g; j += 32) mapping[__ldg(ga + pcol[j] - 1)] = __ldg(ga + j - 1);
            __syncwarp();
        }

        // ---- 3. QCP on the relabelled selection (_lprmsd.pyx:210-222)
        double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = lane; k < n; k += 32) {
            const int t = mapping[k];
            const double x = tgt[3 * t], y = tgt[3 * t + 1], z = tgt[3 * t + 2];
            const double rx = ref[3 * k], ry = ref[3 * k + 1], rz = ref[3 * k + 2];
            M[0] += x * rx; M[1] += x * ry; M[2] += x * rz;
            M[3] += y * rx; M[4] += y * ry; M[5] += y * rz;
            M[6] += z * rx; M[7] += z * ry; M[8] += z * rz;
        }
        QcpInput q;
#pragma unroll
        for (int i = 0; i < 9; ++i) q.M[i] = warp_sum(M[i]);
        q.Ga = G_t;
        q.Gb = G_ref;
        q.inv_n = 1.0 / (double)n;
        float rot2[9];
        const double msd = qcp_solve(q, p.out_rot ? rot2 : nullptr, nullptr);
        if (lane == 0) {
            p.out_rmsd[f] = sqrtf((float)msd);
            if (p.out_rot) {  // sgemm33(rot1, rot2), rotation.cpp:12-24
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        float o = 0.0f;
#pragma unroll
                        for (int k = 0; k < 3; ++k) o += rot1[3 * a + k] * rot2[3 * k + b];
                        p.out_rot[f * 9 + 3 * a + b] = o;
                    }
            }
        }
        if (p.out_map)
            for (int k = lane; k < n; k += 32) p.out_map[f * n + k] = mapping[k];
        __syncwarp();
    }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200rmsd_lprmsd_dev(const float* xyz, int64_t n_frames, int n_atoms, int64_t frame_stride,
                                   const int32_t* idx, int n_sel, const float* ref_sel, const int32_t* dis, int n_dis,
                                   const int32_t* group_atoms, const int32_t* group_off, int n_groups, int max_group,
                                   float* out_rmsd, float* out_rot, int32_t* out_map, void* stream)
{
    if (!xyz || !ref_sel || !out_rmsd || n_frames < 0 || n_atoms <= 0 || frame_stride < 3 * (int64_t)n_atoms || n_dis < 0 ||
        n_groups < 0 || max_group < 0 || (n_dis > 0 && !dis) || (n_groups > 0 && (!group_atoms || !group_off)))
        return set_error(B200RMSD_EINVAL, "lprmsd: bad arguments");
    const int n = idx ? n_sel : n_atoms;
    if (n <= 0 || n_dis > n || max_group > n) return set_error(B200RMSD_EINVAL, "lprmsd: bad selection sizes");
    if (n_frames == 0) return 0;
    int dev = 0, sm = 0, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t fixed = lp_cta_bytes(n, n_dis), per_warp = lp_warp_bytes(n, max_group);
    if (fixed + per_warp > (size_t)smem_max)
        return set_error(B200RMSD_EINVAL, "lprmsd: %d selected atoms (largest permute group %d) do not fit the %d KB of "
                         "shared memory a frame is solved in", n, max_group, smem_max >> 10);
    int warps = (int)(((size_t)smem_max - fixed) / per_warp);
    if (warps > 8) warps = 8;
    LpParams p{};
    p.xyz = xyz; p.n_frames = n_frames; p.frame_stride = frame_stride; p.idx = idx; p.n_sel = n;
    p.ref_sel = ref_sel; p.dis = dis; p.n_dis = n_dis; p.group_atoms = group_atoms; p.group_off = group_off;
    p.n_groups = n_groups; p.g_max = max_group; p.out_rmsd = out_rmsd; p.out_rot = out_rot; p.out_map = out_map;
    p.warps = warps;
    const size_t smem = fixed + (size_t)warps * per_warp;
    // as many CTAs as fit: the assignment is latency-bound (dependent shared-memory round trips), more warps hide it
    int per_sm = (int)((size_t)(228 * 1024 - 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t ctas = (int64_t)sm * per_sm;
    const int64_t need = (n_frames + warps - 1) / warps;
    if (ctas > need) ctas = need;
    cudaError_t e = cudaFuncSetAttribute(lprmsd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        lprmsd_kernel<<<(unsigned)ctas, warps * 32, smem, (cudaStream_t)stream>>>(p);
        e = cudaGetLastError();
    }
    return e == cudaSuccess ? 0 : set_error(B200RMSD_ECUDA, "lprmsd: %s", cudaGetErrorString(e));
}
