// frame_resident.cu -- single-pass read-modify-write kernels: every frame crosses HBM exactly once
// in each direction (24*N bytes per frame).
//
//   OP_SUPERPOSE  Trajectory.superpose (core/trajectory.py:1083-1173) =
//                   gather align subset (:1127) + float64 centroids and shifts (:1135-1147) +
//                   traces (:1149-1150) + superpose_atom_major (_rmsd.pyx:620-674: msd_atom_major with
//                   computeRot=1, rot_atom_major) + re-translation (:1171), fused.
//   OP_CENTER     inplace_center_and_trace_atom_major (center_sse.h:3-112): float64 sums, float32 mean,
//                 float32 subtraction in place, float64 trace of the float32 squares.
//
// One persistent CTA per SM = 16 compute warps + a loader warp + a storer warp.  A CTA owns a contiguous range of frames and
// moves them through a ring of `nbuf` whole-frame shared-memory buffers:
//     loader     : bulk load frame i (cp.async.bulk global->shared, full[i%nbuf] mbarrier) as soon as the storer
//                  reports the buffer drained.
//     storer     : once the compute group signals done[i%nbuf], bulk store the frame back (cp.async.bulk
//                  shared->global), wait for the store to have read the buffer, publish drained[i%nbuf].
//     compute    : the 16 warps form G independent groups of wpf warps; group g takes frames g, g+G, ...
//                  (sums -> float64 solve by the group's first thread -> transform in shared memory), with
//                  group-local named barriers only.  G frames are therefore in different phases at once
//                  and the serial QCP solve of one overlaps the streaming phases of the others; nbuf-G
//                  buffers are in flight to/from HBM.
// Neither the LSU nor L1 sits on the HBM path, and compute threads never wait for a store to drain.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "qcp.cuh"

namespace b200 {

// development aid: when set (B200RMSD_FUSED_TRACE=1), CTA 0 records clock64 stamps per frame:
// [i*8+0] data arrived, [1] sums reduced, [2] solve done, [3] transform done, [4] handed to storer,
// [5] store issued, [6] store drained, [7] load issued
__device__ long long* g_fr_trace = nullptr;
#define FR_STAMP(i, k) do { if (g_fr_trace && blockIdx.x == 0) g_fr_trace[(i) * 8 + (k)] = clock64(); } while (0)

constexpr int kFrWarps = 16;                      // compute warps
constexpr int kFrThreads = kFrWarps * 32 + 64;    // + loader warp + storer warp
constexpr int kPartStride = 17;

struct FrLayout {
    size_t buf_off, ref_off, idx_off, part_off, xf_off, bar_off, total;
    size_t buf_bytes;
};
__host__ __device__ inline size_t fr_align(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline FrLayout fr_layout(int n_pad, int nbuf, int n_sel_pad, int n_idx)
{
    FrLayout L;
    L.buf_bytes = (size_t)n_pad * 12;
    L.buf_off = 0;
    L.ref_off = fr_align((size_t)nbuf * L.buf_bytes, 128);
    L.idx_off = L.ref_off + fr_align((size_t)n_sel_pad * 12, 16);
    L.part_off = fr_align(L.idx_off + (size_t)n_idx * 4, 16);
    L.xf_off = L.part_off + (size_t)kFrWarps * kPartStride * sizeof(double);
    L.bar_off = fr_align(L.xf_off + (size_t)kFrWarps * 24 * sizeof(float), 8);
    L.total = L.bar_off + (size_t)(3 * nbuf + 1) * sizeof(uint64_t);
    return L;
}

__device__ __forceinline__ void group_sync(int group, int wpf)
{
    if (wpf == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(wpf * 32) : "memory");
}

__device__ __forceinline__ void xf_atom(float& x, float& y, float& z, const float* __restrict__ t)
{
    // t: R[0..8], c_hi[9..11], c_lo[12..14], o_hi[15..17], o_lo[18..20]
    const float tx = (x - t[9]) - t[12], ty = (y - t[10]) - t[13], tz = (z - t[11]) - t[14];
    const float rx = tx * t[0] + ty * t[3] + tz * t[6];
    const float ry = tx * t[1] + ty * t[4] + tz * t[7];
    const float rz = tx * t[2] + ty * t[5] + tz * t[8];
    x = (rx + t[15]) + t[18]; y = (ry + t[16]) + t[19]; z = (rz + t[17]) + t[20];
}

template <int OP>
__global__ void __launch_bounds__(kFrThreads, 1) frame_resident_kernel(const FusedParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int n_sel_pad = (p.n_sel + 3) & ~3;
    const FrLayout L = fr_layout(p.n_pad, p.nbuf, OP == OP_SUPERPOSE ? n_sel_pad : 0, p.idx ? p.n_sel : 0);
    const float* ref_s = reinterpret_cast<const float*>(smem + L.ref_off);
    int* idx_s = reinterpret_cast<int*>(smem + L.idx_off);
    double* part_s = reinterpret_cast<double*>(smem + L.part_off);
    float* xf_s = reinterpret_cast<float*>(smem + L.xf_off);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* done = full + p.nbuf;
    uint64_t* drained = done + p.nbuf;
    uint64_t* ref_bar = drained + p.nbuf;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = p.batch;                     // concurrent frame groups; nbuf % G == 0
    const int wpf = kFrWarps / G;              // warps per group (G * wpf <= 16 compute warps take part)
    const int n_cw = G * wpf;
    const int units = p.n_pad >> 2;
    const uint32_t frame_bytes = (uint32_t)p.n_pad * 12u;

    const int64_t f0 = p.n_frames * blockIdx.x / gridDim.x, f1 = p.n_frames * (blockIdx.x + 1) / gridDim.x;
    const int64_t n = f1 - f0;

    unsigned long long t_start_ns = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start_ns));
    if (tid == 0) {
        for (int i = 0; i < p.nbuf; ++i) { mbar_init(&full[i], 1); mbar_init(&done[i], 1); mbar_init(&drained[i], 1); }
        mbar_init(ref_bar, 1);
    }
    fence_mbar_init();
    __syncthreads();

    // ================================================================== DMA warps
    // warp 16 lane 0: loader.  warp 17 lane 0: storer.  Two threads so that waiting for a store to drain
    // (cp.async.bulk.wait_group.read is the only completion mechanism of shared->global bulk copies) never delays
    // the issue of a load into another buffer: the storer publishes drained[buf], the loader refills it at once.
    if (warp == kFrWarps) {
        if (lane == 0) {
            if (OP == OP_SUPERPOSE) {
                const uint32_t bytes = (uint32_t)n_sel_pad * 12u;
                mbar_arrive_expect_tx(ref_bar, bytes);
                bulk_g2s(smem + L.ref_off, p.ref, bytes, ref_bar);
            }
            for (int64_t i = 0; i < n; ++i) {
                const int buf = (int)(i % p.nbuf);
                if (i >= p.nbuf) mbar_wait(&drained[buf], (uint32_t)(((i / p.nbuf) - 1) & 1));
                mbar_arrive_expect_tx(&full[buf], frame_bytes);
                bulk_g2s(smem + L.buf_off + (size_t)buf * L.buf_bytes, p.xyz + (f0 + i) * p.frame_stride, frame_bytes,
                         &full[buf]);
                FR_STAMP(i, 7);
            }
        }
        return;
    }
    if (warp == kFrWarps + 1) {
        if (lane == 0) {
            // Up to K stores in flight: store i is issued, then the storer waits only until store i-K has drained and
            // publishes that buffer.  K = 0 when every buffer is busy computing or loading (nbuf == G); small frames
            // need K > 0 or the drain latency of each 4-12 KB store would serialise the whole pipeline.
            const int K = p.nbuf == G ? 0 : min(3, p.nbuf - G - 1);
            for (int64_t i = 0; i < n; ++i) {
                const int buf = (int)(i % p.nbuf);
                mbar_wait(&done[buf], (uint32_t)((i / p.nbuf) & 1));
                bulk_s2g(p.xyz + (f0 + i) * p.frame_stride, smem + L.buf_off + (size_t)buf * L.buf_bytes, frame_bytes);
                bulk_commit();
                FR_STAMP(i, 5);
                switch (K) {
                    case 0: bulk_wait_read<0>(); break;
                    case 1: bulk_wait_read<1>(); break;
                    case 2: bulk_wait_read<2>(); break;
                    default: bulk_wait_read<3>(); break;
                }
                if (i >= K) mbar_arrive(&drained[(int)((i - K) % p.nbuf)]);  // that buffer may be refilled
                FR_STAMP(i, 6);
            }
            bulk_wait_read<0>();
            for (int64_t i = max((int64_t)0, n - K); i < n; ++i) mbar_arrive(&drained[(int)(i % p.nbuf)]);
            bulk_wait<0>();
            if (g_fr_trace) {  // per-CTA wall time (ns) after the per-frame stamps of CTA 0
                unsigned long long t_end_ns;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end_ns));
                const long long base = (long long)(p.n_frames / gridDim.x + 2) * 8;
                g_fr_trace[base + blockIdx.x * 2] = (long long)t_start_ns;
                g_fr_trace[base + blockIdx.x * 2 + 1] = (long long)t_end_ns;
            }
        }
        return;
    }

    // ================================================================== compute warps
    if (warp >= n_cw) return;  // 16 is not a multiple of every G
    const int g = warp / wpf, sub = warp - g * wpf;
    const int gtid = sub * 32 + lane;  // thread index inside the group
    if (OP == OP_SUPERPOSE && p.idx)
        for (int k = tid; k < p.n_sel; k += n_cw * 32) idx_s[k] = __ldg(p.idx + k);
    float oh[3] = {0, 0, 0}, ol[3] = {0, 0, 0};
    RefStats rs{};
    if (OP == OP_SUPERPOSE) {
        rs = *p.ref_stats;
#pragma unroll
        for (int i = 0; i < 3; ++i) { oh[i] = (float)rs.mean[i]; ol[i] = (float)(rs.mean[i] - (double)oh[i]); }
        mbar_wait(ref_bar, 0);
    }
    asm volatile("bar.sync 0, %0;" ::"r"(n_cw * 32) : "memory");  // idx_s visible to all compute warps

#pragma unroll 1
    for (int64_t i = g; i < n; i += G) {
        const int buf = (int)(i % p.nbuf);
        // nbuf is a multiple of G, so buffer `buf` is only ever used by this group and the group observes every phase
        // of its barrier in order (a parity wait is ambiguous for a waiter that skips phases)
        mbar_wait(&full[buf], (uint32_t)((i / p.nbuf) & 1));
        if (gtid == 0) FR_STAMP(i, 0);
        float* frame_s = reinterpret_cast<float*>(smem + L.buf_off + (size_t)buf * L.buf_bytes);
        float4* xs = reinterpret_cast<float4*>(frame_s);
        const int64_t f = f0 + i;

        if (OP == OP_SUPERPOSE) {
            // ---- phase 1: sums over the align selection (pivot = first selected atom)
            float v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = 0.f;
            const int a0 = p.idx ? idx_s[0] : 0;
            const float px = frame_s[3 * a0], py = frame_s[3 * a0 + 1], pz = frame_s[3 * a0 + 2];
            if (p.idx) {
#pragma unroll 4
                for (int k = gtid; k < p.n_sel; k += wpf * 32) {
                    const int a = idx_s[k];
                    const float ax = frame_s[3 * a] - px, ay = frame_s[3 * a + 1] - py, az = frame_s[3 * a + 2] - pz;
                    const float bx = ref_s[3 * k], by = ref_s[3 * k + 1], bz = ref_s[3 * k + 2];
                    v[0] += ax; v[1] += ay; v[2] += az;
                    v[3] = fmaf(ax, ax, v[3]); v[3] = fmaf(ay, ay, v[3]); v[3] = fmaf(az, az, v[3]);
                    v[4] = fmaf(ax, bx, v[4]); v[5] = fmaf(ax, by, v[5]); v[6] = fmaf(ax, bz, v[6]);
                    v[7] = fmaf(ay, bx, v[7]); v[8] = fmaf(ay, by, v[8]); v[9] = fmaf(ay, bz, v[9]);
                    v[10] = fmaf(az, bx, v[10]); v[11] = fmaf(az, by, v[11]); v[12] = fmaf(az, bz, v[12]);
                }
            } else {
                const float4* ys = reinterpret_cast<const float4*>(ref_s);
                for (int u = gtid; u < units; u += wpf * 32) {
                    const float4 a0v = xs[3 * u], a1v = xs[3 * u + 1], a2v = xs[3 * u + 2];
                    const float4 b0v = ys[3 * u], b1v = ys[3 * u + 1], b2v = ys[3 * u + 2];
                    const int nvalid = p.n_atoms - 4 * u;
                    const float ax[4] = {a0v.x, a0v.w, a1v.z, a2v.y}, ay[4] = {a0v.y, a1v.x, a1v.w, a2v.z},
                                az[4] = {a0v.z, a1v.y, a2v.x, a2v.w};
                    const float bx[4] = {b0v.x, b0v.w, b1v.z, b2v.y}, by[4] = {b0v.y, b1v.x, b1v.w, b2v.z},
                                bz[4] = {b0v.z, b1v.y, b2v.x, b2v.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q < nvalid) {
                            const float dx = ax[q] - px, dy = ay[q] - py, dz = az[q] - pz;
                            v[0] += dx; v[1] += dy; v[2] += dz;
                            v[3] = fmaf(dx, dx, v[3]); v[3] = fmaf(dy, dy, v[3]); v[3] = fmaf(dz, dz, v[3]);
                            v[4] = fmaf(dx, bx[q], v[4]); v[5] = fmaf(dx, by[q], v[5]); v[6] = fmaf(dx, bz[q], v[6]);
                            v[7] = fmaf(dy, bx[q], v[7]); v[8] = fmaf(dy, by[q], v[8]); v[9] = fmaf(dy, bz[q], v[9]);
                            v[10] = fmaf(dz, bx[q], v[10]); v[11] = fmaf(dz, by[q], v[11]); v[12] = fmaf(dz, bz[q], v[12]);
                        }
                    }
                }
            }
            if (gtid == 0) { v[13] = px; v[14] = py; v[15] = pz; }
            warp_reduce_scatter16(v, lane);
            if (!(lane & 1)) part_s[warp * kPartStride + (lane >> 1)] = (double)v[0];
            group_sync(g, wpf);
            // ---- solve: the group's first thread
            if (gtid == 0) {
                FR_STAMP(i, 1);
                double rec[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) rec[q] = 0.0;
                for (int s = 0; s < wpf; ++s)
#pragma unroll
                    for (int q = 0; q < 16; ++q) rec[q] += part_s[(g * wpf + s) * kPartStride + q];
                const double invn = 1.0 / (double)p.n_sel;
                const double mx = rec[0] * invn, my = rec[1] * invn, mz = rec[2] * invn;
                QcpInput q;
                q.n_atoms = p.n_sel;
                q.Gb = rs.G;
                const double ga = rec[3] - (rec[0] * mx + rec[1] * my + rec[2] * mz);
                q.Ga = ga > 0.0 ? ga : 0.0;
                q.M[0] = rec[4] - mx * rs.sum[0];  q.M[1] = rec[5] - mx * rs.sum[1];  q.M[2] = rec[6] - mx * rs.sum[2];
                q.M[3] = rec[7] - my * rs.sum[0];  q.M[4] = rec[8] - my * rs.sum[1];  q.M[5] = rec[9] - my * rs.sum[2];
                q.M[6] = rec[10] - mz * rs.sum[0]; q.M[7] = rec[11] - mz * rs.sum[1]; q.M[8] = rec[12] - mz * rs.sum[2];
                float R[9];
                bool degen = false;
                const double msd = qcp_solve(q, R, &degen);
                if (p.out_rmsd) p.out_rmsd[f] = sqrtf((float)msd);
                if (p.out_rot) {
#pragma unroll
                    for (int c = 0; c < 9; ++c) p.out_rot[f * 9 + c] = R[c];
                }
                if (degen && p.degenerate) atomicAdd(p.degenerate, 1u);
                float* t = xf_s + g * 24;
#pragma unroll
                for (int c = 0; c < 9; ++c) t[c] = R[c];
                const double cen[3] = {rec[13] + mx, rec[14] + my, rec[15] + mz};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float h = (float)cen[c];
                    t[9 + c] = h; t[12 + c] = (float)(cen[c] - (double)h);
                    t[15 + c] = oh[c]; t[18 + c] = ol[c];
                }
                FR_STAMP(i, 2);
            }
            group_sync(g, wpf);
            // ---- phase 2: transform every atom of the frame in shared memory
            {
                float t[21];
#pragma unroll
                for (int c = 0; c < 21; ++c) t[c] = xf_s[g * 24 + c];
#pragma unroll 2
                for (int u = gtid; u < units; u += wpf * 32) {
                    float4 a0v = xs[3 * u], a1v = xs[3 * u + 1], a2v = xs[3 * u + 2];
                    const int nvalid = p.n_atoms - 4 * u;
                    xf_atom(a0v.x, a0v.y, a0v.z, t);
                    if (nvalid > 1) xf_atom(a0v.w, a1v.x, a1v.y, t);
                    if (nvalid > 2) xf_atom(a1v.z, a1v.w, a2v.x, t);
                    if (nvalid > 3) xf_atom(a2v.y, a2v.z, a2v.w, t);
                    xs[3 * u] = a0v; xs[3 * u + 1] = a1v; xs[3 * u + 2] = a2v;
                }
            }
        } else {  // ---------------------------------------------------------------- OP_CENTER
            double sx = 0, sy = 0, sz = 0;
            for (int u = gtid; u < units; u += wpf * 32) {
                const float4 a0v = xs[3 * u], a1v = xs[3 * u + 1], a2v = xs[3 * u + 2];
                sx += (double)a0v.x; sy += (double)a0v.y; sz += (double)a0v.z;
                sx += (double)a0v.w; sy += (double)a1v.x; sz += (double)a1v.y;
                sx += (double)a1v.z; sy += (double)a1v.w; sz += (double)a2v.x;
                sx += (double)a2v.y; sy += (double)a2v.z; sz += (double)a2v.w;
            }
            sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
            if (wpf > 1) {
                if (lane == 0) { part_s[warp * kPartStride] = sx; part_s[warp * kPartStride + 1] = sy; part_s[warp * kPartStride + 2] = sz; }
                group_sync(g, wpf);
                sx = sy = sz = 0;
                for (int s = 0; s < wpf; ++s) {
                    sx += part_s[(g * wpf + s) * kPartStride]; sy += part_s[(g * wpf + s) * kPartStride + 1];
                    sz += part_s[(g * wpf + s) * kPartStride + 2];
                }
            }
            const float mx = (float)(sx / p.n_atoms), my = (float)(sy / p.n_atoms), mz = (float)(sz / p.n_atoms);
            double tr = 0;
            for (int u = gtid; u < units; u += wpf * 32) {
                float4 a0v = xs[3 * u], a1v = xs[3 * u + 1], a2v = xs[3 * u + 2];
                const int nvalid = p.n_atoms - 4 * u;
                a0v.x -= mx; a0v.y -= my; a0v.z -= mz;
                tr += (double)(a0v.x * a0v.x); tr += (double)(a0v.y * a0v.y); tr += (double)(a0v.z * a0v.z);
                if (nvalid > 1) {
                    a0v.w -= mx; a1v.x -= my; a1v.y -= mz;
                    tr += (double)(a0v.w * a0v.w); tr += (double)(a1v.x * a1v.x); tr += (double)(a1v.y * a1v.y);
                }
                if (nvalid > 2) {
                    a1v.z -= mx; a1v.w -= my; a2v.x -= mz;
                    tr += (double)(a1v.z * a1v.z); tr += (double)(a1v.w * a1v.w); tr += (double)(a2v.x * a2v.x);
                }
                if (nvalid > 3) {
                    a2v.y -= mx; a2v.z -= my; a2v.w -= mz;
                    tr += (double)(a2v.y * a2v.y); tr += (double)(a2v.z * a2v.z); tr += (double)(a2v.w * a2v.w);
                }
                xs[3 * u] = a0v; xs[3 * u + 1] = a1v; xs[3 * u + 2] = a2v;
            }
            tr = warp_sum(tr);
            if (wpf > 1) {
                if (lane == 0) part_s[warp * kPartStride + 3] = tr;
                group_sync(g, wpf);
                if (gtid == 0) {
                    tr = 0;
                    for (int s = 0; s < wpf; ++s) tr += part_s[(g * wpf + s) * kPartStride + 3];
                }
            }
            if (gtid == 0 && p.traces) p.traces[f] = (float)tr;
        }
        // ---- publish the modified frame to the async proxy and hand the buffer to the DMA thread
        if (gtid == 0) FR_STAMP(i, 3);
        fence_proxy_async_smem();
        group_sync(g, wpf);
        if (gtid == 0) { mbar_arrive(&done[buf]); FR_STAMP(i, 4); }
    }
}

static bool fr_fits(const FusedParams& p, int op, int nbuf)
{
    const int n_sel_pad = (p.n_sel + 3) & ~3;
    return fr_layout(p.n_pad, nbuf, op == OP_SUPERPOSE ? n_sel_pad : 0, p.idx ? p.n_sel : 0).total <= 232448;
}

// Geometry: G frame groups (wpf = 16/G warps each) and a ring of nbuf = G*m buffers.  nbuf must be a multiple of G so
// that a buffer always belongs to one group (see the parity-wait note in the kernel).  Either every group has a second
// buffer to prefetch into (m >= 2) or there are >= 3 groups so that, while one computes, the others' single buffers
// are loading and storing (m == 1: the large-frame case, e.g. three 60 KB buffers at N = 5000).
bool fused_config(FusedParams& p, int op)
{
    const size_t frame_bytes = (size_t)p.n_pad * 12;
    if (frame_bytes >= (1u << 20)) return false;
    int nbuf_max = 0;
    while (nbuf_max < 48 && fr_fits(p, op, nbuf_max + 1)) ++nbuf_max;
    if (nbuf_max < 3) return false;  // cannot overlap load, compute and store
    static const int kGroups[] = {16, 8, 4, 3, 2, 1};
    const int g_cap = 8;  // measured at N = 1000: throughput grows with G up to 8 for both operations
    // first choice: m >= 2 with at least ~48 KB of prefetch in flight
    for (int G : kGroups) {
        if (G > g_cap || G < 2) continue;
        int m = nbuf_max / G;
        if (m < 2) continue;
        if ((size_t)(m - 1) * G * frame_bytes < 49152) continue;
        if (m > 4) m = 4;
        p.batch = G;
        p.nbuf = G * m;
        return true;
    }
    // large frames: one buffer per group, >= 3 groups
    for (int G : kGroups) {
        if (G <= nbuf_max && G >= 3 && G <= g_cap + 1) {
            p.batch = G;
            p.nbuf = G;
            return true;
        }
    }
    // last resort: one group of 16 warps with a 3-deep ring (measured 0.64x of HBM peak at N = 5000)
    p.batch = 1;
    p.nbuf = 3;
    return true;
}

// development override: G concurrent frame groups and ring depth (validated against the shared-memory budget)
bool fused_override(FusedParams& p, int op, int G, int nbuf)
{
    if (G <= 0) G = p.batch;
    if (nbuf <= 0) nbuf = p.nbuf;
    if (G > 16 || nbuf < G || nbuf % G != 0 || nbuf < 2) return false;
    if (!fr_fits(p, op, nbuf)) return false;
    p.batch = G;
    p.nbuf = nbuf;
    return true;
}

long long* fr_trace_buffer(size_t n_frames_cta)
{
    static long long* buf = nullptr;
    static size_t cap = 0;
    if (!getenv("B200RMSD_FUSED_TRACE")) return nullptr;
    if (cap < n_frames_cta * 8) {
        if (buf) cudaFree(buf);
        cudaMalloc((void**)&buf, n_frames_cta * 8 * sizeof(long long));
        cap = n_frames_cta * 8;
    }
    cudaMemset(buf, 0, cap * sizeof(long long));
    cudaMemcpyToSymbol(g_fr_trace, &buf, sizeof(buf));
    return buf;
}

extern "C" int b200rmsd_debug_fused_trace(long long* host_out, size_t n)
{
    long long* buf = nullptr;
    cudaMemcpyFromSymbol(&buf, g_fr_trace, sizeof(buf));
    if (!buf) return -1;
    cudaDeviceSynchronize();
    return cudaMemcpy(host_out, buf, n * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}

cudaError_t launch_frame_resident(const FusedParams& p, int op, int sm_count, cudaStream_t st)
{
    if (p.n_frames <= 0) return cudaSuccess;
    fr_trace_buffer((size_t)(p.n_frames / (sm_count > 0 ? sm_count : 1) + 2) + 64);
    const int n_sel_pad = (p.n_sel + 3) & ~3;
    const FrLayout L = fr_layout(p.n_pad, p.nbuf, op == OP_SUPERPOSE ? n_sel_pad : 0, p.idx ? p.n_sel : 0);
    auto kern = op == OP_SUPERPOSE ? frame_resident_kernel<OP_SUPERPOSE> : frame_resident_kernel<OP_CENTER>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) return e;
    int64_t ctas = sm_count;
    const int64_t need = (p.n_frames + p.batch - 1) / p.batch;
    if (ctas > need) ctas = need;
    kern<<<(unsigned)ctas, kFrThreads, L.total, st>>>(p);
    return cudaGetLastError();
}

}  // namespace b200
