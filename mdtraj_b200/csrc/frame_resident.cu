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
// One persistent CTA per SM = 16 compute warps + a loader warp + a storer warp.  A CTA owns a contiguous range of frames,
// cuts it into *slots* of `fpb` consecutive frames (fpb > 1 only for frames that are contiguous in memory; one bulk copy
// then moves fpb frames, which is what keeps small frames off the per-copy issue limit) and moves the slots through a
// ring of `nbuf` shared-memory buffers:
//     loader     : bulk load slot s (cp.async.bulk global->shared, full[s%nbuf] mbarrier) as soon as the storer
//                  reports the buffer drained.
//     storer     : once the compute group signals done[s%nbuf], bulk store the slot back (cp.async.bulk
//                  shared->global), wait for the store to have read the buffer, publish drained[s%nbuf].
//     compute    : the 16 warps form G independent groups of wpf warps; group g takes slots g, g+G, ...
//                  (sums of every frame of the slot -> float64 solves, one lane per frame, side by side -> transform
//                  in shared memory), with group-local named barriers only.  Inside a group the warps either share
//                  each frame (team_warps == wpf) or take whole frames each (team_warps == 1, fpb >= wpf).
//                  G slots are in different phases at once, so the serial QCP solve latency of one overlaps the
//                  streaming phases of the others; nbuf-G buffers are in flight to/from HBM.
// Neither the LSU nor L1 sits on the HBM path, and compute threads never wait for a store to drain.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "qcp.cuh"

namespace b200 {

// development aid: when set (B200RMSD_FUSED_TRACE=1), CTA 0 records clock64 stamps per slot:
// [i*8+0] data arrived, [1] sums reduced, [2] solve done, [3] transform done, [4] handed to storer,
// [5] store issued, [6] store drained, [7] load issued
__device__ long long* g_fr_trace = nullptr;
#define FR_STAMP(i, k) do { if (g_fr_trace && blockIdx.x == 0) g_fr_trace[(i) * 8 + (k)] = clock64(); } while (0)

extern int g_fr_pipe_min_bytes;
constexpr int kFrWarps = 16;                      // compute warps
constexpr int kFrThreads = kFrWarps * 32 + 64;    // + loader warp + storer warp
constexpr int kRecStride = 25;  // floats per frame record (odd: the solver lanes read one record each, conflict-free)

struct FrLayout {
    size_t buf_off, ref_off, idx_off, rec_off, dsum_off, cpart_off, bar_off, rs_off, total;
    size_t buf_bytes;
};
__host__ __device__ inline size_t fr_align(size_t x, size_t a) { return (x + a - 1) / a * a; }
// rec: one record per (group, frame of the slot, warp of the team): first the 16 float32 partial sums, then -- written by
// the solver lane of that frame into the team's first record -- the 15 floats of its transform;
// dsum: when several warps share a frame, their partials combined in float64 (16 per frame, by 16 lanes in parallel);
// cpart: float64 partials of OP_CENTER, 4 per compute warp
__host__ __device__ inline FrLayout fr_layout(int n_pad, int nbuf, int n_sel_pad, int n_idx, int G, int fpb, int team_warps,
                                              bool records)
{
    FrLayout L;
    L.buf_bytes = (size_t)n_pad * 12 * fpb;
    L.buf_off = 0;
    L.ref_off = fr_align((size_t)nbuf * L.buf_bytes, 128);
    L.idx_off = L.ref_off + fr_align((size_t)n_sel_pad * 12, 16);
    L.rec_off = fr_align(L.idx_off + (size_t)n_idx * 4, 16);
    L.dsum_off = fr_align(L.rec_off + (records ? (size_t)G * fpb * team_warps * kRecStride * sizeof(float) : 0), 8);
    L.cpart_off = L.dsum_off + (records && team_warps > 1 ? (size_t)G * fpb * 16 * sizeof(double) : 0);
    L.bar_off = L.cpart_off + (size_t)kFrWarps * 4 * sizeof(double);
    L.rs_off = L.bar_off + (size_t)(3 * nbuf + 1) * sizeof(uint64_t);  // RefStats copy (OP_SUPERPOSE)
    L.total = L.rs_off + 64;
    return L;
}

__device__ __forceinline__ void group_sync(int group, int wpf)
{
    if (wpf == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(wpf * 32) : "memory");
}

// x' = (x - c) R + o with the float64 centroid c = c_hi + c_lo and the float64 reference mean o (the reference does both
// broadcasts in float64 numpy, trajectory.py:1140,1171) as 12 float32 operations per atom:
//   t = x - c_hi  (a difference of nearby float32 numbers: at most half an ulp of the centred coordinate),
//   x' = t R + d,  d = o - c_lo R  evaluated in float64 by the solver lane, added first in the FMA chain.
// Worst case ~2 ulp of the result (1e-6 nm at |x'| = 5 nm) against the reference's ~0.5 ulp + 1.5 ulp of t.
__device__ __forceinline__ void xf_atom(float& x, float& y, float& z, const float* __restrict__ t)
{
    // t: R[0..8], c_hi[9..11], d[12..14]
    const float tx = x - t[9], ty = y - t[10], tz = z - t[11];
    x = fmaf(tx, t[0], fmaf(ty, t[3], fmaf(tz, t[6], t[12])));
    y = fmaf(tx, t[1], fmaf(ty, t[4], fmaf(tz, t[7], t[13])));
    z = fmaf(tx, t[2], fmaf(ty, t[5], fmaf(tz, t[8], t[14])));
}

// reduce the 16 lane partials over the L lanes that share a frame and drop them into the frame's record
template <int L>
__device__ __forceinline__ void team_reduce_store(float (&v)[16], int lane, bool act, float* rec)
{
    const int base = group_reduce_scatter16<L>(v, lane);
    if (act) {
#pragma unroll
        for (int k = 0; k < 16 / L; ++k) rec[base + k] = v[k];
    }
}

__device__ __forceinline__ double lanes_sum(double x, int L)  // all-reduce over aligned groups of L lanes
{
    for (int h = L >> 1; h >= 1; h >>= 1) x += __shfl_xor_sync(0xffffffffu, x, h);
    return x;
}


// One frame of a slot: sums record -> RMSD, rotation and the transform record of xf_atom, by one lane.
// rec_f: the frame's float32 record (sums 0..12, pivot 13..15) or nullptr when rec_d holds the float64 sums of several
// warps; t: where the 15 floats of the transform go -- the frame's own (first) record, so everything needed from the
// sums is read before the first write.  The centroid and the pivot are read a second time after the solve instead of
// being carried across it, and the reference's statistics stay in shared memory: the QCP solve needs ~90 registers of
// float64 on its own, and what it pushed to local memory was re-read at L2 latency in the middle of a 6700-cycle chain
// (tools/fused_trace.py; ncu: local loads with a 48 % L1 hit rate).
__device__ __forceinline__ void fr_solve_frame(const float* rec_f, const double* rec_d, float* t, const RefStats* rs_s,
                                               double invn, int64_t f, const FusedParams& p)
{
    auto sum = [&](int q) -> double { return rec_f ? (double)rec_f[q] : rec_d[q]; };
    float R[9];
    {
        const double s0 = sum(0), s1 = sum(1), s2 = sum(2);
        const double mx = s0 * invn, my = s1 * invn, mz = s2 * invn;
        const double r0 = rs_s->sum[0], r1 = rs_s->sum[1], r2 = rs_s->sum[2];
        QcpInput q;
        q.inv_n = invn;
        q.Gb = rs_s->G;
        const double ga = sum(3) - (s0 * mx + s1 * my + s2 * mz);
        q.Ga = ga > 0.0 ? ga : 0.0;
        q.M[0] = sum(4) - mx * r0;  q.M[1] = sum(5) - mx * r1;  q.M[2] = sum(6) - mx * r2;
        q.M[3] = sum(7) - my * r0;  q.M[4] = sum(8) - my * r1;  q.M[5] = sum(9) - my * r2;
        q.M[6] = sum(10) - mz * r0; q.M[7] = sum(11) - mz * r1; q.M[8] = sum(12) - mz * r2;
        bool degen = false;
        const double msd = qcp_solve(q, R, &degen);
        if (p.out_rmsd) p.out_rmsd[f] = sqrtf((float)msd);
        if (degen && p.degenerate) atomicAdd(p.degenerate, 1u);
    }
    if (p.out_rot) {
#pragma unroll
        for (int c = 0; c < 9; ++c) p.out_rot[f * 9 + c] = R[c];
    }
    // centroid = pivot + mean shift, split into a float32 part (subtracted per atom) and a float64 remainder
    double clo[3];
    float chi[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double cen = sum(13 + c) + sum(c) * invn;
        chi[c] = (float)cen;
        clo[c] = cen - (double)chi[c];
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) t[c] = R[c];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        t[9 + c] = chi[c];
        t[12 + c] = (float)(rs_s->mean[c] - (clo[0] * (double)R[c] + clo[1] * (double)R[3 + c] + clo[2] * (double)R[6 + c]));
    }
}

template <int OP>
__global__ void __launch_bounds__(kFrThreads, 1) frame_resident_kernel(const FusedParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int n_sel_pad = (p.n_sel + 3) & ~3;
    const int G = p.batch;                     // concurrent slot groups; nbuf % G == 0
    const int fpb = p.fpb;                     // frames per slot
    const int wpf = kFrWarps / G;              // warps per group (G * wpf <= 16 compute warps take part)
    const int tw = p.team_warps;               // warps sharing one frame: wpf (whole group) or 1
    const FrLayout L = fr_layout(p.n_pad, p.nbuf, OP == OP_SUPERPOSE ? n_sel_pad : 0, p.idx ? p.n_sel : 0, G, fpb, tw,
                                 OP == OP_SUPERPOSE);
    const float* ref_s = reinterpret_cast<const float*>(smem + L.ref_off);
    int* idx_s = reinterpret_cast<int*>(smem + L.idx_off);
    float* rec_all = reinterpret_cast<float*>(smem + L.rec_off);
    double* cpart_s = reinterpret_cast<double*>(smem + L.cpart_off);
    double* dsum_s = reinterpret_cast<double*>(smem + L.dsum_off);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* done = full + p.nbuf;
    uint64_t* drained = done + p.nbuf;
    uint64_t* ref_bar = drained + p.nbuf;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_cw = G * wpf;
    const int units = p.n_pad >> 2;
    const uint32_t frame_bytes = (uint32_t)p.n_pad * 12u;
    const int frame_floats = p.n_pad * 3;

    const int64_t f0 = p.n_frames * blockIdx.x / gridDim.x, f1 = p.n_frames * (blockIdx.x + 1) / gridDim.x;
    const int64_t n = f1 - f0;
    const int64_t n_slots = (n + fpb - 1) / fpb;

    unsigned long long t_start_ns = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start_ns));
    RefStats* rs_s = reinterpret_cast<RefStats*>(smem + L.rs_off);
    if (tid == 0) {
        for (int i = 0; i < p.nbuf; ++i) { mbar_init(&full[i], 1); mbar_init(&done[i], 1); mbar_init(&drained[i], 1); }
        mbar_init(ref_bar, 1);
        if (OP == OP_SUPERPOSE) *rs_s = *p.ref_stats;
    }
    fence_mbar_init();
    __syncthreads();

    // ================================================================== DMA warps
    // warp 16 lane 0: loader.  warp 17 lane 0: storer.  Two threads so that waiting for a store to drain
    // (cp.async.bulk.wait_group.read is the only completion mechanism of shared->global bulk copies) never delays
    // the issue of a load into another buffer: the storer publishes drained[buf], the loader refills it at once.
    // Superpose with one buffer per group on 8 or more groups (frames up to ~25 KB): the group's own first
    // thread stores its slot, waits for the store to have read the buffer and loads the group's next slot into it.  The
    // group can do nothing else in between anyway, and the two DMA threads' queues drop out of the cycle: with eight
    // groups handing over at random times a slot waited 2400 cycles for the storer thread, which blocks ~1100 cycles
    // per store in cp.async.bulk.wait_group.read (12 % of the 20 000-cycle cycle of a group, tools/fused_trace.py).
    // Measured (gpurun r02_fused_sweep vs r02_fused_sweep2): +0.01-0.025 of HBM peak there; centring and the 3-5 group
    // geometries of larger frames lose up to 0.15 without the DMA threads, so they keep them.
    const bool self_dma = OP == OP_SUPERPOSE && p.nbuf == G && G >= 8;
    if (warp == kFrWarps) {
        if (lane == 0) {
            if (OP == OP_SUPERPOSE) {
                const uint32_t bytes = (uint32_t)n_sel_pad * 12u;
                mbar_arrive_expect_tx(ref_bar, bytes);
                bulk_g2s(smem + L.ref_off, p.ref, bytes, ref_bar);
            }
            for (int64_t s = 0; s < (self_dma ? (n_slots < G ? n_slots : G) : n_slots); ++s) {  // self_dma: first round only
                const int buf = (int)(s % p.nbuf);
                const int64_t left = n - s * fpb;
                const uint32_t bytes = (uint32_t)(left < fpb ? left : fpb) * frame_bytes;
                if (s >= p.nbuf) mbar_wait(&drained[buf], (uint32_t)(((s / p.nbuf) - 1) & 1));
                mbar_arrive_expect_tx(&full[buf], bytes);
                bulk_g2s(smem + L.buf_off + (size_t)buf * L.buf_bytes, p.xyz + (f0 + s * fpb) * p.frame_stride, bytes,
                         &full[buf]);
                FR_STAMP(s, 7);
            }
        }
        return;
    }
    if (warp == kFrWarps + 1) {
        if (lane == 0 && !self_dma) {
            // Up to K stores in flight: store s is issued, then the storer waits only until store s-K has drained and
            // publishes that buffer.  K = 0 when every buffer is busy computing or loading (nbuf == G); small slots
            // need K > 0 or the drain latency of each 4-12 KB store would serialise the whole pipeline.
            const int K = p.nbuf == G ? 0 : min(3, p.nbuf - G - 1);
            for (int64_t s = 0; s < n_slots; ++s) {
                const int buf = (int)(s % p.nbuf);
                const int64_t left = n - s * fpb;
                const uint32_t bytes = (uint32_t)(left < fpb ? left : fpb) * frame_bytes;
                mbar_wait(&done[buf], (uint32_t)((s / p.nbuf) & 1));
                bulk_s2g(p.xyz + (f0 + s * fpb) * p.frame_stride, smem + L.buf_off + (size_t)buf * L.buf_bytes, bytes);
                bulk_commit();
                FR_STAMP(s, 5);
                switch (K) {
                    case 0: bulk_wait_read<0>(); break;
                    case 1: bulk_wait_read<1>(); break;
                    case 2: bulk_wait_read<2>(); break;
                    default: bulk_wait_read<3>(); break;
                }
                if (s >= K) mbar_arrive(&drained[(int)((s - K) % p.nbuf)]);  // that buffer may be refilled
                FR_STAMP(s, 6);
            }
            bulk_wait_read<0>();
            for (int64_t s = max((int64_t)0, n_slots - K); s < n_slots; ++s) mbar_arrive(&drained[(int)(s % p.nbuf)]);
            bulk_wait<0>();
            if (g_fr_trace) {  // per-CTA wall time (ns) after the per-slot stamps of CTA 0
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
    // teams: tw warps share a frame; a group has wpf/tw teams, team `team` takes frames team, team+n_teams, ... of a slot
    const int n_teams = wpf / tw, team = sub / tw, ts = sub - team * tw;
    // inside a team: tw > 1 -> all tw*32 threads stride over one frame; tw == 1 -> the warp is cut into 32/Lt lane groups
    // of Lt lanes and passes over 32/Lt frames at a time (small frames would leave most of a warp idle)
    const int Lt = tw > 1 ? 32 : p.lanes;
    const int fpi = 32 / Lt;                       // frames per warp pass
    const int sj = lane / Lt;                      // which of them this lane works on
    const int ttid = tw > 1 ? ts * 32 + lane : lane - sj * Lt, tthreads = tw > 1 ? tw * 32 : Lt;
    if (OP == OP_SUPERPOSE && p.idx)
        for (int k = tid; k < p.n_sel; k += n_cw * 32) idx_s[k] = __ldg(p.idx + k);
    if (OP == OP_SUPERPOSE) mbar_wait(ref_bar, 0);
    asm volatile("bar.sync 0, %0;" ::"r"(n_cw * 32) : "memory");  // idx_s visible to all compute warps

#pragma unroll 1
    for (int64_t s = g; s < n_slots; s += G) {
        const int buf = (int)(s % p.nbuf);
        // nbuf is a multiple of G, so buffer `buf` is only ever used by this group and the group observes every phase
        // of its barrier in order (a parity wait is ambiguous for a waiter that skips phases)
        mbar_wait(&full[buf], (uint32_t)((s / p.nbuf) & 1));
        if (gtid == 0) FR_STAMP(s, 0);
        float* slot_s = reinterpret_cast<float*>(smem + L.buf_off + (size_t)buf * L.buf_bytes);
        const int64_t left = n - s * fpb;
        const int cnt = (int)(left < fpb ? left : fpb);
        const int64_t fbase = f0 + s * fpb;

        if (OP == OP_SUPERPOSE) {
            // ---- phase 1: sums over the align selection (pivot = first selected atom), frame by frame
#pragma unroll 1
            for (int j0 = team * fpi; j0 < cnt; j0 += n_teams * fpi) {
                const bool act = j0 + sj < cnt;            // idle lane groups of a partial pass still take part in shuffles
                const int j = act ? j0 + sj : j0;
                const int n_sel_t = act ? p.n_sel : 0, units_t = act ? units : 0;
                const float* frame_s = slot_s + (size_t)j * frame_floats;
                const float4* xs = reinterpret_cast<const float4*>(frame_s);
                float v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = 0.f;
                const int a0 = p.idx ? idx_s[0] : 0;
                const float px = frame_s[3 * a0], py = frame_s[3 * a0 + 1], pz = frame_s[3 * a0 + 2];
                if (p.idx) {
#pragma unroll 4
                    for (int k = ttid; k < n_sel_t; k += tthreads) {
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
#pragma unroll 2
                    for (int u = ttid; u < units_t; u += tthreads) {
                        const float4 a0v = xs[3 * u], a1v = xs[3 * u + 1], a2v = xs[3 * u + 2];
                        const float4 b0v = ys[3 * u], b1v = ys[3 * u + 1], b2v = ys[3 * u + 2];
                        acc_unit<false>(v, a0v, a1v, a2v, b0v, b1v, b2v, px, py, pz, p.n_atoms - 4 * u);
                    }
                }
                if (ttid == 0 && act) { v[13] = px; v[14] = py; v[15] = pz; }
                float* rec_s = rec_all + (((size_t)g * fpb + j) * tw + ts) * kRecStride;
                switch (Lt) {
                    case 2: team_reduce_store<2>(v, lane, act, rec_s); break;
                    case 4: team_reduce_store<4>(v, lane, act, rec_s); break;
                    case 8: team_reduce_store<8>(v, lane, act, rec_s); break;
                    case 16: team_reduce_store<16>(v, lane, act, rec_s); break;
                    default:
                        warp_reduce_scatter16(v, lane);
                        if (!(lane & 1)) rec_s[lane >> 1] = v[0];
                }
            }
            group_sync(g, wpf);
            if (tw > 1 && sub == 0) {
                // several warps per frame: 16 lanes add up the team's partials in float64, one value each (the solver
                // lane would otherwise convert and add 16*tw numbers one after the other on the critical path)
                if (lane < 16) {
                    for (int j = 0; j < cnt; ++j) {
                        double d = 0.0;
                        for (int w = 0; w < tw; ++w) d += (double)rec_all[(((size_t)g * fpb + j) * tw + w) * kRecStride + lane];
                        dsum_s[((size_t)g * fpb + j) * 16 + lane] = d;
                    }
                }
                __syncwarp();
            }
            // ---- solve: one lane per frame of the slot, side by side
            if (gtid < cnt) {
                if (gtid == 0) FR_STAMP(s, 1);
                const int j = gtid;
                float* rec0 = rec_all + ((size_t)g * fpb + j) * tw * kRecStride;   // the frame's first record
                fr_solve_frame(tw > 1 ? nullptr : rec0, dsum_s + ((size_t)g * fpb + j) * 16, rec0, rs_s, p.inv_n_sel,
                               fbase + j, p);
                if (gtid == 0) FR_STAMP(s, 2);
            }
            group_sync(g, wpf);
            // ---- phase 2: transform every atom of every frame of the slot in shared memory
#pragma unroll 1
            for (int j0 = team * fpi; j0 < cnt; j0 += n_teams * fpi) {
                const int j = j0 + sj;
                if (j >= cnt) continue;
                float4* xs = reinterpret_cast<float4*>(slot_s + (size_t)j * frame_floats);
                float t[15];
#pragma unroll
                for (int c = 0; c < 15; ++c) t[c] = rec_all[((size_t)g * fpb + j) * tw * kRecStride + c];
#pragma unroll 2
                for (int u = ttid; u < units; u += tthreads) {
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
            // tw == wpf: the group's warps share each frame (cnt frames one after the other, the loop is uniform across
            // the group so the named barriers inside are safe); tw == 1: every warp centres whole frames by itself
#pragma unroll 1
            for (int j0 = team * fpi; j0 < cnt; j0 += n_teams * fpi) {
                const bool act = j0 + sj < cnt;
                const int j = act ? j0 + sj : j0;
                const int units_t = act ? units : 0;
                float4* xs = reinterpret_cast<float4*>(slot_s + (size_t)j * frame_floats);
                double sx = 0, sy = 0, sz = 0;
                for (int u = ttid; u < units_t; u += tthreads) {
                    const float4 a0v = xs[3 * u], a1v = xs[3 * u + 1], a2v = xs[3 * u + 2];
                    sx += (double)a0v.x; sy += (double)a0v.y; sz += (double)a0v.z;
                    sx += (double)a0v.w; sy += (double)a1v.x; sz += (double)a1v.y;
                    sx += (double)a1v.z; sy += (double)a1v.w; sz += (double)a2v.x;
                    sx += (double)a2v.y; sy += (double)a2v.z; sz += (double)a2v.w;
                }
                sx = lanes_sum(sx, Lt); sy = lanes_sum(sy, Lt); sz = lanes_sum(sz, Lt);
                if (tw > 1) {
                    if (lane == 0) { cpart_s[warp * 4] = sx; cpart_s[warp * 4 + 1] = sy; cpart_s[warp * 4 + 2] = sz; }
                    group_sync(g, wpf);
                    sx = sy = sz = 0;
                    for (int w = 0; w < tw; ++w) {
                        sx += cpart_s[(g * wpf + w) * 4]; sy += cpart_s[(g * wpf + w) * 4 + 1];
                        sz += cpart_s[(g * wpf + w) * 4 + 2];
                    }
                }
                const float mx = (float)(sx / p.n_atoms), my = (float)(sy / p.n_atoms), mz = (float)(sz / p.n_atoms);
                double tr = 0;
                for (int u = ttid; u < units_t; u += tthreads) {
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
                tr = lanes_sum(tr, Lt);
                if (tw > 1) {
                    if (lane == 0) cpart_s[warp * 4 + 3] = tr;
                    group_sync(g, wpf);
                    if (ttid == 0) {
                        tr = 0;
                        for (int w = 0; w < tw; ++w) tr += cpart_s[(g * wpf + w) * 4 + 3];
                    }
                }
                if (ttid == 0 && act && p.traces) p.traces[fbase + j] = (float)tr;
            }
        }
        // ---- publish the modified slot to the async proxy and hand the buffer to the DMA thread
        if (gtid == 0) FR_STAMP(s, 3);
        fence_proxy_async_smem();
        group_sync(g, wpf);
        if (gtid == 0) {
            if (!self_dma) {
                mbar_arrive(&done[buf]);
                FR_STAMP(s, 4);
            } else {
                FR_STAMP(s, 4);
                const uint32_t bytes = (uint32_t)cnt * frame_bytes;
                bulk_s2g(p.xyz + fbase * p.frame_stride, slot_s, bytes);
                bulk_commit();
                FR_STAMP(s, 5);
                bulk_wait_read<0>();  // the buffer has been read: it may be refilled
                FR_STAMP(s, 6);
                const int64_t s2 = s + G;
                if (s2 < n_slots) {
                    const int64_t left2 = n - s2 * fpb;
                    const uint32_t bytes2 = (uint32_t)(left2 < fpb ? left2 : fpb) * frame_bytes;
                    mbar_arrive_expect_tx(&full[buf], bytes2);
                    bulk_g2s(slot_s, p.xyz + (f0 + s2 * fpb) * p.frame_stride, bytes2, &full[buf]);
                    FR_STAMP(s2, 7);
                }
            }
        }
    }
    // self_dma: the stores of this thread must have completed (not just been read) before the CTA's memory goes away
    if (self_dma && gtid == 0) bulk_wait<0>();
}


// =====================================================================================================================
// Stage-pipelined superpose.  The ring kernel above gives every slot to ONE group of 1-2 warps from arrival to hand-over:
// a 25 KB slot then sits in shared memory for ~19 500 cycles (tools/fused_trace.py, N = 300: load 3800, sums 2800, the
// float64 solve 6700 with the group's other warp idle at its barrier, transform 4400, store 1300) while the HBM rate
// only allows 8 x 2200, and everything but N = 500 / 1000 / 3000 stayed at 0.78 of peak.  Here the stages of a slot are
// taken by different warps, so that the streaming stages cost a slot ~1/8 of that time and nobody waits for a solve:
//
//   loader (warp 16)        bulk load slot s into buffer s % nbuf as soon as it is drained               -> full[b]
//   streaming warps 0..T-1  ALL of them on one slot at a time: sums of slot i (one lane group or W warps per frame)
//                           -> reduced[b]; then the transform of slot i - depth, whose solve has finished meanwhile
//                           (waits on solved[b])                                                         -> done[b]
//   solver warps T..15      warp v takes slots v, v + S, ...: one lane per frame, float64 QCP solve + transform record
//                           (waits on reduced[b])                                                        -> solved[b]
//   storer (warp 17)        bulk store, then drained[b] once the copy has read the buffer
//
// A buffer is in flight to or from HBM, in one of the two short streaming stages, or waiting for its solve; `depth`
// slots are between the sums and the transform, S solver warps work on different slots at once (S ~ solve latency /
// slot period + 1: 8 for 22-atom frames where the solves dominate, 3 for 2000 atoms).
// mbarrier parity: the streaming warps, the loader and the storer see every phase of every barrier they wait on, in
// order.  A solver warp only waits on the slots it owns; the phase before one of those was completed by the streaming
// warps before they completed the solver's previous slot (slots are reduced in order and nbuf >= S), and the phase after
// it cannot start before the solver has finished (the buffer is not reloaded until its slot is stored).
// =====================================================================================================================
constexpr int kPipeMaxFrames = 32;  // frames per slot: one solver lane each

struct PipeLayout {
    size_t buf_off, ref_off, idx_off, rec_off, dsum_off, bar_off, rs_off, total;
    size_t buf_bytes;
};
// rec: per buffer and frame, W records of kRecStride floats (partial sums per streaming warp of the frame; the first
// one is overwritten with the frame's transform by its solver lane); dsum: W > 1 only, the partials added in float64
__host__ __device__ inline PipeLayout pipe_layout(int n_pad, int nbuf, int n_sel_pad, int n_idx, int fpb, int W)
{
    PipeLayout L;
    L.buf_bytes = (size_t)n_pad * 12 * fpb;
    L.buf_off = 0;
    L.ref_off = fr_align((size_t)nbuf * L.buf_bytes, 128);
    L.idx_off = L.ref_off + fr_align((size_t)n_sel_pad * 12, 16);
    L.rec_off = fr_align(L.idx_off + (size_t)n_idx * 4, 16);
    L.dsum_off = fr_align(L.rec_off + (size_t)nbuf * fpb * W * kRecStride * sizeof(float), 8);
    L.bar_off = L.dsum_off + (W > 1 ? (size_t)nbuf * fpb * 16 * sizeof(double) : 0);
    L.rs_off = L.bar_off + (size_t)(5 * nbuf + 1) * sizeof(uint64_t);  // RefStats copy
    L.total = L.rs_off + 64;
    return L;
}

__global__ void __launch_bounds__(kFrThreads, 1) superpose_pipe_kernel(const FusedParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int n_sel_pad = (p.n_sel + 3) & ~3;
    const int S = p.batch, T = kFrWarps - S, D = p.depth, W = p.team_warps, fpb = p.fpb, nbuf = p.nbuf;
    const PipeLayout L = pipe_layout(p.n_pad, nbuf, p.ref_global ? 0 : n_sel_pad, p.idx ? p.n_sel : 0, fpb, W);
    const float* ref_s = p.ref_global ? p.ref : reinterpret_cast<const float*>(smem + L.ref_off);
    int* idx_s = reinterpret_cast<int*>(smem + L.idx_off);
    float* rec_all = reinterpret_cast<float*>(smem + L.rec_off);
    double* dsum_s = reinterpret_cast<double*>(smem + L.dsum_off);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* reduced = full + nbuf;
    uint64_t* solved = reduced + nbuf;
    uint64_t* done = solved + nbuf;
    uint64_t* drained = done + nbuf;
    uint64_t* ref_bar = drained + nbuf;

    // warp-uniform by construction (a shuffle), so that the DMA warps' loops stay on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int units = p.n_pad >> 2;
    const uint32_t frame_bytes = (uint32_t)p.n_pad * 12u;
    const int frame_floats = p.n_pad * 3;
    const int64_t f0 = p.n_frames * blockIdx.x / gridDim.x, f1 = p.n_frames * (blockIdx.x + 1) / gridDim.x;
    const int64_t n = f1 - f0;
    const int n_slots = (int)((n + fpb - 1) / fpb);

    if (threadIdx.x == 0) {
        for (int i = 0; i < nbuf; ++i) {
            mbar_init(&full[i], 1); mbar_init(&reduced[i], T); mbar_init(&solved[i], 1); mbar_init(&done[i], T);
            mbar_init(&drained[i], 1);
        }
        mbar_init(ref_bar, 1);
        *reinterpret_cast<RefStats*>(smem + L.rs_off) = *p.ref_stats;
    }
    fence_mbar_init();
    __syncthreads();

    // ================================================================== loader
    if (warp == kFrWarps) {
        if (elect_one_sync()) {
            if (p.ref_global) {
                mbar_arrive(ref_bar);
            } else {
                const uint32_t bytes = (uint32_t)n_sel_pad * 12u;
                mbar_arrive_expect_tx(ref_bar, bytes);
                bulk_g2s(smem + L.ref_off, p.ref, bytes, ref_bar);
            }
        }
        for (int s = 0; s < n_slots; ++s) {
            const int buf = s % nbuf;
            if (s >= nbuf) mbar_wait(&drained[buf], (uint32_t)(((s / nbuf) - 1) & 1));
            if (elect_one_sync()) {
                const int64_t left = n - (int64_t)s * fpb;
                const uint32_t bytes = (uint32_t)(left < fpb ? left : fpb) * frame_bytes;
                mbar_arrive_expect_tx(&full[buf], bytes);
                bulk_g2s(smem + L.buf_off + (size_t)buf * L.buf_bytes, p.xyz + (f0 + (int64_t)s * fpb) * p.frame_stride, bytes,
                         &full[buf]);
            }
        }
        return;
    }
    // ================================================================== storer
    if (warp == kFrWarps + 1) {
        // up to K stores in flight: store s is issued, then the storer waits until store s-K has been read out of its
        // buffer and publishes that buffer (see the ring kernel above)
        int K = nbuf - D - 3;
        K = K < 0 ? 0 : (K > 3 ? 3 : K);
        for (int s = 0; s < n_slots; ++s) {
            const int buf = s % nbuf;
            mbar_wait(&done[buf], (uint32_t)((s / nbuf) & 1));
            if (elect_one_sync()) {
                const int64_t left = n - (int64_t)s * fpb;
                const uint32_t bytes = (uint32_t)(left < fpb ? left : fpb) * frame_bytes;
                bulk_s2g(p.xyz + (f0 + (int64_t)s * fpb) * p.frame_stride, smem + L.buf_off + (size_t)buf * L.buf_bytes, bytes);
                bulk_commit();
                switch (K) {
                    case 0: bulk_wait_read<0>(); break;
                    case 1: bulk_wait_read<1>(); break;
                    case 2: bulk_wait_read<2>(); break;
                    default: bulk_wait_read<3>(); break;
                }
                if (s >= K) mbar_arrive(&drained[(s - K) % nbuf]);  // that buffer may be refilled
            }
        }
        if (elect_one_sync()) {
            bulk_wait_read<0>();
            for (int s = n_slots - K < 0 ? 0 : n_slots - K; s < n_slots; ++s) mbar_arrive(&drained[s % nbuf]);
            bulk_wait<0>();  // the stores must have completed before the CTA's shared memory goes away
        }
        return;
    }
    // ================================================================== solver warps
    if (warp >= T) {
        const RefStats* rs_s = reinterpret_cast<const RefStats*>(smem + L.rs_off);
#pragma unroll 1
        for (int s = warp - T; s < n_slots; s += S) {
            const int buf = s % nbuf;
            const int64_t left = n - (int64_t)s * fpb;
            const int cnt = (int)(left < fpb ? left : fpb);
            const int64_t fbase = f0 + (int64_t)s * fpb;
            float* rec_b = rec_all + (size_t)buf * fpb * W * kRecStride;
            mbar_wait(&reduced[buf], (uint32_t)((s / nbuf) & 1));
            if (W > 1) {
                // several warps per frame: 16 lanes add up the partials in float64, one value each
                if (lane < 16) {
                    for (int j = 0; j < cnt; ++j) {
                        double d = 0.0;
                        for (int w = 0; w < W; ++w) d += (double)rec_b[((size_t)j * W + w) * kRecStride + lane];
                        dsum_s[((size_t)buf * fpb + j) * 16 + lane] = d;
                    }
                }
                __syncwarp();
            }
            if (lane < cnt) {
                float* rec0 = rec_b + (size_t)lane * W * kRecStride;
                fr_solve_frame(W > 1 ? nullptr : rec0, dsum_s + ((size_t)buf * fpb + lane) * 16, rec0, rs_s, p.inv_n_sel,
                               fbase + lane, p);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&solved[buf]);
        }
        return;
    }

    // ================================================================== streaming warps
    // W == 1: a warp is cut into 32/Lt lane groups, one frame each: T * 32/Lt frames per pass of the T warps;
    // W  > 1: W warps share a frame (fpb * W <= T, warps beyond that idle), all 32 lanes on it
    const int Lt = W > 1 ? 32 : p.lanes;
    const int fpi = 32 / Lt;
    const int sj = lane / Lt;
    const int wj = W > 1 ? warp / W : 0, ts = W > 1 ? warp - wj * W : 0;
    const int ttid = W > 1 ? ts * 32 + lane : lane - sj * Lt, tthreads = W > 1 ? W * 32 : Lt;
    const int j_first = W > 1 ? wj : warp * fpi, j_step = W > 1 ? kFrWarps : T * fpi;  // W > 1: a single pass
    if (p.idx)
        for (int k = threadIdx.x; k < p.n_sel; k += T * 32) idx_s[k] = __ldg(p.idx + k);
    mbar_wait(ref_bar, 0);
    asm volatile("bar.sync 1, %0;" ::"r"(T * 32) : "memory");  // idx_s visible to all streaming warps

#pragma unroll 1
    for (int i = 0; i < n_slots + D; ++i) {
        if (i < n_slots) {
            // ---- sums over the align selection of every frame of slot i (pivot = first selected atom)
            const int buf = i % nbuf;
            const int64_t left = n - (int64_t)i * fpb;
            const int cnt = (int)(left < fpb ? left : fpb);
            const float* slot_s = reinterpret_cast<const float*>(smem + L.buf_off + (size_t)buf * L.buf_bytes);
            float* rec_b = rec_all + (size_t)buf * fpb * W * kRecStride;
            mbar_wait(&full[buf], (uint32_t)((i / nbuf) & 1));
#pragma unroll 1
            for (int j0 = j_first; j0 < cnt; j0 += j_step) {
                const bool act = j0 + sj < cnt;            // idle lane groups of a partial pass still take part in shuffles
                const int j = act ? j0 + sj : j0;
                const int n_sel_t = act ? p.n_sel : 0, units_t = act ? units : 0;
                const float* frame_s = slot_s + (size_t)j * frame_floats;
                const float4* xs = reinterpret_cast<const float4*>(frame_s);
                float v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = 0.f;
                const int a0 = p.idx ? idx_s[0] : 0;
                const float px = frame_s[3 * a0], py = frame_s[3 * a0 + 1], pz = frame_s[3 * a0 + 2];
                if (p.idx) {
#pragma unroll 4
                    for (int k = ttid; k < n_sel_t; k += tthreads) {
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
#pragma unroll 2
                    for (int u = ttid; u < units_t; u += tthreads) {
                        const float4 a0v = xs[3 * u], a1v = xs[3 * u + 1], a2v = xs[3 * u + 2];
                        const float4 b0v = ys[3 * u], b1v = ys[3 * u + 1], b2v = ys[3 * u + 2];
                        acc_unit<false>(v, a0v, a1v, a2v, b0v, b1v, b2v, px, py, pz, p.n_atoms - 4 * u);
                    }
                }
                if (ttid == 0 && act) { v[13] = px; v[14] = py; v[15] = pz; }
                float* rec_s = rec_b + ((size_t)j * W + ts) * kRecStride;
                switch (Lt) {
                    case 2: team_reduce_store<2>(v, lane, act, rec_s); break;
                    case 4: team_reduce_store<4>(v, lane, act, rec_s); break;
                    case 8: team_reduce_store<8>(v, lane, act, rec_s); break;
                    case 16: team_reduce_store<16>(v, lane, act, rec_s); break;
                    default:
                        warp_reduce_scatter16(v, lane);
                        if (!(lane & 1)) rec_s[lane >> 1] = v[0];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&reduced[buf]);
        }
        const int k = i - D;
        if (k >= 0) {
            // ---- transform of every atom of every frame of slot k, in shared memory
            const int buf = k % nbuf;
            const int64_t left = n - (int64_t)k * fpb;
            const int cnt = (int)(left < fpb ? left : fpb);
            float* slot_s = reinterpret_cast<float*>(smem + L.buf_off + (size_t)buf * L.buf_bytes);
            const float* rec_b = rec_all + (size_t)buf * fpb * W * kRecStride;
            mbar_wait(&solved[buf], (uint32_t)((k / nbuf) & 1));
#pragma unroll 1
            for (int j0 = j_first; j0 < cnt; j0 += j_step) {
                const int j = j0 + sj;
                if (j >= cnt) continue;
                float4* xs = reinterpret_cast<float4*>(slot_s + (size_t)j * frame_floats);
                float t[15];
#pragma unroll
                for (int c = 0; c < 15; ++c) t[c] = rec_b[(size_t)j * W * kRecStride + c];
#pragma unroll 2
                for (int u = ttid; u < units; u += tthreads) {
                    float4 a0v = xs[3 * u], a1v = xs[3 * u + 1], a2v = xs[3 * u + 2];
                    const int nvalid = p.n_atoms - 4 * u;
                    xf_atom(a0v.x, a0v.y, a0v.z, t);
                    if (nvalid > 1) xf_atom(a0v.w, a1v.x, a1v.y, t);
                    if (nvalid > 2) xf_atom(a1v.z, a1v.w, a2v.x, t);
                    if (nvalid > 3) xf_atom(a2v.y, a2v.z, a2v.w, t);
                    xs[3 * u] = a0v; xs[3 * u + 1] = a1v; xs[3 * u + 2] = a2v;
                }
            }
            fence_proxy_async_smem();  // the modified slot must be visible to the bulk store
            __syncwarp();
            if (lane == 0) mbar_arrive(&done[buf]);
        }
    }
}

static bool pipe_fits(const FusedParams& p, int nbuf, int fpb, int W, bool ref_global = false)
{
    const int n_sel_pad = (p.n_sel + 3) & ~3;
    return pipe_layout(p.n_pad, nbuf, ref_global ? 0 : n_sel_pad, p.idx ? p.n_sel : 0, fpb, W).total <= 232448;
}

// Geometry of the stage-pipelined superpose: slots of ~20 KB (at most 32 frames), as many buffers as fit (<= 12),
// S = solve latency / slot period + 1 solver warps, the same number of slots between the sums and the transform.
static bool pipe_config(FusedParams& p)
{
    const size_t frame_bytes = (size_t)p.n_pad * 12;
    if (frame_bytes > 100000) return false;
    const bool contiguous = p.frame_stride == (int64_t)p.n_pad * 3;
    int fpb = contiguous ? (int)(20480 / frame_bytes) : 1;
    fpb = fpb < 1 ? 1 : (fpb > kPipeMaxFrames ? kPipeMaxFrames : fpb);
    const double kSolveCycles = 6000.0;   // float64 QCP solve with rotation, one frame per lane (tools/fused_trace.py)
    for (;; --fpb) {
        const double period = (double)fpb * frame_bytes * 2.0 / 22.5;  // cycles a slot takes at the HBM rate of one SM
        int S = (int)(kSolveCycles / period) + 2;
        S = S < 2 ? 2 : (S > 8 ? 8 : S);
        int T = kFrWarps - S;
        int W = 1, lanes = 32;
        if (fpb >= T) {
            while (lanes > 2 && T * (32 / lanes) < fpb) lanes >>= 1;  // the widest lane group that covers the slot in one pass
        } else {
            W = T / fpb;
        }
        int nbuf = 12;
        while (nbuf >= 2 && !pipe_fits(p, nbuf, fpb, W)) --nbuf;
        p.ref_global = 0;
        if (nbuf < 3 && fpb == 1) {
            // one frame per slot and fewer than three buffers next to a resident reference (all atoms selected on
            // ~5,000-atom frames): leave the reference in global memory -- every CTA reads the same <= 100 KB, so it stays
            // in L2 -- and keep load / compute / store of three frames overlapped instead of two passes over HBM
            int nb = 12;
            while (nb >= 3 && !pipe_fits(p, nb, fpb, W, true)) --nb;
            if (nb >= 3) { nbuf = nb; p.ref_global = 1; }
        }
        if (nbuf < 2) {
            if (fpb > 1) continue;
            return false;
        }
        if (S > nbuf) { S = nbuf; T = kFrWarps - S; if (W > 1) W = T / fpb; }
        int D = (int)(kSolveCycles / period) + 2;
        if (D > nbuf - 3) D = nbuf - 3;
        if (D < 1) D = 1;
        p.pipe = 1; p.batch = S; p.nbuf = nbuf; p.fpb = fpb; p.team_warps = W; p.lanes = lanes; p.depth = D;
        return true;
    }
}

static int fr_team_warps(int G, int fpb)
{
    const int wpf = kFrWarps / G;
    return fpb >= wpf ? 1 : wpf;  // enough frames per slot to give every warp its own, else share each frame
}

static bool fr_fits(const FusedParams& p, int op, int G, int fpb, int nbuf)
{
    const int n_sel_pad = (p.n_sel + 3) & ~3;
    return fr_layout(p.n_pad, nbuf, op == OP_SUPERPOSE ? n_sel_pad : 0, p.idx ? p.n_sel : 0, G, fpb, fr_team_warps(G, fpb),
                     op == OP_SUPERPOSE).total <= 232448;
}

// Work split inside a group for (G, fpb):
//   fpb <  wpf : the group's wpf warps share each frame (team_warps = wpf, lanes = 32);
//   fpb >= wpf : every warp takes whole frames (team_warps = 1), and a warp is cut into lane groups of `lanes` lanes, one
//                frame each, so that a warp passes over 32/lanes frames at a time.
static void fr_set(FusedParams& p, int G, int fpb, int nbuf, int lanes)
{
    p.batch = G; p.nbuf = nbuf; p.fpb = fpb; p.team_warps = fr_team_warps(G, fpb);
    p.lanes = p.team_warps == 1 ? lanes : 32;
}

// Lanes per frame when every warp takes whole frames: the widest lane group that still covers the slot in ONE pass of the
// group's warps.  Every pass costs ~1800 cycles (reduction, pivot and transform loads, loop set-up, barriers) against
// 350-950 per unit a lane walks: N = 50 with 32 frames per slot runs at 0.88x of HBM peak on 2 lanes per frame (one pass)
// and 0.78x on 4 (two passes); 39 frames per slot (two passes on 2 lanes) 0.80x.
static int fr_one_pass_lanes(int G, int fpb)
{
    const int wpf = kFrWarps / G;
    int lanes = 32;
    while (lanes > 2 && wpf * (32 / lanes) < fpb) lanes >>= 1;
    return lanes;
}

// frames per slot: one solver lane per frame, the solver lanes are the group's first min(64, wpf*32) threads
static int fr_cap(int G, bool contiguous) { return !contiguous ? 1 : (kFrWarps / G >= 2 ? 64 : 32); }

static int fr_max_fpb(const FusedParams& p, int op, int G, int nbuf, int cap)
{
    int fpb = 0;
    while (fpb < cap && fr_fits(p, op, G, fpb + 1, nbuf)) ++fpb;  // the layout grows monotonically with fpb
    return fpb;
}

// Cost model of one group working through one single-buffer slot, least-squares fit (median error 5 %) to ~900
// geometries per operation measured with tools/fused_sweep.py (N = 22 ... 1000).  Nothing overlaps inside a group, so a
// slot costs  F (load/store latency + the float64 solve, ~4400 cycles however many frames are solved side by side)
//           + S per KB moved + per pass of the group's warps over the slot ( P + c per unit a lane walks ).
// Returns bytes per clock per SM (22.5 = HBM peak).
static double fr_model(const FusedParams& p, int op, int G, int fpb, int lanes)
{
    const int wpf = kFrWarps / G, units = p.n_pad / 4, tw = fr_team_warps(G, fpb);
    const double sel = (op == OP_SUPERPOSE && p.idx) ? (double)p.n_sel / p.n_atoms : 1.0;
    const double F = op == OP_SUPERPOSE ? 7250.0 : 2800.0, S = op == OP_SUPERPOSE ? 110.0 : 90.0;
    const double c = op == OP_SUPERPOSE ? 216.0 + 552.0 * sel : 690.0;
    double P = op == OP_SUPERPOSE ? 1740.0 : 1980.0;
    int passes, iters;
    if (tw == 1) {
        const int frames_per_pass = wpf * (32 / lanes);
        passes = (fpb + frames_per_pass - 1) / frames_per_pass;
        iters = (units + lanes - 1) / lanes;
    } else {
        passes = fpb;
        iters = (units + wpf * 32 - 1) / (wpf * 32);
        P += op == OP_SUPERPOSE ? 50.0 : 600.0;  // named barriers between the warps of the group
    }
    const double slot_kb = (double)fpb * p.n_pad * 12.0 / 1000.0;
    const double t = F + S * slot_kb + passes * (P + iters * c);
    return (double)G * slot_kb * 2000.0 / t;
}

// Geometry (measured with tools/fused_sweep.py, N = 22 ... 5000):
//  * nbuf is a multiple of G so that a buffer always belongs to one group (see the parity-wait note in the kernel);
//  * superpose: the serial float64 solve makes the compute phase long, so what pays is many groups with ONE slot each,
//    as large as shared memory allows: contiguous frames are packed fpb to a slot (N = 1000: 2 frames = 24 KB per slot on
//    8 groups, N = 300: 12 frames on 5 groups, N = 22: 64).  That also takes small frames off the DMA threads' issue
//    limit (~1 bulk copy per 100-150 ns each way);
//    (G, fpb) is the best of the cost model above, lanes the one-pass choice;
//  * centring: 5 single-slot groups up to 12 KB frames, the model up to 20 KB, larger frames two slots per group (one
//    computing, one in flight).
bool fused_config(FusedParams& p, int op)
{
    const size_t frame_bytes = (size_t)p.n_pad * 12;
    if (frame_bytes >= (1u << 20)) return false;
    p.pipe = 0;
    if (op == OP_SUPERPOSE && (int64_t)frame_bytes >= g_fr_pipe_min_bytes && pipe_config(p)) return true;
    const bool contiguous = p.frame_stride == (int64_t)p.n_pad * 3;
    if (op == OP_CENTER && frame_bytes <= 12288) {
        // centring is light enough to be HBM-bound in the model whatever the geometry; measured best (0.96-0.99x for
        // N = 50 ... 1000) are 5 groups with one slot of 24-45 KB each
        const int fit = fr_max_fpb(p, op, 5, 5, fr_cap(5, contiguous));
        if ((size_t)fit * frame_bytes >= 24576) {
            fr_set(p, 5, fit, 5, fr_one_pass_lanes(5, fit));
            return true;
        }
    }
    if (op == OP_SUPERPOSE || frame_bytes <= 20480) {
        static const int kGroups[] = {8, 5, 4, 3};
        double best = 0.0;
        for (int G : kGroups) {  // in descending order: fewer, larger slots only for a clear (2 %) modelled gain
            const int fit = fr_max_fpb(p, op, G, G, fr_cap(G, contiguous));
            double best_g = 0.0;
            int fpb_g = 0;
            for (int fpb = 1; fpb <= fit; ++fpb) {
                const double r = fr_model(p, op, G, fpb, fr_one_pass_lanes(G, fpb));
                if (r >= best_g) { best_g = r; fpb_g = fpb; }
            }
            if (fpb_g && best_g > best * 1.02) {
                best = best_g;
                fr_set(p, G, fpb_g, G, fr_one_pass_lanes(G, fpb_g));
            }
        }
        if (best > 0.0) {
            // One frame per slot on 8 groups leaves shared memory half empty for 13-22 KB frames (1100-1800 atoms).  More
            // groups of a single warp each fill it (the cost model does not cover them): measured N=1400 0.80x -> 0.94x
            // with 12 groups, N=1600 0.90x -> 0.98x with 10; centring N=1400-1600 0.94-0.98x -> 0.97-0.99x with 10.
            if (p.batch == 8 && p.fpb == 1) {
                if (op == OP_SUPERPOSE && fr_fits(p, op, 12, 1, 12)) fr_set(p, 12, 1, 12, 32);
                else if (fr_fits(p, op, 10, 1, 10)) fr_set(p, 10, 1, 10, 32);
            }
            return true;
        }
    } else {
        static const int kGroups[] = {8, 4, 3, 2};
        const int want = (int)(16384 / frame_bytes) < 1 ? 1 : (int)(16384 / frame_bytes);
        for (int G : kGroups) {  // the shallowest ring with at least ~48 KB of prefetch in flight
            for (int m = 2; m <= 4; ++m) {
                const int cap = fr_cap(G, contiguous);
                const int fpb = fr_max_fpb(p, op, G, G * m, want < cap ? want : cap);
                if (fpb < 1 || (size_t)(m - 1) * G * fpb * frame_bytes < 49152) continue;
                fr_set(p, G, fpb, G * m, fr_one_pass_lanes(G, fpb));
                return true;
            }
        }
        for (int G : kGroups) {  // large frames: one buffer per group, >= 3 groups
            if (G >= 3 && fr_fits(p, op, G, 1, G)) {
                fr_set(p, G, 1, G, 32);
                return true;
            }
        }
    }
    if (fr_fits(p, op, 1, 1, 3)) {  // one group of 16 warps with a 3-deep ring (measured 0.64x of HBM peak at N = 5000)
        fr_set(p, 1, 1, 3, 32);
        return true;
    }
    return false;
}

// Frames of at least this many bytes take the stage-pipelined superpose kernel, smaller ones the ring kernel: a slot of
// small frames costs every streaming warp of the pipelined kernel its fixed per-pass overhead (~900 cycles against a slot
// period of 800-1600), which the ring kernel pays once per slot (profiles/r02_fused_sweep_pipe*.jsonl).
int g_fr_pipe_min_bytes = 30000;
extern "C" int b200rmsd_debug_fused_pipe_min_bytes(int min_frame_bytes)  // development / sweeps; returns the old value
{
    const int was = g_fr_pipe_min_bytes;
    g_fr_pipe_min_bytes = min_frame_bytes;
    return was;
}

// development override: frames per slot, G concurrent groups, ring depth, lanes per frame (checked against shared memory)
bool fused_override(FusedParams& p, int op, int G, int nbuf, int fpb, int lanes)
{
    if (p.pipe) {  // pipelined superpose: G = solver warps, nbuf, fpb; lanes = depth
        if (G > 0) p.batch = G;
        if (nbuf > 0) p.nbuf = nbuf;
        if (lanes > 0) p.depth = lanes;
        if (fpb > 0 && fpb != p.fpb) return false;
        if (p.batch < 1 || p.batch > 8 || p.nbuf < 2 || p.nbuf < p.batch || p.depth < 1 || p.depth > p.nbuf - 1) return false;
        const int T = kFrWarps - p.batch;
        if (p.team_warps > 1) p.team_warps = T / p.fpb;
        else if (T * (32 / p.lanes) < p.fpb) return false;
        return p.team_warps >= 1 && pipe_fits(p, p.nbuf, p.fpb, p.team_warps);
    }
    if (G <= 0) G = p.batch;
    if (fpb <= 0) fpb = p.fpb;
    if (nbuf <= 0) nbuf = G * (p.nbuf / p.batch);
    if (G > 16 || G < 1 || fpb > fr_cap(G, p.frame_stride == (int64_t)p.n_pad * 3)) return false;
    if (nbuf < G || nbuf % G != 0 || nbuf < 2) return false;
    if (!fr_fits(p, op, G, fpb, nbuf)) return false;
    if (lanes > 0 && ((lanes & (lanes - 1)) || lanes < 2 || lanes > 32)) return false;
    if (lanes > 0 && lanes != 32 && fr_team_warps(G, fpb) != 1) return false;
    fr_set(p, G, fpb, nbuf, lanes > 0 ? lanes : fr_one_pass_lanes(G, fpb));
    return true;
}

long long* fr_trace_buffer(size_t n_frames_cta)
{
    static long long* buf = nullptr;
    static size_t cap = 0;
#ifndef B200RMSD_DEV_SWITCHES
    (void)buf; (void)cap; (void)n_frames_cta;
    return nullptr;
#else
    if (!getenv("B200RMSD_FUSED_TRACE")) return nullptr;
    if (cap < n_frames_cta * 8) {
        if (buf) cudaFree(buf);
        cudaMalloc((void**)&buf, n_frames_cta * 8 * sizeof(long long));
        cap = n_frames_cta * 8;
    }
    cudaMemset(buf, 0, cap * sizeof(long long));
    cudaMemcpyToSymbol(g_fr_trace, &buf, sizeof(buf));
    return buf;
#endif
}

// development / test hook (no device needed): the geometry fused_config picks; out = {G, nbuf, fpb, team_warps, lanes,
// shared-memory bytes}.  Returns 0 when the single-pass kernel does not apply.
extern "C" int b200rmsd_debug_fused_geometry(int op, int n_atoms, int n_sel, int has_idx, int contiguous, int* out)
{
    FusedParams p{};
    p.n_atoms = n_atoms;
    p.n_pad = (n_atoms + 3) / 4 * 4;
    p.frame_stride = contiguous ? (int64_t)p.n_pad * 3 : (int64_t)p.n_pad * 3 + 4;
    p.n_sel = has_idx ? n_sel : n_atoms;
    p.idx = has_idx ? reinterpret_cast<const int*>(0x10) : nullptr;  // only tested against nullptr
    if (!fused_config(p, op)) return 0;
    const int n_sel_pad = (p.n_sel + 3) & ~3;
    if (p.pipe) {  // returns 2: {solver warps, nbuf, fpb, streaming warps per frame, lanes per frame, bytes}; depth in out[6]
        const PipeLayout PL = pipe_layout(p.n_pad, p.nbuf, p.ref_global ? 0 : n_sel_pad, p.idx ? p.n_sel : 0, p.fpb, p.team_warps);
        out[0] = p.batch; out[1] = p.nbuf; out[2] = p.fpb; out[3] = p.team_warps; out[4] = p.lanes; out[5] = (int)PL.total;
        out[6] = p.depth; out[7] = p.ref_global;
        return 2;
    }
    const FrLayout L = fr_layout(p.n_pad, p.nbuf, op == OP_SUPERPOSE ? n_sel_pad : 0, p.idx ? p.n_sel : 0, p.batch, p.fpb,
                                 p.team_warps, op == OP_SUPERPOSE);
    out[0] = p.batch; out[1] = p.nbuf; out[2] = p.fpb; out[3] = p.team_warps; out[4] = p.lanes; out[5] = (int)L.total;
    return 1;
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
    if (op == OP_SUPERPOSE && p.pipe) {
        const int n_sel_pad = (p.n_sel + 3) & ~3;
        const PipeLayout L = pipe_layout(p.n_pad, p.nbuf, p.ref_global ? 0 : n_sel_pad, p.idx ? p.n_sel : 0, p.fpb, p.team_warps);
        cudaError_t e = cudaFuncSetAttribute(superpose_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
        if (e != cudaSuccess) return e;
        int64_t ctas = sm_count;
        const int64_t need = (p.n_frames + p.fpb - 1) / p.fpb;
        if (ctas > need) ctas = need;
        FusedParams q = p;
        q.inv_n_sel = 1.0 / (double)(p.n_sel > 0 ? p.n_sel : 1);
        superpose_pipe_kernel<<<(unsigned)ctas, kFrThreads, L.total, st>>>(q);
        return cudaGetLastError();
    }
    fr_trace_buffer((size_t)(p.n_frames / (sm_count > 0 ? sm_count : 1) + 2) + 64);
    const int n_sel_pad = (p.n_sel + 3) & ~3;
    const FrLayout L = fr_layout(p.n_pad, p.nbuf, op == OP_SUPERPOSE ? n_sel_pad : 0, p.idx ? p.n_sel : 0, p.batch, p.fpb,
                                 p.team_warps, op == OP_SUPERPOSE);
    auto kern = op == OP_SUPERPOSE ? frame_resident_kernel<OP_SUPERPOSE> : frame_resident_kernel<OP_CENTER>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) return e;
    int64_t ctas = sm_count;
    const int64_t per_round = (int64_t)p.batch * p.fpb;
    const int64_t need = (p.n_frames + per_round - 1) / per_round;
    if (ctas > need) ctas = need;
    FusedParams q = p;
    q.inv_n_sel = 1.0 / (double)(p.n_sel > 0 ? p.n_sel : 1);
    kern<<<(unsigned)ctas, kFrThreads, L.total, st>>>(q);
    return cudaGetLastError();
}

}  // namespace b200
