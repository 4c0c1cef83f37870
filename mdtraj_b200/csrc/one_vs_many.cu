// one_vs_many.cu -- the HBM-bound streaming kernels of the RMSD hot path (sm_100a).
//
//   ovm_tma_kernel      every frame of a padded atom-major trajectory against ONE
//                       reference frame: centroid + trace G + 3x3 inner product M in
//                       a single pass over HBM, then the QCP solve in registers.
//                       Replaces the three-pass CPU sequence
//                         inplace_center_and_trace_atom_major   center_sse.h:3-112
//                         msd_atom_major                        theobald_rmsd_sse.h:184-335
//                         msdFromMandG                          theobald_rmsd.cpp:217-334
//                       driven by the prange loop at _rmsd.pyx:217-224.
//   ovm_gather_kernel   same maths for an atom_indices selection (replaces the
//                       fancy-index copy at _rmsd.pyx:197 + the loop above).
//   ovm_finish_kernel   combines per-segment partial sums when a frame is split.
//
// Design:
//   * persistent grid, one 512-thread CTA per SM, static contiguous frame ranges
//     per warp (output stays coalesced, imbalance < 1 frame per warp);
//   * each warp owns a ring of shared-memory stages filled by 1-D bulk async copies
//     (cp.async.bulk -> UBLKCP) that complete on mbarriers; lane 0 is the producer,
//     all 32 lanes consume with conflict-free 128-bit shared loads (48-byte lane
//     stride = 4 whole atoms per lane);
//   * the centred reference (<= 4096 atoms per segment) is resident in shared memory
//     for the whole kernel; longer frames are split into atom segments across CTAs
//     and finished by ovm_finish_kernel;
//   * single pass: sums are taken about a per-frame pivot p close to the centroid (the
//     mean of 32 atoms spread over what is at hand: the frame's first chunk in shared
//     memory, the whole frame for short or very long frames), so that x - p is exact in
//     float32 and  G = sum|x-p|^2 - N|mu-p|^2  loses a fraction of a bit; since the
//     reference frame is centred, M needs no centring of the target;
//   * float32 lane partials (<= a few dozen terms each), with |x-p|^2 accumulated minus
//     the reference's mean-square radius so that this all-positive sum stays
//     fluctuation-sized; the cross-lane sums of the 16-value reduce-scatter are float64
//     (24 shuffles per frame); then everything -- recentring, polynomial, Newton, the
//     G_a+G_b-2*lambda cancellation, quaternion -- in float64 on one lane per frame,
//     kBatch frames at a time.  Against the float64 truth on MD-like frames this is 5-10x
//     closer than the reference's float32 SSE accumulation.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "qcp.cuh"

// Development variants (tools/build_variants.sh compiles them side by side); the defaults are the shipped kernel.
#ifndef OVM_PIVOT
#define OVM_PIVOT 2   // 0: the frame's first atom, 2: mean of 8 atoms spread over the frame
#endif
#ifndef OVM_REDUCE
#define OVM_REDUCE 1  // 0: float32 butterfly, 1: last three stages in float64
#endif
#ifndef OVM_SHIFT
#define OVM_SHIFT 1   // mean-square shift of the |x - p|^2 accumulation
#endif
#ifndef OVM_FLUSH
#define OVM_FLUSH 1   // fold the lane partials into float64 every kFlushUnits units per lane
#endif

namespace b200 {

// Pivot of the float32 accumulation: the mean of 8 atoms spread over the frame.  Lane l holds atom (l & 7) of the eight
// (lanes 8..31 duplicate them); three butterfly stages leave the bit-identical mean on every lane.  Eight atoms put the
// pivot within ~0.35 Rg of the centroid, after which the pivot is no longer what limits the accuracy (32 atoms: same
// error against the float64 truth) -- and 8 sectors per frame cost nothing measurable where 32
// cost 7-10 % of the HBM rate.
__device__ __forceinline__ float pivot_mean8(float x)
{
#if OVM_PIVOT
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    return x * 0.125f;
#else
    return x;
#endif
}
__device__ __forceinline__ double reduce16(float (&v)[16], int lane)
{
#if OVM_REDUCE == 1
    return warp_reduce_scatter16_mixed(v, lane);
#else
    warp_reduce_scatter16(v, lane);
    return (double)v[0];
#endif
}

// ---------------------------------------------------------------------------
// per-frame epilogue: 16 float32 sums -> rmsd (+ rotation, centroid)
//   rec[0..2]  = sum (x - p)          rec[3] = sum |x - p|^2
//   rec[4..12] = sum (x - p)_i y_j    rec[13..15] = pivot p
// ---------------------------------------------------------------------------
// mean-square shift of the |x - p|^2 accumulation (acc_unit_shift): a function of the reference alone
__device__ __forceinline__ float msq_shift(const OvmParams& p)
{
#if OVM_SHIFT
    return (float)(p.ref_stats->G / (double)p.n_atoms);
#else
    return 0.f;
#endif
}

template <bool PRE>
__device__ __forceinline__ void finish_frame(const double rec[16], int64_t f, const OvmParams& p, float cshift)
{
    const RefStats rs = *p.ref_stats;
    QcpInput q;
    const double invn = p.inv_n;
    q.inv_n = invn;
    q.Gb = rs.G;
    double cx = 0, cy = 0, cz = 0;
    if (PRE) {
        q.Ga = (double)p.traces[f];
#pragma unroll
        for (int i = 0; i < 9; ++i) q.M[i] = rec[4 + i];
    } else {
        const double mx = rec[0] * invn, my = rec[1] * invn, mz = rec[2] * invn;  // mean relative to pivot
        const double g_raw = rec[3] + (double)cshift * (double)p.n_atoms;  // undo the caller's shift, exactly
        double ga = g_raw - (rec[0] * mx + rec[1] * my + rec[2] * mz);
        q.Ga = ga > 0.0 ? ga : 0.0;
        // M_c = M' - N mu' (mu_y)^T ; mu_y is the float32 residual of the centred reference
        q.M[0] = rec[4] - mx * rs.sum[0];  q.M[1] = rec[5] - mx * rs.sum[1];  q.M[2] = rec[6] - mx * rs.sum[2];
        q.M[3] = rec[7] - my * rs.sum[0];  q.M[4] = rec[8] - my * rs.sum[1];  q.M[5] = rec[9] - my * rs.sum[2];
        q.M[6] = rec[10] - mz * rs.sum[0]; q.M[7] = rec[11] - mz * rs.sum[1]; q.M[8] = rec[12] - mz * rs.sum[2];
        cx = rec[13] + mx; cy = rec[14] + my; cz = rec[15] + mz;
    }
    float R[9];
    bool degen = false;
    const double msd = qcp_solve(q, p.out_rot ? R : nullptr, &degen);
    p.out_rmsd[f] = sqrtf((float)msd);
    if (p.out_rot) {
#pragma unroll
        for (int i = 0; i < 9; ++i) p.out_rot[f * 9 + i] = R[i];
        if (degen && p.degenerate) atomicAdd(p.degenerate, 1u);
    }
    if (p.out_centroid) {
        p.out_centroid[f * 3 + 0] = cx;
        p.out_centroid[f * 3 + 1] = cy;
        p.out_centroid[f * 3 + 2] = cz;
    }
}

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct OvmSmemLayout {
    size_t ref_off, ring_off, sums_off, bar_off, total;
    size_t stage_bytes;
};
__host__ __device__ inline OvmSmemLayout ovm_layout(int seg_units, int chunk_units, int stages)
{
    OvmSmemLayout L;
    L.stage_bytes = (size_t)chunk_units * 48;
    L.ref_off = 0;
    L.ring_off = align_up((size_t)seg_units * 48, 128);
    L.sums_off = L.ring_off + (size_t)kWarpsPerCta * stages * L.stage_bytes;
    L.sums_off = align_up(L.sums_off, 8);
    L.bar_off = align_up(L.sums_off + (size_t)kWarpsPerCta * kBatch * kSumStride * sizeof(double), 8);
    L.total = L.bar_off + ((size_t)kWarpsPerCta * stages + 1) * sizeof(uint64_t);
    return L;
}
size_t ovm_tma_smem_bytes(const OvmParams& p) { return ovm_layout(p.seg_units, p.chunk_units, p.stages).total; }

// ---------------------------------------------------------------------------
// the streaming kernel
// grid = n_seg * ctas_per_seg ; CTA b works on segment (b % n_seg)
// ---------------------------------------------------------------------------
// FLUSH: frames longer than kFlushUnits units per lane (1024 atoms) fold their lane partials into float64 on the way
template <bool PRE, bool FLUSH>
__global__ void __launch_bounds__(kThreadsPerCta, 1) ovm_tma_kernel(const OvmParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const OvmSmemLayout L = ovm_layout(p.seg_units, p.chunk_units, p.stages);
    const float4* ref_s = reinterpret_cast<const float4*>(smem + L.ref_off);
    double* sums_all = reinterpret_cast<double*>(smem + L.sums_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg = blockIdx.x % p.n_seg;
    const int cta_in_seg = blockIdx.x / p.n_seg;
    const int ctas_per_seg = gridDim.x / p.n_seg;
    const int seg_unit0 = seg * p.seg_units;
    const int upf = min(p.seg_units, p.total_units - seg_unit0);  // units of this segment per frame

    unsigned char* ring = smem + L.ring_off + (size_t)warp * p.stages * L.stage_bytes;
    uint64_t* my_bars = bars + warp * p.stages;
    uint64_t* ref_bar = bars + kWarpsPerCta * p.stages;
    double* sums = sums_all + warp * kBatch * kSumStride;

    if (lane == 0)
        for (int s = 0; s < p.stages; ++s) mbar_init(&my_bars[s], 1);
    if (threadIdx.x == 0) mbar_init(ref_bar, 1);
    fence_mbar_init();
    __syncthreads();

    if (threadIdx.x == 0) {  // reference segment: loaded once, stays resident
        const uint32_t bytes = (uint32_t)upf * 48u;
        mbar_arrive_expect_tx(ref_bar, bytes);
        bulk_g2s(smem + L.ref_off, p.ref + (size_t)seg_unit0 * kUnitFloats, bytes, ref_bar);
    }

    // static contiguous frame range of this warp; everything below counts frames and chunks relative to it in 32 bits
    // (a warp's share of even 10^11 frames fits), which keeps the kernel under the 128-register ceiling of 512 threads
    const int64_t W = (int64_t)ctas_per_seg * kWarpsPerCta;
    const int64_t gw = (int64_t)cta_in_seg * kWarpsPerCta + warp;
    const int64_t f_begin = p.n_frames * gw / W;
    const int n_my = (int)(p.n_frames * (gw + 1) / W - f_begin);
    const float* my_xyz = p.xyz + f_begin * p.frame_stride + (size_t)seg_unit0 * kUnitFloats;  // this warp's first frame, this segment
    const int cpf = (upf + p.chunk_units - 1) / p.chunk_units;  // bulk copies per frame
    const int total_chunks = n_my * cpf;

    // producer cursor (meaningful on lane 0 only)
    int pf = 0, pc = 0, issued = 0;
    const uint64_t pol = l2_policy_evict_first();
    auto issue = [&](int stage) {
        const int units = min(p.chunk_units, upf - pc * p.chunk_units);
        const uint32_t bytes = (uint32_t)units * 48u;
        const float* src = my_xyz + (int64_t)pf * p.frame_stride + (size_t)(pc * p.chunk_units) * kUnitFloats;
        mbar_arrive_expect_tx(&my_bars[stage], bytes);
        bulk_g2s_hint(ring + (size_t)stage * L.stage_bytes, src, bytes, &my_bars[stage], pol);
        if (++pc == cpf) { pc = 0; ++pf; }
        ++issued;
    };
    if (lane == 0)
        for (int s = 0; s < p.stages && issued < total_chunks; ++s) issue(s);

    mbar_wait(ref_bar, 0);

    int stage = 0;
    uint32_t phase = 0;
    int slot = 0;
    int batch_i0 = 0;
    const float cshift = PRE ? 0.f : msq_shift(p);
    // Pivot = mean of 8 atoms spread over the WHOLE frame (pivot_mean8), read from global memory one frame ahead of the
    // consumer: one 16-byte load per lane, 8 distinct sectors per frame, the same sectors the bulk copy of that frame
    // brings through L2.  Every segment's CTA of a long frame computes the bit-identical pivot (the finish kernel adds
    // their partial sums).  A pivot taken from the frame's first chunk alone is 2-3x worse on chain-like structures.
    const float4* piv_ptr = reinterpret_cast<const float4*>(p.xyz + f_begin * p.frame_stride) +
                            3 * (OVM_PIVOT ? (int)((((int64_t)(2 * (lane & 7) + 1)) * p.total_units) >> 4) : 0);
    // (an 8-byte and a 4-byte load, not one of 16: the unused fourth register of a 16-byte destination gets reused by the
    // compiler for the chunk counter, whose first write then waits a whole load latency on it, every frame -- ncu source
    // page of round 2, 12-17 % of all stall samples on that one MOV)
    float npx = 0.f, npy = 0.f, npz = 0.f;
    auto load_pivot = [&]() {
        const float2 t = __ldg(reinterpret_cast<const float2*>(piv_ptr));
        npx = t.x; npy = t.y;
        npz = __ldg(reinterpret_cast<const float*>(piv_ptr) + 2);
    };
    if (!PRE && n_my > 0) load_pivot();
    // lane partials are folded into float64 every kFlushUnits units per lane (1024 atoms per warp)
    const int flush_chunks = max(1, (kFlushUnits * 32) / p.chunk_units);

#pragma unroll 1
    for (int fi = 0; fi < n_my; ++fi) {
        float px = 0.f, py = 0.f, pz = 0.f;
        if (!PRE) {
            px = pivot_mean8(npx); py = pivot_mean8(npy); pz = pivot_mean8(npz);
            if (fi + 1 < n_my) {  // prefetch the next frame's pivot atoms
                piv_ptr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(piv_ptr) + p.frame_stride);
                load_pivot();
            }
        }
        double tot = 0.0;
        int until_flush = flush_chunks;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;

#pragma unroll 1
        for (int c = 0; c < cpf; ++c) {
            const int unit0 = c * p.chunk_units;
            const int units = min(p.chunk_units, upf - unit0);
            mbar_wait(&my_bars[stage], phase);
            const float4* xs = reinterpret_cast<const float4*>(ring + (size_t)stage * L.stage_bytes);
            const float4* ys = ref_s + (size_t)unit0 * 3;
            const int atom0 = (seg_unit0 + unit0) * 4;
#pragma unroll 2
            for (int u = lane; u < units; u += 32) {
                const float4 a0 = xs[3 * u], a1 = xs[3 * u + 1], a2 = xs[3 * u + 2];
                const float4 b0 = ys[3 * u], b1 = ys[3 * u + 1], b2 = ys[3 * u + 2];
                acc_unit_shift<PRE>(v, a0, a1, a2, b0, b1, b2, px, py, pz, p.n_atoms - (atom0 + 4 * u), cshift);
            }
            __syncwarp();
            if (lane == 0 && issued < total_chunks) {
                fence_proxy_async_smem();
                issue(stage);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            if (FLUSH && OVM_FLUSH && --until_flush == 0 && c + 1 < cpf) {
                tot += reduce16(v, lane);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
                until_flush = flush_chunks;
            }
        }

        if (lane == 0) { v[13] = px; v[14] = py; v[15] = pz; }
        if (FLUSH) tot += reduce16(v, lane);
        else tot = reduce16(v, lane);

        if (p.n_seg > 1) {
            if (!(lane & 1)) p.partials[((size_t)(f_begin + fi) * p.n_seg + seg) * 16 + (lane >> 1)] = tot;
        } else {
            if (!(lane & 1)) sums[slot * kSumStride + (lane >> 1)] = tot;
            ++slot;
            if (slot == kBatch || fi + 1 == n_my) {
                __syncwarp();
                if (lane < slot) {
                    double rec[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) rec[i] = sums[lane * kSumStride + i];
                    finish_frame<PRE>(rec, f_begin + batch_i0 + lane, p, cshift);
                }
                __syncwarp();
                slot = 0;
                batch_i0 = fi + 1;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Short frames (<= 256 atoms): L lanes per frame, 32/L frames per warp iteration.
// A whole-warp pass over a 22- or 100-atom frame leaves most lanes idle and pays the reduction, the mbarrier
// round trip and the copy issue once per frame; here one bulk copy brings 32/L consecutive frames, each lane group
// reduces over log2(L) shuffle stages only, and the float64 solve runs on full 32-frame batches.
// ---------------------------------------------------------------------------
constexpr int kGroupBatch = 32;  // frames per solve batch (one per lane)

struct OvmGroupLayout {
    size_t ref_off, idx_off, ring_off, sums_off, bar_off, total, stage_bytes;
};
// units: 4-atom units of a frame in memory; ref_units: units of the packed reference (== units without a selection);
// n_idx: length of the selection kept in shared memory (0 without)
__host__ __device__ inline OvmGroupLayout ovm_group_layout(int units, int ref_units, int n_idx, int fpi, int stages, int warps)
{
    OvmGroupLayout G;
    G.stage_bytes = (size_t)fpi * units * 48;
    G.ref_off = 0;
    G.idx_off = align_up((size_t)ref_units * 48, 16);
    G.ring_off = align_up(G.idx_off + (size_t)n_idx * 4, 128);
    G.sums_off = G.ring_off + (size_t)warps * stages * G.stage_bytes;
    G.bar_off = align_up(G.sums_off + (size_t)warps * kGroupBatch * kSumStride * sizeof(float), 8);
    G.total = G.bar_off + ((size_t)warps * stages + 1) * sizeof(uint64_t);
    return G;
}

template <int L, bool PRE>
__global__ void __launch_bounds__(kThreadsPerCta, 1) ovm_group_kernel(const OvmParams p)
{
    constexpr int FPI = 32 / L;  // frames per warp iteration
    extern __shared__ __align__(128) unsigned char smem[];
    const int units = p.total_units;           // units of a frame in memory
    const bool sel = !PRE && p.idx != nullptr;  // selection gathered out of the staged frame (never with traces)
    const int ref_units = sel ? (p.n_atoms + 3) / 4 : units;
    const int n_warps = blockDim.x >> 5;  // 16 for short frames, fewer when two ring stages of 32/L frames need more room
    const OvmGroupLayout G = ovm_group_layout(units, ref_units, sel ? p.n_atoms : 0, FPI, p.stages, n_warps);
    const float4* ref_s = reinterpret_cast<const float4*>(smem + G.ref_off);
    int* idx_s = reinterpret_cast<int*>(smem + G.idx_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G.bar_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / L, j = lane % L;
    unsigned char* ring = smem + G.ring_off + (size_t)warp * p.stages * G.stage_bytes;
    uint64_t* my_bars = bars + warp * p.stages;
    uint64_t* ref_bar = bars + n_warps * p.stages;
    float* sums = reinterpret_cast<float*>(smem + G.sums_off) + warp * kGroupBatch * kSumStride;
    const uint32_t frame_bytes = (uint32_t)units * 48u;
    // pivot (L >= 8, i.e. frames of more than 67 atoms): mean of 8 atoms spread over the frame (the selection), see
    // pivot_mean8; shorter frames keep their first atom -- their sums are too small for the pivot to matter.  No
    // mean-square shift here either: sum |x - p|^2 of <= 704 atoms is a few hundred.
    constexpr bool kMeanPivot = L >= 8 && OVM_PIVOT != 0;
    const int piv_k = kMeanPivot ? (int)(((int64_t)(2 * (j & 7) + 1) * p.n_atoms) >> 4) : 0;

    if (lane == 0)
        for (int s = 0; s < p.stages; ++s) mbar_init(&my_bars[s], 1);
    if (threadIdx.x == 0) mbar_init(ref_bar, 1);
    if (sel)
        for (int k = threadIdx.x; k < p.n_atoms; k += blockDim.x) idx_s[k] = __ldg(p.idx + k);
    fence_mbar_init();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(ref_bar, (uint32_t)ref_units * 48u);
        bulk_g2s(smem + G.ref_off, p.ref, (uint32_t)ref_units * 48u, ref_bar);
    }

    const int64_t W = (int64_t)gridDim.x * n_warps;
    const int64_t gw = (int64_t)blockIdx.x * n_warps + warp;
    const int64_t f_begin = p.n_frames * gw / W, f_end = p.n_frames * (gw + 1) / W;
    const int64_t n_iter = (f_end - f_begin + FPI - 1) / FPI;
    const uint64_t pol = l2_policy_evict_first();

    int64_t issued = 0;
    auto issue = [&](int stage) {  // lane 0: one bulk copy = FPI consecutive frames (contiguous in HBM)
        const int64_t fb = f_begin + issued * FPI;
        const int nfr = (int)min((int64_t)FPI, f_end - fb);
        const uint32_t bytes = (uint32_t)nfr * frame_bytes;
        mbar_arrive_expect_tx(&my_bars[stage], bytes);
        bulk_g2s_hint(ring + (size_t)stage * G.stage_bytes, p.xyz + fb * p.frame_stride, bytes, &my_bars[stage], pol);
        ++issued;
    };
    if (lane == 0)
        for (int s = 0; s < p.stages && issued < n_iter; ++s) issue(s);
    mbar_wait(ref_bar, 0);

    int stage = 0;
    uint32_t phase = 0;
    int slot = 0;
    int64_t batch_f0 = f_begin;
#pragma unroll 1
    for (int64_t it = 0; it < n_iter; ++it) {
        const int64_t fb = f_begin + it * FPI;
        const bool active = fb + g < f_end;
        mbar_wait(&my_bars[stage], phase);
        const float4* xs = reinterpret_cast<const float4*>(ring + (size_t)stage * G.stage_bytes + (size_t)g * frame_bytes);
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (!PRE) {
            // every lane of the warp takes part in the shuffles; inactive groups read frame 0 of the stage (harmless)
            const float* xf = reinterpret_cast<const float*>(active ? xs : reinterpret_cast<const float4*>(
                                                                              ring + (size_t)stage * G.stage_bytes));
            const int a0 = sel ? idx_s[piv_k] : piv_k;
            px = xf[3 * a0]; py = xf[3 * a0 + 1]; pz = xf[3 * a0 + 2];
            if (kMeanPivot) { px = pivot_mean8(px); py = pivot_mean8(py); pz = pivot_mean8(pz); }
        }
        if (active && sel) {
            // the whole frame is in shared memory: gather the selection from there
            const float* xf = reinterpret_cast<const float*>(xs);
            const float* yf = reinterpret_cast<const float*>(ref_s);
#pragma unroll 2
            for (int k = j; k < p.n_atoms; k += L) {
                const int a = idx_s[k];
                acc_atom<false>(v, xf[3 * a], xf[3 * a + 1], xf[3 * a + 2], yf[3 * k], yf[3 * k + 1], yf[3 * k + 2], px, py, pz);
            }
            if (j == 0) { v[13] = px; v[14] = py; v[15] = pz; }
        } else if (active) {
#pragma unroll 2
            for (int u = j; u < units; u += L) {
                const float4 a0 = xs[3 * u], a1 = xs[3 * u + 1], a2 = xs[3 * u + 2];
                const float4 b0 = ref_s[3 * u], b1 = ref_s[3 * u + 1], b2 = ref_s[3 * u + 2];
                acc_unit<PRE>(v, a0, a1, a2, b0, b1, b2, px, py, pz, p.n_atoms - 4 * u);
            }
            if (j == 0) { v[13] = px; v[14] = py; v[15] = pz; }
        }
        __syncwarp();
        if (lane == 0 && issued < n_iter) {
            fence_proxy_async_smem();
            issue(stage);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }

        // frames of <= 704 atoms: <= 44 atoms per lane and sums of a few hundred -- float32 throughout is already in the
        // 1e-6 nm class here; the float64 stages of the long-frame kernels would cost 5-15 % of HBM rate
        const int base = group_reduce_scatter16<L>(v, lane);
#pragma unroll
        for (int k = 0; k < 16 / L; ++k) sums[(slot + g) * kSumStride + base + k] = v[k];
        slot += FPI;
        if (slot == kGroupBatch || it + 1 == n_iter) {
            __syncwarp();
            const int64_t f = batch_f0 + lane;
            if (lane < slot && f < f_end) {
                double rec[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) rec[i] = (double)sums[lane * kSumStride + i];
                finish_frame<PRE>(rec, f, p, 0.f);
            }
            __syncwarp();
            slot = 0;
            batch_f0 = fb + FPI;
        }
    }
}

// combine segment partials (float64) and solve; one thread per frame
template <bool PRE>
__global__ void ovm_finish_kernel(const OvmParams p)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.n_frames) return;
    double rec[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) rec[i] = 0.0;
    for (int s = 0; s < p.n_seg; ++s) {
        const double2* src = reinterpret_cast<const double2*>(p.partials + ((size_t)f * p.n_seg + s) * 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double2 t = src[i];
            rec[2 * i] += t.x; rec[2 * i + 1] += t.y;
        }
    }
    // every segment carried the same pivot in slots 13..15
    const double inv = 1.0 / p.n_seg;
    rec[13] *= inv; rec[14] *= inv; rec[15] *= inv;
    finish_frame<PRE>(rec, f, p, PRE ? 0.f : msq_shift(p));
}

// ---------------------------------------------------------------------------
// atom_indices path: gather 12-byte atoms by index straight from the full frame.
// One warp per frame at a time, grid-stride over contiguous frame ranges.
// ---------------------------------------------------------------------------
template <bool PRE>
__global__ void __launch_bounds__(256) ovm_gather_kernel(const OvmParams p)
{
    __shared__ double sums_all[8 * kBatch * kSumStride];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sums = sums_all + warp * kBatch * kSumStride;
    const int64_t W = (int64_t)gridDim.x * 8;
    const int64_t gw = (int64_t)blockIdx.x * 8 + warp;
    const int64_t f_begin = p.n_frames * gw / W, f_end = p.n_frames * (gw + 1) / W;
    const int n = p.n_atoms;
    // pivot: mean of 32 selected atoms spread over the selection, one per lane (their sectors are read again below)
    const int piv_k = (int)(((int64_t)lane * n) >> 5);
    const int piv = p.idx ? __ldg(p.idx + piv_k) : piv_k;
    const float cshift = PRE ? 0.f : msq_shift(p);
    int slot = 0;
    int64_t batch_f0 = f_begin;

#pragma unroll 1
    for (int64_t f = f_begin; f < f_end; ++f) {
        const float* fr = p.xyz + f * p.frame_stride;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (!PRE) {
            px = warp_sum(__ldg(fr + 3 * piv)) * 0.03125f;
            py = warp_sum(__ldg(fr + 3 * piv + 1)) * 0.03125f;
            pz = warp_sum(__ldg(fr + 3 * piv + 2)) * 0.03125f;
        }
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
        double tot = 0.0;
#pragma unroll 1
        for (int k0 = 0; k0 < n; k0 += kFlushUnits * 64) {  // float32 lane partials over 512 atoms, then float64
            const int k1 = min(n, k0 + kFlushUnits * 64);
#pragma unroll 4
            for (int k = k0 + lane; k < k1; k += 32) {
                const int a = p.idx ? __ldg(p.idx + k) : k;
                const float ax = __ldg(fr + 3 * a), ay = __ldg(fr + 3 * a + 1), az = __ldg(fr + 3 * a + 2);
                const float bx = __ldg(p.ref + 3 * k), by = __ldg(p.ref + 3 * k + 1), bz = __ldg(p.ref + 3 * k + 2);
                acc_atom<PRE>(v, ax, ay, az, bx, by, bz, px, py, pz);
                if (!PRE) v[3] -= cshift;
            }
            if (k1 < n) {
                tot += warp_reduce_scatter16_mixed(v, lane);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
        }
        if (lane == 0) { v[13] = px; v[14] = py; v[15] = pz; }
        tot += warp_reduce_scatter16_mixed(v, lane);
        if (!(lane & 1)) sums[slot * kSumStride + (lane >> 1)] = tot;
        ++slot;
        if (slot == kBatch || f + 1 == f_end) {
            __syncwarp();
            if (lane < slot) {
                double rec[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) rec[i] = sums[lane * kSumStride + i];
                finish_frame<PRE>(rec, batch_f0 + lane, p, cshift);
            }
            __syncwarp();
            slot = 0;
            batch_f0 = f + 1;
        }
    }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
cudaError_t launch_ovm_tma(const OvmParams& p, bool precentered, int sm_count, cudaStream_t st)
{
    if (p.n_frames <= 0) return cudaSuccess;
    const size_t smem = ovm_tma_smem_bytes(p);
    const bool flush = p.seg_units > kFlushUnits * 32;
    auto kern = precentered ? (flush ? ovm_tma_kernel<true, true> : ovm_tma_kernel<true, false>)
                            : (flush ? ovm_tma_kernel<false, true> : ovm_tma_kernel<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int ctas_per_seg = sm_count / p.n_seg;
    if (ctas_per_seg < 1) ctas_per_seg = 1;
    // do not launch more warps than frames
    const int64_t need = (p.n_frames + kWarpsPerCta - 1) / kWarpsPerCta;
    if ((int64_t)ctas_per_seg > need) ctas_per_seg = (int)need;
    kern<<<p.n_seg * ctas_per_seg, kThreadsPerCta, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (p.n_seg > 1) {
        auto fin = precentered ? ovm_finish_kernel<true> : ovm_finish_kernel<false>;
        const int threads = 128;
        fin<<<(unsigned)((p.n_frames + threads - 1) / threads), threads, 0, st>>>(p);
        e = cudaGetLastError();
    }
    return e;
}

template <int L>
static cudaError_t launch_group_L(const OvmParams& p, bool precentered, int sm_count, int warps, cudaStream_t st)
{
    const bool sel = !precentered && p.idx != nullptr;
    const OvmGroupLayout G = ovm_group_layout(p.total_units, sel ? (p.n_atoms + 3) / 4 : p.total_units, sel ? p.n_atoms : 0,
                                              32 / L, p.stages, warps);
    auto kern = precentered ? ovm_group_kernel<L, true> : ovm_group_kernel<L, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.total);
    if (e != cudaSuccess) return e;
    int64_t ctas = sm_count;
    const int64_t need = (p.n_frames + warps * (32 / L) - 1) / (warps * (32 / L));
    if (ctas > need) ctas = need;
    kern<<<(unsigned)ctas, warps * 32, G.total, st>>>(p);
    return cudaGetLastError();
}

// Short- and mid-size-frame path.  Returns false (and launches nothing) when the shape is not eligible.
//   frames <= 3 KB (256 atoms): 16 warps, the fewest lanes per frame (= most frames per bulk copy and per reduction) whose
//     per-warp ring still holds >= 2 stages;
//   frames <= 8.25 KB (704 atoms): 16 lanes per frame (two frames per warp pass, 94 % of the lanes busy at N = 300 against
//     78 % for a warp per frame, one reduction and one copy per two frames) on as many warps (15 ... 6) as leave every
//     warp a two-stage ring of frame pairs: N = 300 0.67x -> 0.98x of HBM peak, N = 516 0.77x -> 1.01x.
bool launch_ovm_group(OvmParams& p, bool precentered, int sm_count, cudaStream_t st, cudaError_t* err)
{
    const bool sel = p.idx != nullptr;
    if (sel && (precentered || p.frame_atoms <= 0)) return false;
    p.total_units = ((sel ? p.frame_atoms : p.n_atoms) + 3) / 4;
    const size_t frame_bytes = (size_t)p.total_units * 48;
    // a selection is gathered out of the staged frame when that costs no more HBM traffic than gathering from global
    // memory would (every selected atom touches a 32-byte sector or two) or when the frames are so small that the
    // gather kernel's warp per frame is the limit (2.0e9 frames/s whatever N)
    if (sel && frame_bytes > 3072 && (size_t)p.n_atoms * 64 < frame_bytes) return false;
    const size_t sel_bytes = sel ? align_up((size_t)((p.n_atoms + 3) / 4) * 48, 16) + (size_t)p.n_atoms * 4 + 128 : 0;
    size_t max_bytes = 8448;  // 704 atoms: measured 0.97-1.05x of HBM peak up to here, the chunked kernel wins from ~800 atoms
#ifdef B200RMSD_DEV_SWITCHES
    if (const char* mb = getenv("B200RMSD_GROUP_MAX_BYTES")) max_bytes = (size_t)atol(mb);  // development override
#endif
    if (p.frame_stride != (int64_t)p.total_units * 12 || frame_bytes > max_bytes || p.n_seg > 1) return false;
    const size_t budget = 232448;
    auto per_warp_bytes = [&](int warps) {
        const size_t fixed = (sel ? sel_bytes : align_up(frame_bytes, 128)) +
                             (size_t)warps * kGroupBatch * kSumStride * sizeof(float) + 1024;
        return fixed < budget ? (budget - fixed) / warps : (size_t)0;
    };
    int Lsel = 0, stages = 0, warps = kWarpsPerCta;
    for (int L = 2; L <= 16; L <<= 1) {
        const int st_n = (int)(per_warp_bytes(warps) / ((size_t)(32 / L) * frame_bytes));
        if (st_n >= 2) { Lsel = L; stages = st_n > 4 ? 4 : st_n; break; }
    }
    if (Lsel == 0) {
        for (warps = kWarpsPerCta - 1; warps >= 4; --warps)
            if (per_warp_bytes(warps) / (2 * frame_bytes) >= 2) { Lsel = 16; stages = 2; break; }
    }
#ifdef B200RMSD_DEV_SWITCHES
    if (const char* force = getenv("B200RMSD_GROUP_LANES")) {  // development override
        const int L = atoi(force);
        if (L == 2 || L == 4 || L == 8 || L == 16) {
            const int st_n = (int)(per_warp_bytes(kWarpsPerCta) / ((size_t)(32 / L) * frame_bytes));
            if (st_n >= 2) { Lsel = L; stages = st_n > 4 ? 4 : st_n; warps = kWarpsPerCta; }
        }
    }
    if (const char* force = getenv("B200RMSD_GROUP_WARPS")) {  // development override: warps per CTA for L = 16
        const int w = atoi(force);
        if (w >= 4 && w <= kWarpsPerCta && per_warp_bytes(w) / (2 * frame_bytes) >= 2) {
            Lsel = 16; warps = w;
            stages = (int)std::min<size_t>(4, per_warp_bytes(w) / (2 * frame_bytes));
        }
    }
#endif
    if (Lsel == 0) return false;
    p.stages = stages;
    switch (Lsel) {
        case 2: *err = launch_group_L<2>(p, precentered, sm_count, warps, st); break;
        case 4: *err = launch_group_L<4>(p, precentered, sm_count, warps, st); break;
        case 8: *err = launch_group_L<8>(p, precentered, sm_count, warps, st); break;
        default: *err = launch_group_L<16>(p, precentered, sm_count, warps, st); break;
    }
    return true;
}

cudaError_t launch_ovm_gather(const OvmParams& p, bool precentered, int sm_count, cudaStream_t st)
{
    if (p.n_frames <= 0) return cudaSuccess;
    auto kern = precentered ? ovm_gather_kernel<true> : ovm_gather_kernel<false>;
    int64_t ctas = (int64_t)sm_count * 8;  // 8 x 256 threads = full occupancy
    const int64_t need = (p.n_frames + 7) / 8;
    if (ctas > need) ctas = need;
    kern<<<(unsigned)ctas, 256, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace b200
