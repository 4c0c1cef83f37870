// allpairs_tc.cu -- all-pairs RMSD matrix on the 5th-generation tensor cores.
//
// D[i][j] = rmsd(frame j onto frame i) needs the 3x3 inner products M_ij = X_i X_j^T of the centred
// frames: one dense contraction (3F x A).(A x 3F).  This kernel computes it as a tcgen05 GEMM:
//
//   operands   K-major fp32 rows (one row per frame component, K = atoms padded to 32), pre-split into
//              tf32 "hi" and "lo" parts by allpairs_tc_prepare_kernel (hi = rna_tf32(x), lo = rna_tf32(x-hi));
//              rows are grouped 10 frames (30 rows) + 2 zero rows per 32, so that the three rows of a frame
//              always sit in one warp's TMEM lane quarter;
//   loads      TMA tiled copies (cp.async.bulk.tensor.2d, SWIZZLE_128B, 128 rows x 32 floats per box) into a
//              3-stage shared-memory ring, mbarrier full/empty pipeline;
//   MMA        one elected thread issues tcgen05.mma.cta_group::1.kind::tf32, M=128 N=128 K=8, three per
//              K-step (lo.hi, hi.lo, hi.hi: "3xTF32", the dropped lo.lo term is ~2^-22 relative) accumulating
//              fp32 in TMEM; tcgen05.commit releases smem stages and publishes finished accumulators;
//   epilogue   8 warps read the accumulator with tcgen05.ld.32x32b (one TMEM lane = one row per thread),
//              regroup 3x3 blocks with warp shuffles, run the QCP solve and write only the RMSD (4 bytes per
//              pair); TMEM is double buffered so the epilogue of tile t overlaps the MMAs of tile t+1.
//
// The inner products never touch HBM.  Replaces the Python loop of F md.rmsd calls
// (examples/clustering.ipynb:78-81); arithmetic of each pair == msdFromMandG (theobald_rmsd.cpp:217-334).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "../../include/b200rmsd.h"
#include "allpairs_layout.cuh"
#include "common.cuh"
#include "kernels.cuh"
#include "qcp.cuh"
#include "tc_ptx.cuh"

namespace b200 {

constexpr int kBM = 128, kBN = 128, kBK = 32, kStages = 3;
constexpr int kFramesPerTile = 40;               // 4 warps x 10 frames
constexpr int kOperandBytes = kBM * kBK * 4;     // 16 KB: one 128 x 32 fp32 box
constexpr int kStageBytes = 4 * kOperandBytes;   // A_hi, A_lo, B_hi, B_lo
constexpr int kAccCols = kBN;                    // fp32 accumulator columns per stage
constexpr uint32_t kTmemCols = 256;              // 2 accumulator stages

// ---------------------------------------------------------------------------------------------
// prepare: centre every frame (center_generic.h:3-44 semantics), write tf32 hi/lo rows + traces.
// one warp per frame.  Row of (frame f, component c) = 32*(f/10) + 3*(f%10) + c.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) allpairs_tc_prepare_kernel(const float* __restrict__ xyz, int64_t n_frames,
                                                                  int64_t frame_stride, const int* __restrict__ idx,
                                                                  int n_sel, int k_pad, float* __restrict__ hi,
                                                                  float* __restrict__ lo, float* __restrict__ traces)
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
        const int64_t row = ap_tc_row(f, 0);
        double tr = 0;
        for (int k = lane; k < k_pad; k += 32) {
            float v[3] = {0.f, 0.f, 0.f};
            if (k < n_sel) {
                const int a = idx ? __ldg(idx + k) : k;
                v[0] = __ldg(fr + 3 * a) - mx; v[1] = __ldg(fr + 3 * a + 1) - my; v[2] = __ldg(fr + 3 * a + 2) - mz;
                tr += (double)(v[0] * v[0]); tr += (double)(v[1] * v[1]); tr += (double)(v[2] * v[2]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float h = rna_tf32(v[c]);
                hi[(row + c) * k_pad + k] = h;
                lo[(row + c) * k_pad + k] = rna_tf32(v[c] - h);
            }
        }
        tr = warp_sum(tr);
        if (lane == 0) traces[f] = (float)tr;
    }
}

// ---------------------------------------------------------------------------------------------
struct TcParams {
    const float* traces;
    float* out;
    int64_t ld;
    int64_t n_frames;
    int64_t row0, row1;      // output rows [row0, row1)
    int64_t col0, col1;      // output columns [col0, col1)
    float* out_t;            // optional transposed copy of the block: out_t[(j-col0)*ld_t + (i-row0)]
    int64_t ld_t;
    int tiles_j0;            // first j-tile (= col0 / 40)
    int n_sel;
    int nk;                  // K blocks of 32
    int tiles_i0;            // first i-tile (= row0 / 40)
    int tiles_i, tiles_j;    // tile grid
    int symmetric;           // full square matrix: compute tiles tj >= ti only and mirror them
    int n_super;             // super-blocks per side (symmetric mode)
    unsigned flags;
};

// Tile rasterisation: super-blocks of kSuper x kSuper tiles, row-major inside a block and across blocks, so that the
// ~150 CTAs in flight share a working set of 2*kSuper operand tiles (10 MB at K=320) that stays L2-resident,
// instead of sweeping the whole operand (hundreds of MB) once per tile row.
constexpr int kSuper = 16;
__device__ __forceinline__ bool tile_coords(int64_t t, const TcParams& p, int& ti, int& tj)
{
    if (p.symmetric) {
        // slots: (super-block pair bi <= bj) x (kSuper x kSuper tiles); slots outside the matrix or below the
        // diagonal are skipped (identically by all three warp roles)
        const int64_t blk = t / (kSuper * kSuper);
        const int in = (int)(t % (kSuper * kSuper));
        const double nb2 = 2.0 * p.n_super + 1.0;
        int bi = (int)((nb2 - sqrt(nb2 * nb2 - 8.0 * (double)blk)) * 0.5);
        // first slot of block row bi is bi*n_super - bi*(bi-1)/2; fix rounding of the square root
        while (bi > 0 && (int64_t)bi * p.n_super - (int64_t)bi * (bi - 1) / 2 > blk) --bi;
        while ((int64_t)(bi + 1) * p.n_super - (int64_t)(bi + 1) * bi / 2 <= blk) ++bi;
        const int bj = bi + (int)(blk - ((int64_t)bi * p.n_super - (int64_t)bi * (bi - 1) / 2));
        const int li = bi * kSuper + in / kSuper, lj = bj * kSuper + in % kSuper;  // tile offsets inside the square block
        ti = p.tiles_i0 + li;
        tj = p.tiles_i0 + lj;
        return li < p.tiles_i && lj < p.tiles_i && lj >= li;
    }
    const int bj_count = (p.tiles_j + kSuper - 1) / kSuper;
    const int64_t full_row = (int64_t)kSuper * p.tiles_j;          // tiles in one full block row
    const int bi = (int)(t / full_row);
    const int rows_here = min(kSuper, p.tiles_i - bi * kSuper);
    int64_t r = t - (int64_t)bi * full_row;                        // index inside this block row
    const int64_t per_block = (int64_t)rows_here * kSuper;
    int bj = (int)(r / per_block);
    if (bj >= bj_count) bj = bj_count - 1;
    r -= (int64_t)bj * per_block;
    const int cols_here = min(kSuper, p.tiles_j - bj * kSuper);
    ti = p.tiles_i0 + bi * kSuper + (int)(r / cols_here);
    tj = p.tiles_j0 + bj * kSuper + (int)(r % cols_here);
    return true;
}

__device__ __forceinline__ float sel3(int c, float a0, float a1, float a2) { return c == 0 ? a0 : (c == 1 ? a1 : a2); }

// EPI_WARPS in {4, 8, 16}: epilogue warps (each TMEM lane quarter is served by EPI_WARPS/4 warps that split the
// four 32-column chunks of the accumulator); NP in {1, 2}: independent solves interleaved per lane.
template <int EPI_WARPS, int NP>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
allpairs_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const TcParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment is required by SWIZZLE_128B; dynamic smem base is not guaranteed to have it
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty = full + kStages;
    uint64_t* tfull = empty + kStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = p.symmetric ? (int64_t)p.n_super * (p.n_super + 1) / 2 * (kSuper * kSuper)
                                        : (int64_t)p.tiles_i * p.tiles_j;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], EPI_WARPS); }
        fence_mbar_init();
    }
    if (warp == EPI_WARPS) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Warp roles: warps [0, EPI_WARPS) epilogue, warp EPI_WARPS = TMA producer, warp EPI_WARPS+1 = MMA issuer.
    // The two single-thread roles get the highest warp ids: the SMSP arbiter favours higher warp ids, and a late
    // TMA or MMA issue stalls the whole pipeline while a late epilogue instruction does not.
    if (warp == EPI_WARPS) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int ti, tj;
                if (!tile_coords(t, p, ti, tj)) continue;
                for (int kb = 0; kb < p.nk; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    unsigned char* st = smem + stage * kStageBytes;
                    mbar_arrive_expect_tx(&full[stage], kStageBytes);
                    tma_load_2d(st, &map_hi, &full[stage], kb * kBK, ti * kBM);
                    tma_load_2d(st + kOperandBytes, &map_lo, &full[stage], kb * kBK, ti * kBM);
                    tma_load_2d(st + 2 * kOperandBytes, &map_hi, &full[stage], kb * kBK, tj * kBN);
                    tma_load_2d(st + 3 * kOperandBytes, &map_lo, &full[stage], kb * kBK, tj * kBN);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_tf32_idesc(kBM, kBN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int ti_unused, tj_unused;
                if (!tile_coords(t, p, ti_unused, tj_unused)) continue;
                mbar_wait(&tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
                for (int kb = 0; kb < p.nk; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    unsigned char* st = smem + stage * kStageBytes;
                    const uint64_t a_hi = make_sw128_kmajor_desc(st), a_lo = make_sw128_kmajor_desc(st + kOperandBytes);
                    const uint64_t b_hi = make_sw128_kmajor_desc(st + 2 * kOperandBytes),
                                   b_lo = make_sw128_kmajor_desc(st + 3 * kOperandBytes);
#pragma unroll
                    for (int ks = 0; ks < ((p.flags & 0x200u) ? 0 : kBK / 8); ++ks) {  // 0x200: development, skip MMAs
                        const uint64_t off = (uint64_t)(ks * 2);  // 8 floats = 32 bytes = 2 x 16-byte units
                        umma_tf32(d_tmem, a_lo + off, b_hi + off, idesc, (kb | ks) != 0 ? 1u : 0u);
                        umma_tf32(d_tmem, a_hi + off, b_lo + off, idesc, 1u);
                        umma_tf32(d_tmem, a_hi + off, b_hi + off, idesc, 1u);
                    }
                    umma_commit(&empty[stage]);  // frees this smem stage when the MMAs above have read it
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&tfull[acc]);        // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================================================== epilogue (TMEM -> QCP -> HBM)
        const int ew = warp & 3;                 // TMEM lane quarter this warp may read (warp_id % 4)
        constexpr int kChunksPerWarp = 16 / EPI_WARPS;  // 4, 2 or 1 of the four 32-column chunks
        const int part = warp >> 2;              // which share of the chunks
        const int c = lane % 3, tq = lane / 3;   // component row and frame slot of this lane
        const bool row_valid = lane < 30;
        const int src1 = lane - c + (c + 1) % 3, src2 = lane - c + (c + 2) % 3;
        const float inv_n = 1.0f / (float)p.n_sel;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int ti, tj;
            if (!tile_coords(t, p, ti, tj)) continue;
            const int64_t fi = (int64_t)ti * kFramesPerTile + ew * 10 + tq;  // row frame of this lane
            const bool i_ok = row_valid && fi >= p.row0 && fi < p.row1;
            const float Gi = i_ok ? __ldg(p.traces + fi) : 1.0f;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < kChunksPerWarp; ++cc) {
                const int chunk = part * kChunksPerWarp + cc;
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * kAccCols + chunk * 32), r);
#pragma unroll
                for (int jp = 0; jp < 4 / NP; ++jp) {
                    // NP j-frame groups per pass so that NP independent solves interleave
                    float M[NP][9], Ga[NP], Gb[NP], res[NP];
                    int64_t fj[NP];
                    bool ok[NP], trusted[NP];
                    // rows (c, c+1, c+2) mod 3 of the 3x3 block of pair (i_t, j-frame 3*jg + c): a cyclic permutation of
                    // x,y,z = a proper rotation of frame i, which leaves the RMSD unchanged.  Warp-collective (shuffles).
                    auto gather = [&](int jg, float (&m)[9]) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const float m0 = __uint_as_float(r[(jg < 3 ? 9 * jg : 27) + q]);
                            const float m1 = __uint_as_float(r[(jg < 3 ? 9 * jg + 3 : 27) + q]);
                            const float m2 = __uint_as_float(r[(jg < 3 ? 9 * jg + 6 : 27) + q]);
                            m[q] = sel3(c, m0, m1, m2);
                            const float send1 = sel3(c, m2, m0, m1);  // for the reader whose component is (c+2)%3
                            const float send2 = sel3(c, m1, m2, m0);  // for the reader whose component is (c+1)%3
                            m[3 + q] = __shfl_sync(0xffffffffu, send1, src1);
                            m[6 + q] = __shfl_sync(0xffffffffu, send2, src2);
                        }
                    };
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        const int jg = NP * jp + u;  // group of three j-frames (the last group holds one frame only)
                        gather(jg, M[u]);
                        const int jl = 3 * jg + c;  // j-frame slot inside the chunk
                        fj[u] = (int64_t)tj * kFramesPerTile + chunk * 10 + jl;
                        ok[u] = i_ok && jl < 10 && fj[u] >= p.col0 && fj[u] < p.col1;
                        Ga[u] = ok[u] ? __ldg(p.traces + fj[u]) : 1.0f;
                        Gb[u] = Gi;
                        trusted[u] = true;
                    }
                    if (p.flags & 0x100u) {  // development: skip the solve to time the GEMM main loop alone
#pragma unroll
                        for (int u = 0; u < NP; ++u) res[u] = M[u][0];
                    } else if (p.flags & B200RMSD_FAST_SOLVE) {  // all-float32 solve (reference-class precision)
                        qcp_msd_f32<NP>(M, Ga, Gb, ok, inv_n, res, trusted);
                    } else {
                        qcp_msd_fast<NP>(M, Ga, Gb, ok, inv_n, res, trusted);
                    }
                    bool all_trusted = true;
#pragma unroll
                    for (int u = 0; u < NP; ++u) all_trusted = all_trusted && trusted[u];
                    if (!__all_sync(0xffffffffu, all_trusted)) {
                        // rare (collinear atoms, two-atom selections: a double largest root): fetch the blocks again --
                        // M is dead by now, which keeps it out of the solver's register budget -- and take the closed form
#pragma unroll
                        for (int u = 0; u < NP; ++u) {  // unrolled: the accumulator registers must stay statically indexed
                            float m[9];
                            gather(NP * jp + u, m);
                            if (!trusted[u]) res[u] = qcp_rmsd_closed(m, Ga[u], Gb[u], inv_n);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < NP; ++u)
                        if (ok[u]) {
                            const float v = (fi == fj[u] && (p.flags & B200RMSD_DIAG_ZERO)) ? 0.f : res[u];
                            if (!p.symmetric) {
                                p.out[(size_t)(fi - p.row0) * p.ld + fj[u]] = v;
                                if (p.out_t) p.out_t[(size_t)(fj[u] - p.col0) * p.ld_t + (fi - p.row0)] = v;
                            } else if (fj[u] >= fi) {  // each unordered pair once, mirrored: D is exactly symmetric
                                p.out[(size_t)(fi - p.row0) * p.ld + fj[u]] = v;
                                if (fj[u] != fi) p.out[(size_t)(fj[u] - p.row0) * p.ld + fi] = v;
                            }
                        }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
cudaError_t launch_allpairs_tc_prepare(const float* xyz, int64_t n_frames, int64_t frame_stride, const int* idx, int n_sel,
                                       int k_pad, float* hi, float* lo, float* traces, int64_t rows_pad, int sm_count,
                                       cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(hi, 0, (size_t)rows_pad * k_pad * 4, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(lo, 0, (size_t)rows_pad * k_pad * 4, st);
    if (e != cudaSuccess) return e;
    int64_t ctas = (int64_t)sm_count * 8;
    const int64_t need = (n_frames + 7) / 8;
    if (ctas > need) ctas = need;
    allpairs_tc_prepare_kernel<<<(unsigned)ctas, 256, 0, st>>>(xyz, n_frames, frame_stride, idx, n_sel, k_pad, hi, lo, traces);
    return cudaGetLastError();
}

int launch_allpairs_tc_block(const float* hi, const float* lo, const float* traces, int64_t n_frames, int n_sel, int k_pad,
                             int64_t rows_pad, int64_t row0, int64_t row1, int64_t col0, int64_t col1, float* out,
                             int64_t ld, float* out_t, int64_t ld_t, unsigned flags, int sm_count, cudaStream_t st)
{
    CUtensorMap map_hi, map_lo;
    if (!make_operand_map(&map_hi, hi, rows_pad, k_pad, kBM) || !make_operand_map(&map_lo, lo, rows_pad, k_pad, kBM))
        return set_error(B200RMSD_ECUDA, "allpairs: cuTensorMapEncodeTiled failed");
    TcParams p{};
    p.traces = traces;
    p.out = out;
    p.ld = ld;
    p.n_frames = n_frames;
    p.row0 = row0;
    p.row1 = row1;
    p.col0 = col0;
    p.col1 = col1;
    p.out_t = out_t;
    p.ld_t = ld_t;
    p.n_sel = n_sel;
    p.nk = k_pad / kBK;
    p.tiles_i0 = (int)(row0 / kFramesPerTile);
    p.tiles_i = (int)((row1 + kFramesPerTile - 1) / kFramesPerTile) - p.tiles_i0;
    p.tiles_j0 = (int)(col0 / kFramesPerTile);
    p.tiles_j = (int)((col1 + kFramesPerTile - 1) / kFramesPerTile) - p.tiles_j0;
    p.flags = flags;
    // a square block on the diagonal: compute tiles tj >= ti only and mirror them (exactly symmetric, half the flops)
    p.symmetric = (row0 == col0 && row1 == col1 && !out_t && !getenv("B200RMSD_NO_SYMMETRIC")) ? 1 : 0;
    p.n_super = (p.tiles_i + kSuper - 1) / kSuper;
    if (const char* dbg = getenv("B200RMSD_TC_DEBUG")) p.flags |= (unsigned)strtoul(dbg, nullptr, 0) & 0xff02u;
    const size_t smem = (size_t)kStages * kStageBytes + 1024 + 256;
    int64_t ctas = sm_count;
    const int64_t n_tiles = p.symmetric ? ((int64_t)p.tiles_i * (p.tiles_i + 1)) / 2 : (int64_t)p.tiles_i * p.tiles_j;
    if (ctas > n_tiles) ctas = n_tiles;
    if (ctas < 1) ctas = 1;
    const char* cfg = getenv("B200RMSD_TC_EPILOGUE");  // development: "<warps>x<np>", e.g. 16x1
    int ew = 16, np = 2;  // measured best on B200 (16x2 > 16x1 > 8x2 > 8x1)
    if (cfg) sscanf(cfg, "%dx%d", &ew, &np);
    cudaError_t e = cudaSuccess;
#define B200_LAUNCH_TC(EW, NP)                                                                                         \
    do {                                                                                                               \
        e = cudaFuncSetAttribute(allpairs_tc_kernel<EW, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) allpairs_tc_kernel<EW, NP><<<(unsigned)ctas, 64 + 32 * EW, smem, st>>>(map_hi, map_lo, p); \
    } while (0)
    if (ew == 4 && np == 2) B200_LAUNCH_TC(4, 2);
    else if (ew == 8 && np == 1) B200_LAUNCH_TC(8, 1);
    else if (ew == 8 && np == 2) B200_LAUNCH_TC(8, 2);
    else if (ew == 16 && np == 1) B200_LAUNCH_TC(16, 1);
    else B200_LAUNCH_TC(16, 2);
#undef B200_LAUNCH_TC
    if (e != cudaSuccess) return set_error(B200RMSD_ECUDA, "allpairs: %s", cudaGetErrorString(e));
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error(B200RMSD_ECUDA, "allpairs: %s", cudaGetErrorString(e));
}

}  // namespace b200
