// allpairs_tc144.cu -- all-pairs RMSD matrix on the 5th-generation tensor cores, dense-row layout.
//
// D[i][j] = rmsd(frame j onto frame i) needs the 3x3 inner products M_ij = X_i X_j^T of the centred
// frames: one dense contraction (3F x A).(A x 3F), computed as a tcgen05 GEMM whose epilogue solves the
// QCP polynomial of every 3x3 block and writes only the RMSD (arithmetic of each pair == msdFromMandG,
// theobald_rmsd.cpp:217-334; replaces the Python loop of F md.rmsd calls, examples/clustering.ipynb:78-81).
//
//   operands   K-major fp32 matrices in tf32 "hi" and "lo" parts (hi = rna_tf32(x), lo = rna_tf32(x-hi)), row 3f+c =
//              component c of frame f, no padding rows.  A = every frame aligned onto its nearest reference structure
//              (allpairs_refs.cu), B = its difference from that reference; eight extra K columns per reference, in
//              matrices of their own, add X'_i c_r^T to the blocks whose column frame refers to c_r, so that the
//              accumulator holds X'_i D_j^T (fluctuation-sized) until the last K-steps -- the tensor core's fp32
//              accumulation truncates, and the bias is proportional to the running sum (allpairs_tc144_prepare_kernel);
//   tile       40 i-frames x 48 j-frames.  The A operand (M = 128 TMEM lanes) is four 32-row quarters starting at rows
//              120*ti + 30*w, so that each epilogue warp's lane quarter holds 10 whole frames (lanes 30 and 31 of a
//              quarter are zero-filled and ignored); the B operand (N = 144
//              accumulator columns) is one 144-row box = 48 whole frames.  Columns are registers of the reading
//              thread, so frames may sit at any column: each of the 16 epilogue warps takes 36 columns = 12 j-frames
//              = 4 passes of 3 frames per lane triplet, every pass full (the 128-column layout of allpairs_tc.cu
//              padded every 32 columns to 10 frames and left the fourth pass two-thirds empty: 1600 pairs per tile
//              for the same epilogue passes that now deliver 1920);
//   loads      two cp.async.bulk.tensor copies per K block, SWIZZLE_128B: the A box (4-d: 32 floats x 32 rows x 4 quarters
//              x {hi, lo}, see make_a_operand_map) and the B box (3-d: 32 floats x 144 rows x {hi, lo}); 3-stage
//              shared-memory ring (68 KB per stage), mbarrier full/empty pipeline;
//   MMA        one elected thread issues tcgen05.mma.cta_group::1.kind::tf32, M=128 N=144 K=8, three per K-step
//              (lo.hi, hi.lo, hi.hi: "3xTF32", the dropped lo.lo term is ~2^-22 relative) accumulating fp32 in TMEM;
//              tcgen05.commit releases smem stages and publishes finished accumulators;
//   epilogue   tcgen05.ld.32x32b (one TMEM lane = one row per thread), 3x3 blocks regrouped with warp shuffles,
//              QCP solve, 4 bytes per pair written; two accumulator stages in TMEM so that the epilogue of tile t
//              overlaps the MMAs of tile t+1.
//
// The inner products never touch HBM.
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
namespace {

constexpr int kM = 128, kN = 144, kK = 32, kRing = 3;
constexpr int kIFrames = 40;                    // 4 lane quarters x 10 frames
constexpr int kJFrames = 48;                    // 144 columns
constexpr int kQuarterBytes = 32 * kK * 4;      // 4 KB: one 32-row A box
constexpr int kABytes = kM * kK * 4;            // 16 KB
constexpr int kBBytes = kN * kK * 4;            // 18 KB
constexpr int kStage = 2 * kABytes + 2 * kBBytes;  // A_hi, A_lo, B_hi, B_lo = 68 KB (a multiple of 1024)
constexpr int kAccStages = 3;                   // accumulator stages in TMEM: the epilogue warps of a tile finish at different
                                                // times (Newton trip counts, the rare fall-back pass); with two stages the
                                                // fastest warp could run at most one tile ahead of the slowest
constexpr uint32_t kAccStride = 160;            // TMEM columns between accumulator stages (144 used; 3 x 160 <= 512)
constexpr uint32_t kTmem = 512;                 // allocation (power of two >= 256 + 144)
constexpr int kSegCols = 36;                    // columns per epilogue segment = 12 j-frames
// CTA-pair mode (cta_group::2): each CTA of a 2-CTA cluster holds its own 128 rows of A and HALF of the B box (72 rows);
// one M = 256 MMA issued by the leader feeds both accumulators.  A stage shrinks to 50 KB, the ring grows to 4 stages,
// and the L2 -> SM traffic per tile falls from 272 to 200 operand rows.
constexpr int kRingPair = 4;
constexpr int kBHalfBytes = (kN / 2) * kK * 4;  // 9 KB
constexpr int kStagePair = 2 * kABytes + 2 * kBHalfBytes;  // 50 KB (a multiple of 1024)
// Tile rasterisation: super-blocks of 24 x 20 tiles = 960 x 960 frames, row-major inside a block and across blocks, so
// that the ~150 CTAs in flight share a working set of 2 x 960 frames (15 MB of operands at K = 320) that stays
// L2-resident instead of sweeping the whole operand once per tile row.
constexpr int kSupI = 24, kSupJ = 20;

struct Tc144Params {
    const float* traces;
    float* out;
    int64_t ld;
    int64_t row0, row1;      // output rows [row0, row1)
    int64_t col0, col1;      // output columns [col0, col1)
    float* out_t;            // optional transposed copy of the block: out_t[(j-col0)*ld_t + (i-row0)]
    int64_t ld_t;
    float* out_rot;          // optional (ROT kernels): 9 floats per pair, out_rot[((i-row0)*(col1-col0) + (j-col0))*9 + 3a+b]
    const float* frame_rot;  // the rotation every frame got in the prepare step (x' = (x - centroid) R_f), 9 floats each
    int64_t n_slots;         // super-block slots to walk (tiles outside the block or under the diagonal are skipped)
    int tiles_i0, tiles_j0;  // first i-tile (= row0 / 40) and j-tile (= col0 / 48)
    int tiles_i, tiles_j;    // tile grid
    int n_bj;                // super-blocks per block row
    int n_sel;
    int nk;                  // K blocks of 32 (atoms)
    const int2* tile_aug;    // per absolute j-tile: first and last augmentation K block (first > last: none)
    int pair;                // CTA-pair mode: a slot is TWO i-tiles (ti, ti + 1) x one j-tile, one per CTA of the cluster
    int symmetric;           // square block on the diagonal: tiles that hold no pair j >= i are skipped, values mirrored
    unsigned flags;
};

// slot -> tile; false for slots outside the tile grid and, in symmetric mode, for tiles entirely under the diagonal.
// Evaluated identically by the three warp roles.
__host__ __device__ __forceinline__ bool tile_of_slot(int64_t t, const Tc144Params& p, int& ti, int& tj)
{
    const int sup_i = p.pair ? kSupI / 2 : kSupI;  // rows of slots per super-block (a pair slot is two tiles high)
    const int64_t blk = t / (sup_i * kSupJ);
    const int in = (int)(t - blk * (sup_i * kSupJ));
    const int bi = (int)(blk / p.n_bj), bj = (int)(blk - (int64_t)bi * p.n_bj);
    const int li = (bi * sup_i + in / kSupJ) * (p.pair ? 2 : 1), lj = bj * kSupJ + in % kSupJ;
    ti = p.tiles_i0 + li;  // pair mode: the leader's tile; the peer takes ti + 1
    tj = p.tiles_j0 + lj;
    if (li >= p.tiles_i || lj >= p.tiles_j) return false;
    // symmetric: the tile is needed iff its last j-frame is not before its first i-frame
    return !p.symmetric || (int64_t)tj * kJFrames + (kJFrames - 1) >= (int64_t)ti * kIFrames;
}

// The same walk without the 64-bit divisions: every thread of the three warp roles visits every slot (skipped ones
// included, about half of them in symmetric mode), and t / (sup_i * kSupJ), blk / n_bj were 6 % of the kernel's stall
// samples.  The step between a CTA's slots is smaller than a super-block, so the position advances by carries.
struct SlotWalk {
    int64_t t;
    int in, bi, bj, per_block, step;
    __device__ __forceinline__ SlotWalk(int64_t t0, int step_, const Tc144Params& p) : t(t0), step(step_)
    {
        per_block = (p.pair ? kSupI / 2 : kSupI) * kSupJ;
        const int64_t blk = t0 / per_block;
        in = (int)(t0 - blk * per_block);
        bi = (int)(blk / p.n_bj);
        bj = (int)(blk - (int64_t)bi * p.n_bj);
    }
    __device__ __forceinline__ void next(const Tc144Params& p)
    {
        t += step;
        in += step;
        while (in >= per_block) {
            in -= per_block;
            if (++bj == p.n_bj) { bj = 0; ++bi; }
        }
    }
    __device__ __forceinline__ bool tile(const Tc144Params& p, int& ti, int& tj) const
    {
        const int row = in / kSupJ, col = in - row * kSupJ;
        const int li = (bi * (p.pair ? kSupI / 2 : kSupI) + row) * (p.pair ? 2 : 1), lj = bj * kSupJ + col;
        ti = p.tiles_i0 + li;
        tj = p.tiles_j0 + lj;
        if (li >= p.tiles_i || lj >= p.tiles_j) return false;
        return !p.symmetric || (int64_t)tj * kJFrames + (kJFrames - 1) >= (int64_t)ti * kIFrames;
    }
};

// every wait of this kernel is long by design (a role whose next tile or stage is not ready): parked, not polled
__device__ __forceinline__ void ap_wait(uint64_t* bar, uint32_t parity)
{
#ifdef B200RMSD_AP_WAIT_SPIN  // development: the polling wait, for A/B timings
    mbar_wait(bar, parity);
#else
    mbar_wait_parked(bar, parity);
#endif
}

__device__ __forceinline__ float sel3(int c, float a0, float a1, float a2) { return c == 0 ? a0 : (c == 1 ? a1 : a2); }

}  // namespace

// ---------------------------------------------------------------------------------------------
// prepare: operands and traces.  One warp per frame.
//
// The tensor core accumulates in fp32 with truncation: measured on B200 (profiles/r01_tc_accumulation_probe.json) the
// sum of 300 products comes out ~17 ulp short of its float64 value, always towards zero.  For frames of one ensemble the
// inner products are large (~N Rg^2 / 3) while the quantity of interest, G_a + G_b - 2 lambda, is N rmsd^2: a 5e-4 bias
// on M costs 4e-5 nm on a 0.24 nm RMSD (N = 300, Rg = 1 nm), outside the 1e-5 / 1e-4 parity tolerance.  The RMSD of a
// pair is unchanged by a rigid motion of either frame, so every frame is first aligned onto the reference structure it
// is closest to (its owner c_o, allpairs_refs.cu): x' = (x - centroid) . R.  With d_j = x'_j - c_o(j),
//
//     M_ij = sum_k x'_ik x'_jk^T = sum_k x'_ik d_jk^T  +  G_i^(o(j)),     G_i^(r) = sum_k x'_ik c_rk^T   (3x3, float64 here),
//
// and the GEMM only has to accumulate X'_i D_j^T, whose entries are fluctuation-sized (~sqrt(N) Rg sigma instead of
// N Rg^2 / 3: 50 times smaller at sigma = 0.1 nm).  G_i^(r) enters the same accumulator through eight augmentation
// columns per reference, held in matrices of their own: A row (i,c) carries the tf32 pieces g1, g2 of G_i^(r)[c][0..2]
// (a third piece g3 rides in the lo matrix), B row (j,q) carries the unit vector e_q twice in the columns of ITS owner
// and zeros elsewhere.  A tile walks the augmentation K blocks of its column frames' owners after the atoms, i.e. G is
// added in the LAST K-steps: one truncation at full magnitude instead of one per accumulation step.
// Frames farther from every reference than half its radius of gyration ("far": iid test data, outliers; owner stored as
// -1-o) keep b_j = x'_j and no unit vectors, because there X'_i D_j^T and G_i would be two large numbers cancelling --
// worse than the plain product.
// Centring follows center_generic.h:3-44 (float64 mean, float32 subtraction, float64 trace of the float32 squares);
// the float32 rotation deforms a frame by ~1e-7 relative, 1e-7 nm of RMSD.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
allpairs_tc144_prepare_kernel(const float* __restrict__ xyz, int64_t n_frames, int64_t frame_stride,
                              const int* __restrict__ idx, int n_sel, int k_pad, const float* __restrict__ refs,
                              int64_t ref_stride, int n_refs, const int* __restrict__ owner,
                              const float* __restrict__ rot, const double* __restrict__ centroid,
                              float* __restrict__ a_hi, float* __restrict__ a_lo, float* __restrict__ b_hi,
                              float* __restrict__ b_lo, float* __restrict__ aug_a_hi, float* __restrict__ aug_a_lo,
                              float* __restrict__ aug_b, float* __restrict__ traces)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t)gridDim.x * 8;
    for (int64_t f = (int64_t)blockIdx.x * 8 + warp; f < n_frames; f += n_warps) {
        const float* fr = xyz + f * frame_stride;
        float R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(rot + f * 9 + i);
        const float mx = (float)__ldg(centroid + f * 3), my = (float)__ldg(centroid + f * 3 + 1),
                    mz = (float)__ldg(centroid + f * 3 + 2);
        const int64_t row = ap_tc144_row(f, 0);
        const int raw = __ldg(owner + f);
        const bool near = raw >= 0;  // warp-uniform
        const int own = near ? raw : -1 - raw;
        const float* cown = refs + own * ref_stride;
        double tr = 0;
        for (int k = lane; k < n_sel; k += 32) {
            const int a = idx ? __ldg(idx + k) : k;
            const float tx = __ldg(fr + 3 * a) - mx, ty = __ldg(fr + 3 * a + 1) - my, tz = __ldg(fr + 3 * a + 2) - mz;
#pragma unroll
            for (int c = 0; c < 3; ++c) {  // row vector x R (rotation_generic.h:40-42)
                const float v = fmaf(tz, R[6 + c], fmaf(ty, R[3 + c], tx * R[c]));
                tr += (double)(v * v);
                const float h = rna_tf32(v);
                a_hi[(row + c) * k_pad + k] = h;
                a_lo[(row + c) * k_pad + k] = rna_tf32(v - h);
                const float d = near ? v - __ldg(cown + 3 * k + c) : v;
                const float dh = rna_tf32(d);
                b_hi[(row + c) * k_pad + k] = dh;
                b_lo[(row + c) * k_pad + k] = rna_tf32(d - dh);
            }
        }
        tr = warp_sum(tr);
        if (lane == 0) traces[f] = (float)tr;
        if (near && lane < 6) {  // B row q: e_q in columns 8*own + q and 8*own + 3 + q
            const int q = lane % 3;
            aug_b[(row + q) * kApAugCols + 8 * own + lane] = 1.0f;
        }
        __syncwarp();  // the A rows written above are read back below by other lanes of this warp
        // G^(r) = A_i c_r^T with A_i = hi + lo, exactly the operand the tensor core multiplies
        for (int r = 0; r < n_refs; ++r) {
            const float* cr = refs + r * ref_stride;
            double G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int k = lane; k < n_sel; k += 32) {
                const double c0 = (double)__ldg(cr + 3 * k), c1 = (double)__ldg(cr + 3 * k + 1),
                             c2 = (double)__ldg(cr + 3 * k + 2);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double v = (double)a_hi[(row + c) * k_pad + k] + (double)a_lo[(row + c) * k_pad + k];
                    G[3 * c] += v * c0; G[3 * c + 1] += v * c1; G[3 * c + 2] += v * c2;
                }
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) G[i] = warp_sum(G[i]);
            // columns 8r + 3p + m, p = 0,1: A_hi row c <- piece p of G[c][m]; the third piece in A_lo under piece 0
            if (lane < 18) {
                const int piece = lane / 9, c = (lane % 9) / 3, m = lane % 3;
                double g = 0;
#pragma unroll
                for (int i = 0; i < 9; ++i) g = (i == 3 * c + m) ? G[i] : g;  // static indexing keeps G in registers
                // half an fp32 ulp of G towards its sign: G is the last (and by far the largest) term to enter the
                // accumulator, whose truncation then rounds the finished entry to nearest instead of towards zero --
                // unbiased, half the worst case (tests/tc_model.py ROUND_BIAS: 8e-6 -> 3e-6 nm at rmsd 0.04 nm)
                if (g != 0.0) g += copysign(scalbn(1.0, ilogb(g) - 24), g);
                const float g1 = rna_tf32((float)g);
                const float g2 = rna_tf32((float)(g - (double)g1));
                const float g3 = rna_tf32((float)(g - (double)g1 - (double)g2));
                aug_a_hi[(row + c) * kApAugCols + 8 * r + 3 * piece + m] = piece == 0 ? g1 : g2;
                if (piece == 0) aug_a_lo[(row + c) * kApAugCols + 8 * r + m] = g3;
            }
        }
    }
}

// EPI_WARPS in {8, 16}: epilogue warps (each TMEM lane quarter is served by EPI_WARPS/4 warps that split the four
// 36-column segments of the accumulator); NP in {1, 2}: independent solves interleaved per lane.
template <int EPI_WARPS, int NP, bool PAIR, bool ROT>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
allpairs_tc144_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_g_a, const __grid_constant__ CUtensorMap map_g_b,
                      const Tc144Params p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment is required by SWIZZLE_128B; dynamic smem base is not guaranteed to have it
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kRingN = PAIR ? kRingPair : kRing;
    constexpr int kStageN = PAIR ? kStagePair : kStage;
    constexpr int kBBytesN = PAIR ? kBHalfBytes : kBBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kRingN * kStageN);
    uint64_t* empty = full + kRingN;
    uint64_t* tfull = empty + kRingN;
    uint64_t* tempty = tfull + kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kAccStages);

    // the warp index through a shuffle: provably warp-uniform for the compiler, so that everything the two DMA/MMA
    // roles derive from it lives in uniform registers (see the note at the MMA issuer)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // PAIR: rank of this CTA in its 2-CTA cluster (0 = leader: issues the MMAs, owns the full[] and tempty[] barriers
    // both CTAs signal); slots are walked per cluster
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const int64_t slot0 = PAIR ? blockIdx.x >> 1 : blockIdx.x, slot_step = PAIR ? gridDim.x >> 1 : gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kRingN; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], PAIR ? 2 * EPI_WARPS : EPI_WARPS); }
        fence_mbar_init();
    }
    if (warp == EPI_WARPS) {
        if (PAIR) tmem_alloc_pair(tmem_slot, kTmem);
        else tmem_alloc(tmem_slot, kTmem);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();  // the peer's barriers must be initialised before anything arrives on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // Warp roles: warps [0, EPI_WARPS) epilogue, warp EPI_WARPS = TMA producer, warp EPI_WARPS+1 = MMA issuer.
    // The two single-thread roles get the highest warp ids: the SMSP arbiter favours higher warp ids, and a late
    // TMA or MMA issue stalls the whole pipeline while a late epilogue instruction does not.
    // Both roles run their loops with the WHOLE warp (uniform control flow, uniform values) and elect one lane only around
    // the instructions that must be issued once.  Inside `if (lane == 0)` the compiler cannot keep descriptors and
    // addresses in uniform registers, and wrapped every tcgen05.mma in a loop of ELECT + five R2UR.BROADCAST + votes:
    // ~110 cycles of dependent scalar code per 72-cycle MMA -- the issuing thread, not the tensor pipe or the operand
    // delivery, set the pace (ncu source page, round 2: 74 % of that warp's samples inside its own issue code).
    if (warp == EPI_WARPS) {
        // ===================================================== TMA producer
        {
            int stage = 0;
            uint32_t phase = 0;
            // PAIR: both CTAs load (their A rows, their half of the B rows); every load's bytes are counted on the
            // LEADER's full[stage], which the leader arms with the bytes of both CTAs.  A peer load may land before the
            // leader has armed the phase: the transaction count just goes negative for a moment (the phase cannot
            // complete before the leader's own arrival), and never earlier than that, because the peer only refills a
            // stage after the commit of the MMAs that read it, i.e. after the leader's previous phase completed.
            // Two copies per stage: the A box (hi and lo planes, four 30-row quarters each) and the B box (hi and lo planes)
            auto load_a = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int k0, int quarter) {
                if (PAIR) tma_load_4d_pair(dst, map, bar, k0, 0, quarter, 0);
                else tma_load_4d(dst, map, bar, k0, 0, quarter, 0);
            };
            auto load_b = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int k0, int row) {
                if (PAIR) tma_load_3d_pair(dst, map, bar, k0, row, 0);
                else tma_load_3d(dst, map, bar, k0, row, 0);
            };
            auto load_b1 = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int k0, int row) {  // single plane
                if (PAIR) tma_load_2d_pair(dst, map, bar, k0, row);
                else tma_load_2d(dst, map, bar, k0, row);
            };
            for (SlotWalk w(slot0, (int)slot_step, p); w.t < p.n_slots; w.next(p)) {
                int ti, tj;
                if (!w.tile(p, ti, tj)) continue;
                ti += rank;
                const int a_quarter = ti * 4, b_row = tj * kN + rank * (kN / 2);  // quarter q = rows 30q .. 30q+29
                for (int kb = 0; kb < p.nk; ++kb) {
                    ap_wait(&empty[stage], phase ^ 1u);
                    unsigned char* st = smem + stage * kStageN;
                    if (elect_one_sync()) {
                        if (rank == 0) mbar_arrive_expect_tx(&full[stage], PAIR ? 2 * kStageN : kStageN);
                        load_a(st, &map_a, &full[stage], kb * kK, a_quarter);
                        load_b(st + 2 * kABytes, &map_b, &full[stage], kb * kK, b_row);
                    }
                    if (++stage == kRingN) { stage = 0; phase ^= 1u; }
                }
                // augmentation K blocks of the references this tile's column frames are stored against (no B_lo part)
                const int2 aug = __ldg(p.tile_aug + tj);
                for (int g = aug.x; g <= aug.y; ++g) {
                    ap_wait(&empty[stage], phase ^ 1u);
                    unsigned char* st = smem + stage * kStageN;
                    if (elect_one_sync()) {
                        if (rank == 0) mbar_arrive_expect_tx(&full[stage], (PAIR ? 2 : 1) * (kStageN - kBBytesN));
                        load_a(st, &map_g_a, &full[stage], g * kK, a_quarter);
                        load_b1(st + 2 * kABytes, &map_g_b, &full[stage], g * kK, b_row);
                    }
                    if (++stage == kRingN) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ===================================================== MMA issuer
        if (rank == 0) {  // PAIR: the leader issues the M = 256 MMAs for both CTAs
            constexpr uint32_t idesc = make_tf32_idesc(PAIR ? 2 * kM : kM, kN);
            auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc_flag) {
                if (PAIR) umma_tf32_pair(d, a, b, idesc, acc_flag);
                else umma_tf32(d, a, b, idesc, acc_flag);
            };
            auto commit = [&](uint64_t* bar) {  // PAIR: arrives on the barrier at this offset in BOTH CTAs
                if (PAIR) umma_commit_pair(bar);
                else umma_commit(bar);
            };
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (SlotWalk w(slot0, (int)slot_step, p); w.t < p.n_slots; w.next(p)) {
                int ti_unused, tj;
                if (!w.tile(p, ti_unused, tj)) continue;
                ap_wait(&tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
                for (int kb = 0; kb < p.nk; ++kb) {
                    ap_wait(&full[stage], phase);
                    tc_fence_after();
                    bool continue_mma = true;
                    unsigned char* st = smem + stage * kStageN;
                    const uint64_t a_hi = make_sw128_kmajor_desc(st), a_lo = make_sw128_kmajor_desc(st + kABytes);
                    const uint64_t b_hi = make_sw128_kmajor_desc(st + 2 * kABytes),
                                   b_lo = make_sw128_kmajor_desc(st + 2 * kABytes + kBBytesN);
#ifdef B200RMSD_DEV_SWITCHES
                    if (p.flags & 0x200u) continue_mma = false;  // development: skip the MMAs (operand delivery alone)
#endif
                    if (elect_one_sync()) {
#pragma unroll
                        for (int ks = 0; ks < kK / 8; ++ks) {
                            if (!continue_mma) break;
                            const uint64_t off = (uint64_t)(ks * 2);  // 8 floats = 32 bytes = 2 x 16-byte units
                            mma(d_tmem, a_lo + off, b_hi + off, (kb | ks) != 0 ? 1u : 0u);
                            mma(d_tmem, a_hi + off, b_lo + off, 1u);
                            mma(d_tmem, a_hi + off, b_hi + off, 1u);
                        }
                        commit(&empty[stage]);  // frees this smem stage when the MMAs above have read it
                    }
                    if (++stage == kRingN) { stage = 0; phase ^= 1u; }
                }
                const int2 aug = __ldg(p.tile_aug + tj);
                for (int g = aug.x; g <= aug.y; ++g) {  // + X'_i c_r^T for the four references of block g: (g1 + g2 + g3) . e
                    ap_wait(&full[stage], phase);
                    tc_fence_after();
                    unsigned char* st = smem + stage * kStageN;
                    const uint64_t a_hi = make_sw128_kmajor_desc(st), a_lo = make_sw128_kmajor_desc(st + kABytes);
                    const uint64_t b_hi = make_sw128_kmajor_desc(st + 2 * kABytes);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int ks = 0; ks < kK / 8; ++ks) {
                            const uint64_t off = (uint64_t)(ks * 2);
                            mma(d_tmem, a_lo + off, b_hi + off, 1u);
                            mma(d_tmem, a_hi + off, b_hi + off, 1u);
                        }
                        commit(&empty[stage]);
                    }
                    if (++stage == kRingN) { stage = 0; phase ^= 1u; }
                }
                if (elect_one_sync()) commit(&tfull[acc]);  // accumulator complete (both CTAs' epilogues)
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================================================== epilogue (TMEM -> QCP -> HBM)
        const int ew = warp & 3;                 // TMEM lane quarter this warp may read (warp_id % 4)
        constexpr int kSegsPerWarp = 16 / EPI_WARPS;  // 2 or 1 of the four 36-column segments
        const int part = warp >> 2;              // which share of the segments
        const int c = lane % 3, tq = lane / 3;   // component row and frame slot of this lane
        const bool row_valid = lane < 30;        // lanes 30, 31: first rows of the next quarter's frame, ignored
        const int src1 = lane - c + (c + 1) % 3, src2 = lane - c + (c + 2) % 3;
        const float inv_n = 1.0f / (float)p.n_sel;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (SlotWalk w(slot0, (int)slot_step, p); w.t < p.n_slots; w.next(p)) {
            int ti, tj;
            if (!w.tile(p, ti, tj)) continue;
            ti += rank;
            const int64_t fi = (int64_t)ti * kIFrames + ew * 10 + tq;  // row frame of this lane
            const bool i_ok = row_valid && fi >= p.row0 && fi < p.row1;
            const float Gi = i_ok ? __ldg(p.traces + fi) : 1.0f;
            float Ri[ROT ? 9 : 1];
            if constexpr (ROT) {
#pragma unroll
                for (int k = 0; k < 9; ++k) Ri[k] = i_ok ? __ldg(p.frame_rot + fi * 9 + k) : 0.f;
            }
            ap_wait(&tfull[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < kSegsPerWarp; ++cc) {
                const int seg = part * kSegsPerWarp + cc;
                uint32_t r[kSegCols];
                {
                    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)acc * kAccStride +
                                           (uint32_t)(seg * kSegCols);
#pragma unroll
                    for (int q4 = 0; q4 < kSegCols / 4; ++q4)
                        tmem_ld_32x32b_x4(taddr + 4 * q4, r[4 * q4], r[4 * q4 + 1], r[4 * q4 + 2], r[4 * q4 + 3]);
                    tmem_ld_wait();
                }
#pragma unroll
                for (int jp = 0; jp < 4 / NP; ++jp) {
                    // NP groups of three j-frames per pass so that NP independent solves interleave
                    float M[NP][9], Ga[NP], Gb[NP], res[NP];
                    int64_t fj[NP];
                    bool ok[NP], trusted[NP];
                    // column 9*jg + 3*m + q of the segment = (this lane's row) . (component q of j-frame 3*jg + m).
                    // Lane (tq, c) ends up with the block of pair (i_tq, j-frame 3*jg + c) as m[3r + q] =
                    // M[(c+r)%3][(c+q)%3]: rows AND columns in the cyclic order starting at the lane's own component.
                    // That relabels the axes of both frames the same way (x,y,z -> y,z,x is a proper rotation applied to
                    // both), so the key matrix keeps its eigenvalues, tr M and the common orientation the shifted solver
                    // relies on -- and every entry costs ONE lane-dependent select: in round r the lane that holds row
                    // (c+r)%3 publishes the entry of the j-frame its reader works on, already at its permuted column.
                    // Warp-collective (shuffles).
                    auto gather = [&](int jg, float (&m)[9]) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const float e0 = __uint_as_float(r[9 * jg + q]);                    // j-frame 0, column q
                            const float e1 = __uint_as_float(r[9 * jg + 3 + (q + 1) % 3]);      // j-frame 1, column q+1
                            const float e2 = __uint_as_float(r[9 * jg + 6 + (q + 2) % 3]);      // j-frame 2, column q+2
                            m[q] = sel3(c, e0, e1, e2);
                            const float send1 = sel3(c, e2, e0, e1);  // for the reader whose component is (c+2)%3
                            const float send2 = sel3(c, e1, e2, e0);  // for the reader whose component is (c+1)%3
                            m[3 + q] = __shfl_sync(0xffffffffu, send1, src1);
                            m[6 + q] = __shfl_sync(0xffffffffu, send2, src2);
                        }
                    };
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        const int jg = NP * jp + u;
                        gather(jg, M[u]);
                        fj[u] = (int64_t)tj * kJFrames + seg * 12 + 3 * jg + c;
                        ok[u] = i_ok && fj[u] >= p.col0 && fj[u] < p.col1;
                        Ga[u] = ok[u] ? __ldg(p.traces + fj[u]) : 1.0f;
                        Gb[u] = Gi;
                        trusted[u] = true;
                    }
#ifdef B200RMSD_DEV_SWITCHES
                    if (p.flags & 0x100u) {  // development: skip the solve to time the GEMM main loop alone
#pragma unroll
                        for (int u = 0; u < NP; ++u) res[u] = M[u][0];
                    } else
#endif
                    if (p.flags & B200RMSD_FAST_SOLVE) {  // all-float32 solve on lambda (reference-class precision)
                        qcp_msd_f32<NP>(M, Ga, Gb, ok, inv_n, res, trusted);
                    } else {                              // float32 solve on lambda - tr M: 1e-7 nm class on aligned pairs
                        qcp_msd_shift<NP>(M, Ga, Gb, ok, (float)p.n_sel, res, trusted);
                    }
                    bool all_trusted = true;
#pragma unroll
                    for (int u = 0; u < NP; ++u) all_trusted = all_trusted && trusted[u];
                    if (!__all_sync(0xffffffffu, all_trusted)) {
                        // Uncommon.  (1) similar frames that do not share an orientation: the float64-polished solve on
                        // lambda; (2) collinear atoms, two-atom selections (a double largest root): the closed form.
                        // The blocks are fetched again -- M is dead by now, which keeps it out of the fast solver's
                        // register budget (unrolled: the accumulator registers must stay statically indexed).
                        float M2[NP][9], res2[NP];
                        bool trusted2[NP];
#pragma unroll
                        for (int u = 0; u < NP; ++u) gather(NP * jp + u, M2[u]);
                        qcp_msd_fast<NP>(M2, Ga, Gb, ok, inv_n, res2, trusted2);
#pragma unroll
                        for (int u = 0; u < NP; ++u) {
                            if (!trusted[u]) res[u] = trusted2[u] ? res2[u] : qcp_rmsd_closed(M2[u], Ga[u], Gb[u], inv_n);
                        }
                    }
                    if constexpr (ROT) {
                        // The rotation that superposes frame j onto frame i (north_star: "plus a rotation matrix for
                        // superpose"; the convention of b200rmsd_rmsd_dev's out_rot and rotation_generic.h:40-42:
                        // (x_j - centroid_j) U ~ x_i - centroid_i).  The blocks are fetched once more (M is dead after the
                        // solve, as on the fall-back route) and go through the float64 solver with the quaternion
                        // (qcp_solve: cofactors of row 0 of K - lambda I, theobald_rmsd.cpp:280-334), with rows = the frame
                        // that is rotated (j).  The block belongs to the frames as the prepare step aligned them
                        // (x' = (x - centroid) R_f) and its axes are in this lane's cyclic order, so the result W' is
                        // first put back into x,y,z order (W[(c+r)%3][(c+q)%3] = W'[r][q]) and then carried to the frames'
                        // own orientations: X'_j W ~ X'_i  =>  U = R_j W R_i^T.
                        float M3[NP][9];
#pragma unroll
                        for (int u = 0; u < NP; ++u) gather(NP * jp + u, M3[u]);
#pragma unroll
                        for (int u = 0; u < NP; ++u) {
                            if (!ok[u]) continue;
                            float* dst = p.out_rot + ((size_t)(fi - p.row0) * (size_t)(p.col1 - p.col0) + (size_t)(fj[u] - p.col0)) * 9;
                            if (fi == fj[u] && (p.flags & B200RMSD_DIAG_ZERO)) {
#pragma unroll
                                for (int k = 0; k < 9; ++k) dst[k] = (k % 4 == 0) ? 1.0f : 0.0f;
                                continue;
                            }
                            QcpInput qi;
                            qi.inv_n = (double)inv_n;
                            qi.Ga = (double)Ga[u];
                            qi.Gb = (double)Gb[u];
#pragma unroll
                            for (int r = 0; r < 3; ++r)
#pragma unroll
                                for (int q = 0; q < 3; ++q) qi.M[3 * r + q] = (double)M3[u][3 * q + r];
                            float Wp[9], W[9], Rj[9], T[9];
                            qcp_solve(qi, Wp, nullptr);
#pragma unroll
                            for (int a = 0; a < 3; ++a)
#pragma unroll
                                for (int b = 0; b < 3; ++b)
                                    W[3 * a + b] = sel3(c, Wp[3 * a + b], Wp[3 * ((a + 2) % 3) + (b + 2) % 3],
                                                        Wp[3 * ((a + 1) % 3) + (b + 1) % 3]);
#pragma unroll
                            for (int k = 0; k < 9; ++k) Rj[k] = __ldg(p.frame_rot + fj[u] * 9 + k);
#pragma unroll
                            for (int a = 0; a < 3; ++a)
#pragma unroll
                                for (int b = 0; b < 3; ++b)
                                    T[3 * a + b] = fmaf(Rj[3 * a + 2], W[6 + b], fmaf(Rj[3 * a + 1], W[3 + b], Rj[3 * a] * W[b]));
#pragma unroll
                            for (int a = 0; a < 3; ++a)
#pragma unroll
                                for (int b = 0; b < 3; ++b)
                                    dst[3 * a + b] = fmaf(T[3 * a + 2], Ri[3 * b + 2], fmaf(T[3 * a + 1], Ri[3 * b + 1], T[3 * a] * Ri[3 * b]));
                        }
                    }
#ifdef B200RMSD_DEV_SWITCHES
                    if (p.flags & 0x400u) {  // development: no global stores (one per warp so the work is not dead code)
                        float acc_v = 0.f;
#pragma unroll
                        for (int u = 0; u < NP; ++u) acc_v += res[u];
                        if (acc_v == 123.456f) p.out[0] = acc_v;
                    } else
#endif
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
            if (lane == 0) {
                if (PAIR) mbar_arrive_leader(&tempty[acc]);  // the leader's MMA thread waits for both CTAs' epilogues
                else mbar_arrive(&tempty[acc]);
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all();  // neither CTA may exit (or free its TMEM) while the other can still reach into it
    else __syncthreads();
    if (warp == EPI_WARPS) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, kTmem);
        else tmem_dealloc(tmem_base, kTmem);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// tile grid and super-block walk of the block [row0,row1) x [col0,col1)
static Tc144Params tc144_tiling(int64_t row0, int64_t row1, int64_t col0, int64_t col1, bool has_out_t, bool pair = false)
{
    Tc144Params p{};
    p.pair = pair ? 1 : 0;
    p.row0 = row0;
    p.row1 = row1;
    p.col0 = col0;
    p.col1 = col1;
    p.tiles_i0 = (int)(row0 / kIFrames);
    p.tiles_i = (int)((row1 + kIFrames - 1) / kIFrames) - p.tiles_i0;
    p.tiles_j0 = (int)(col0 / kJFrames);
    p.tiles_j = (int)((col1 + kJFrames - 1) / kJFrames) - p.tiles_j0;
    // a square block on the diagonal: compute each unordered pair once and mirror it (exactly symmetric, half the flops)
    p.symmetric = (row0 == col0 && row1 == col1 && !has_out_t) ? 1 : 0;
#ifdef B200RMSD_DEV_SWITCHES
    if (getenv("B200RMSD_NO_SYMMETRIC")) p.symmetric = 0;
#endif
    p.n_bj = (p.tiles_j + kSupJ - 1) / kSupJ;
    // a super-block is kSupI i-tiles high either way: kSupI slots, or kSupI / 2 pair slots
    p.n_slots = (int64_t)((p.tiles_i + kSupI - 1) / kSupI) * p.n_bj * ((pair ? kSupI / 2 : kSupI) * kSupJ);
    return p;
}

// development / test hook (no device needed): the tiles the kernel visits for a block, in slot order, as (ti, tj) pairs
// in tiles[0 .. 2*cap); returns how many there are, and whether the block runs in symmetric mode in *symmetric.
extern "C" long long b200rmsd_debug_allpairs_tiles(long long row0, long long row1, long long col0, long long col1,
                                                   int has_out_t, int* tiles, long long cap, int* symmetric)
{
    const Tc144Params p = tc144_tiling(row0, row1, col0, col1, has_out_t != 0);
    if (symmetric) *symmetric = p.symmetric;
    long long n = 0;
    for (int64_t t = 0; t < p.n_slots; ++t) {
        int ti, tj;
        if (!tile_of_slot(t, p, ti, tj)) continue;
        if (n < cap) { tiles[2 * n] = ti; tiles[2 * n + 1] = tj; }
        ++n;
    }
    return n;
}

cudaError_t launch_allpairs_tc144_prepare(const float* xyz, int64_t n_frames, int64_t frame_stride, const int* idx,
                                          int n_sel, const ApGeometry& g, char* base, int n_refs, int sm_count,
                                          cudaStream_t st)
{
    // K padding, unused augmentation columns and the rows past 3F (read by the last tiles' boxes) are zero
    const size_t atom_bytes = (size_t)g.rows_pad * g.k_pad * 4, aug_bytes = (size_t)g.rows_pad * kApAugCols * 4;
    const size_t offs[7] = {g.a_hi_off, g.a_lo_off, g.b_hi_off, g.b_lo_off, g.aug_a_hi_off, g.aug_a_lo_off, g.aug_b_off};
    for (int i = 0; i < 7; ++i) {
        cudaError_t e = cudaMemsetAsync(base + offs[i], 0, i < 4 ? atom_bytes : aug_bytes, st);
        if (e != cudaSuccess) return e;
    }
    int64_t ctas = (int64_t)sm_count * 8;
    const int64_t need = (n_frames + 7) / 8;
    if (ctas > need) ctas = need;
    allpairs_tc144_prepare_kernel<<<(unsigned)ctas, 256, 0, st>>>(
        xyz, n_frames, frame_stride, idx, n_sel, g.k_pad, (const float*)(base + g.ref_off), (int64_t)(g.ref_stride / 4),
        n_refs, (const int*)(base + g.owner_off), (const float*)(base + g.rot_off), (const double*)(base + g.cen_off),
        (float*)(base + g.a_hi_off), (float*)(base + g.a_lo_off), (float*)(base + g.b_hi_off), (float*)(base + g.b_lo_off),
        (float*)(base + g.aug_a_hi_off), (float*)(base + g.aug_a_lo_off), (float*)(base + g.aug_b_off),
        (float*)(base + g.traces_off));
    return cudaGetLastError();
}

int launch_allpairs_tc144_block(const ApGeometry& g, const char* base, int n_sel, int64_t n_frames, int64_t row0,
                                int64_t row1, int64_t col0, int64_t col1, float* out, int64_t ld, float* out_t,
                                int64_t ld_t, float* out_rot, unsigned flags, int sm_count, cudaStream_t st)
{
    (void)n_frames;
    // rotations: single-CTA geometry, every pair computed (the mirrored entry would need the transposed rotation)
    bool pair = g_ap_cta_pair != 0 && sm_count >= 2 && !out_rot;
#ifdef B200RMSD_DEV_SWITCHES
    if (const char* v = getenv("B200RMSD_TC_PAIR")) pair = atoi(v) != 0 && sm_count >= 2;
#endif
    const int b_box = pair ? kN / 2 : kN;  // CTA-pair mode: each CTA loads half of the B rows of a j-tile
    CUtensorMap map_a, map_b, map_g_a, map_g_b;
    auto op = [&](size_t off) { return (const float*)(base + off); };
    if (!make_a_operand_map(&map_a, op(g.a_hi_off), g.rows_pad, g.k_pad, g.a_lo_off - g.a_hi_off) ||
        !make_b_operand_map(&map_b, op(g.b_hi_off), g.rows_pad, g.k_pad, b_box, g.b_lo_off - g.b_hi_off) ||
        !make_a_operand_map(&map_g_a, op(g.aug_a_hi_off), g.rows_pad, kApAugCols, g.aug_a_lo_off - g.aug_a_hi_off) ||
        !make_operand_map(&map_g_b, op(g.aug_b_off), g.rows_pad, kApAugCols, b_box))
        return set_error(B200RMSD_ECUDA, "allpairs: cuTensorMapEncodeTiled failed");
    Tc144Params p = tc144_tiling(row0, row1, col0, col1, out_t != nullptr || out_rot != nullptr, pair);
    p.out_rot = out_rot;
    p.frame_rot = op(g.rot_off);
    p.traces = op(g.traces_off);
    p.out = out;
    p.ld = ld;
    p.out_t = out_t;
    p.ld_t = ld_t;
    p.n_sel = n_sel;
    p.nk = g.k_pad / kK;
    p.tile_aug = (const int2*)(base + g.tile_aug_off);
    p.flags = flags & (B200RMSD_DIAG_ZERO | B200RMSD_FAST_SOLVE);
    int ew = 16, np = 2;
#ifdef B200RMSD_DEV_SWITCHES
    if (const char* dbg = getenv("B200RMSD_TC_DEBUG")) p.flags |= (unsigned)strtoul(dbg, nullptr, 0) & 0xff00u;
    if (const char* cfg = getenv("B200RMSD_TC_EPILOGUE")) sscanf(cfg, "%dx%d", &ew, &np);  // "<warps>x<np>", e.g. 16x1
#endif
    const size_t smem = (size_t)(pair ? kRingPair * kStagePair : kRing * kStage) + 1024 + 256;
    int64_t ctas = sm_count;
    const int64_t n_tiles = (int64_t)p.tiles_i * p.tiles_j;
    if (ctas > n_tiles) ctas = n_tiles;
    if (ctas < 1) ctas = 1;
    if (pair) ctas = std::max<int64_t>(2, ctas & ~(int64_t)1);  // whole clusters
    cudaError_t e = cudaSuccess;
    // one launch path for both modes: a 2-CTA cluster (the two SMs of a TPC) in pair mode
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define B200_LAUNCH_TC144(EW, NP, PAIR, ROT)                                                                               \
    do {                                                                                                                   \
        auto kern = allpairs_tc144_kernel<EW, NP, PAIR, ROT>;                                                                \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                            \
        cfg.blockDim = dim3(64 + 32 * EW);                                                                                 \
        if (e == cudaSuccess)                                                                                              \
            e = cudaLaunchKernelEx(&cfg, kern, map_a, map_b, map_g_a, map_g_b, p);                                        \
    } while (0)
#ifdef B200RMSD_DEV_SWITCHES
    if (ew == 8 && np == 2 && !out_rot) B200_LAUNCH_TC144(8, 2, false, false);
    else if (ew == 16 && np == 1 && !out_rot) B200_LAUNCH_TC144(16, 1, false, false);
    else
#endif
    if (out_rot) B200_LAUNCH_TC144(16, 2, false, true);
    else if (pair) B200_LAUNCH_TC144(16, 2, true, false);
    else B200_LAUNCH_TC144(16, 2, false, false);
#undef B200_LAUNCH_TC144
    if (e != cudaSuccess) return set_error(B200RMSD_ECUDA, "allpairs: %s", cudaGetErrorString(e));
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error(B200RMSD_ECUDA, "allpairs: %s", cudaGetErrorString(e));
}

}  // namespace b200
