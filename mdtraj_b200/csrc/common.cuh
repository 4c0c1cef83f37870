// common.cuh -- sm_100a PTX wrappers shared by the RMSD kernels.
//
// mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP / SYNCS), cache-hinted
// vector loads/stores, and the 16-value warp reduce-scatter used by every
// streaming kernel.  No library dependencies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// one lane of the (converged) warp: the idiom that lets a warp-uniform loop issue single-thread instructions
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    // make the inits visible to the async proxy before any bulk copy targets them
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// The same wait with a suspend-time hint: the thread is parked until the phase completes (or the hint, in nanoseconds,
// runs out) instead of coming back every ~100 cycles.  For waits that are long by design -- a warp whose next tile is not
// ready: in the all-pairs kernel the polling loops were 16 % of all issued instructions, on the schedulers the
// epilogue warps need.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
    } while (!ok);
}

// generic-proxy accesses before this point are ordered before later async-proxy
// (bulk copy) accesses to the same shared memory
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- 1-D bulk copies (TMA engine, no tensor map needed) ---------------------
// global -> shared, completion counted in bytes on an mbarrier.  src/dst 16-byte
// aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// same, with an L2 eviction-priority policy (streamed-once data: evict_first)
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// shared -> global
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- streaming vector global access -----------------------------------------
__device__ __forceinline__ float4 ldg_stream(const float4* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// ---- 16-value warp reduce-scatter -------------------------------------------
// ---- the 16 running sums of one frame against the reference (see finish_frame in one_vs_many.cu for the record layout)
// accumulate one atom
template <bool PRE>
__device__ __forceinline__ void acc_atom(float (&v)[16], float ax, float ay, float az, float bx, float by, float bz,
                                         float px, float py, float pz)
{
    if (!PRE) {
        ax -= px; ay -= py; az -= pz;
        v[0] += ax; v[1] += ay; v[2] += az;
        v[3] = fmaf(ax, ax, v[3]); v[3] = fmaf(ay, ay, v[3]); v[3] = fmaf(az, az, v[3]);
    }
    v[4] = fmaf(ax, bx, v[4]);   v[5] = fmaf(ax, by, v[5]);   v[6] = fmaf(ax, bz, v[6]);
    v[7] = fmaf(ay, bx, v[7]);   v[8] = fmaf(ay, by, v[8]);   v[9] = fmaf(ay, bz, v[9]);
    v[10] = fmaf(az, bx, v[10]); v[11] = fmaf(az, by, v[11]); v[12] = fmaf(az, bz, v[12]);
}

// one unit = 4 atoms held in 3 float4 (x0 y0 z0 x1 | y1 z1 x2 y2 | z2 x3 y3 z3)
template <bool PRE>
__device__ __forceinline__ void acc_unit(float (&v)[16], const float4& a0, const float4& a1, const float4& a2,
                                         const float4& b0, const float4& b1, const float4& b2, float px, float py,
                                         float pz, int nvalid)
{
    acc_atom<PRE>(v, a0.x, a0.y, a0.z, b0.x, b0.y, b0.z, px, py, pz);
    if (nvalid >= 4) {
        acc_atom<PRE>(v, a0.w, a1.x, a1.y, b0.w, b1.x, b1.y, px, py, pz);
        acc_atom<PRE>(v, a1.z, a1.w, a2.x, b1.z, b1.w, b2.x, px, py, pz);
        acc_atom<PRE>(v, a2.y, a2.z, a2.w, b2.y, b2.z, b2.w, px, py, pz);
    } else {  // ragged last unit of the frame: padding atoms must not see the pivot
        if (nvalid > 1) acc_atom<PRE>(v, a0.w, a1.x, a1.y, b0.w, b1.x, b1.y, px, py, pz);
        if (nvalid > 2) acc_atom<PRE>(v, a1.z, a1.w, a2.x, b1.z, b1.w, b2.x, px, py, pz);
    }
}

// Every lane holds v[0..15]; on return v[0] of lane L is sum over all 32 lanes of
// value index (L >> 1).  16 shuffles instead of 80 for 16 separate all-reduces.
__device__ __forceinline__ void warp_reduce_scatter16(float (&v)[16], int lane)
{
    const unsigned full = 0xffffffffu;
    {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float send = up ? v[j] : v[j + 8];
            const float keep = up ? v[j + 8] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float send = up ? v[j] : v[j + 4];
            const float keep = up ? v[j + 4] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float send = up ? v[j] : v[j + 2];
            const float keep = up ? v[j + 2] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, 4);
        }
    }
    {
        const bool up = lane & 2;
        const float send = up ? v[0] : v[1];
        const float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(full, send, 2);
    }
    v[0] += __shfl_xor_sync(full, v[0], 1);
}

// Same butterfly inside aligned groups of L lanes (L = 2, 4, 8, 16): L lanes share one frame.
template <int L>
__device__ __forceinline__ int group_reduce_scatter16(float (&v)[16], int lane)
{
    // after the call v[0 .. 16/L) of this lane hold the group sums of value indices base .. base + 16/L - 1
    int base = 0;
    int cnt = 8;
#pragma unroll
    for (int h = L / 2; h >= 1; h >>= 1) {
        const bool up = lane & h;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < cnt) {
                const float send = up ? v[j] : v[j + cnt];
                const float keep = up ? v[j + cnt] : v[j];
                v[j] = keep + __shfl_xor_sync(0xffffffffu, send, h);
            }
        }
        base += up ? cnt : 0;
        cnt >>= 1;
    }
    return base;
}

// ---- the same reduction with the LAST stages in float64 -------------------------------------------------------------
// What the float32 butterfly above loses is the rounding of sums whose magnitude is the whole frame's (ulp(1e4) = 1e-3
// against N msd ~ 30): the first two stages add 2 and 4 lane partials (small, float32 is enough, and they carry 12 of
// the 16 exchanges), the last stages add the big numbers and run in float64: 20 SHFL + 12 FADD + 4 DADD instead of
// 16 SHFL + 16 FADD.  The streaming kernels call it once per 1024 atoms (8 units per lane) and add the results in
// float64, so no float32 number ever holds more than 32 atoms' worth of a sum (numpy emulation of the variants:
// DESIGN.md section 4 -- all-float64 stages gain nothing more).
// On return the value of index (lane >> 1), summed over all 32 lanes.
__device__ __forceinline__ double warp_reduce_scatter16_mixed(float (&v)[16], int lane)
{
    const unsigned full = 0xffffffffu;
    {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float send = up ? v[j] : v[j + 8];
            const float keep = up ? v[j + 8] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float send = up ? v[j] : v[j + 4];
            const float keep = up ? v[j + 4] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, 8);
        }
    }
    double d0, d1;
    {
        const bool up = lane & 4;
        const float s0 = up ? v[0] : v[2], k0 = up ? v[2] : v[0];
        const float s1 = up ? v[1] : v[3], k1 = up ? v[3] : v[1];
        d0 = (double)k0 + (double)__shfl_xor_sync(full, s0, 4);   // the exchange itself still moves floats
        d1 = (double)k1 + (double)__shfl_xor_sync(full, s1, 4);
    }
    {
        const bool up = lane & 2;
        const double send = up ? d0 : d1;
        const double keep = up ? d1 : d0;
        d0 = keep + __shfl_xor_sync(full, send, 2);
    }
    return d0 + __shfl_xor_sync(full, d0, 1);
}

// acc_unit with the mean-square shift: every atom's |x - p|^2 enters the running sum minus a constant c (the same for
// all frames: the reference's mean-square radius), so that the float32 lane partial of sum |x - p|^2 -- all terms
// positive, by far the largest of the 13 sums -- stays fluctuation-sized instead of growing linearly; the epilogue adds
// c * n_atoms back in float64.  One FADD per 4 atoms.
template <bool PRE>
__device__ __forceinline__ void acc_unit_shift(float (&v)[16], const float4& a0, const float4& a1, const float4& a2,
                                               const float4& b0, const float4& b1, const float4& b2, float px, float py,
                                               float pz, int nvalid, float c)
{
    acc_unit<PRE>(v, a0, a1, a2, b0, b1, b2, px, py, pz, nvalid);
    if (!PRE) v[3] = fmaf(-c, nvalid >= 4 ? 4.0f : (float)nvalid, v[3]);
}

// Mean of one value per lane over the lanes of `mask` (n of them), bit-identical on all of them, in 5 instructions
// instead of a log2(n)-stage shuffle butterfly: fixed point (2^-7 nm) through the integer REDUX unit.  Only used for the
// accumulation pivot, which may be ANY finite number near the centroid as long as every lane uses the same one (it is
// carried along exactly and added back in float64); out-of-range coordinates (> 5e5 nm) merely give a poor pivot.
__device__ __forceinline__ float lanes_mean_fixed(float x, unsigned mask, float inv_n)
{
    const int q = __float2int_rn(x * 128.0f);  // saturating
    return (float)__reduce_add_sync(mask, q) * (inv_n * 0.0078125f);
}

__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ float warp_sum(float x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

}  // namespace b200
