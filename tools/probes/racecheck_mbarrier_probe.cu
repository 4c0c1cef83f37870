// Probe for compute-sanitizer racecheck: does it order generic shared-memory accesses of different warps that are
// synchronised ONLY by an mbarrier with several arrivals (the hand-over superpose_pipe_kernel uses between its streaming
// warps and its solver warps)?  Kernel A: two producer warps write, lane 0 of each arrives (count 2), a consumer warp
// waits and reads.  Kernel B: one producer warp, count 1.  Both are race-free by the PTX memory model
// (mbarrier.arrive has release, try_wait acquire semantics at CTA scope).
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/probe tools/probes/racecheck_mbarrier_probe.cu
//   compute-sanitizer --tool racecheck /tmp/probe
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
}
template <int PRODUCERS>
__global__ void probe(float* out, int rounds)
{
    __shared__ float data[PRODUCERS * 32];
    __shared__ uint64_t filled, emptied;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&filled, PRODUCERS); mbar_init(&emptied, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    float acc = 0.f;
    for (int r = 0; r < rounds; ++r) {
        if (warp < PRODUCERS) {
            if (r > 0) mbar_wait(&emptied, (r - 1) & 1);
            data[warp * 32 + lane] = (float)(r + lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(&filled);
        } else {
            mbar_wait(&filled, r & 1);
            for (int w = 0; w < PRODUCERS; ++w) acc += data[w * 32 + lane];
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptied);
        }
    }
    if (warp == PRODUCERS) out[lane] = acc;
}
int main()
{
    float* out;
    cudaMalloc(&out, 128);
    probe<2><<<1, 96>>>(out, 50);
    printf("kernel A (2 arrivals): %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    probe<1><<<1, 64>>>(out, 50);
    printf("kernel B (1 arrival):  %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
