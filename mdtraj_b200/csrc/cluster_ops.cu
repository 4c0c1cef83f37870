// cluster_ops.cu -- consumers of the all-pairs matrix that the reference's example notebooks run on the host with numpy /
// scipy (SURVEY.md section 8(f) "next" #3), as HBM-bound reductions over the device-resident matrix:
//
//   matrix_moments   sum d, sum d^2 over all entries            -> distances.std()            (centroids.ipynb:117)
//   exp_rowsum       s_i = sum_j exp(scale * d_ij)              -> np.exp(-beta*d/std).sum(1)  (centroids.ipynb:117)
//   row_argmin       argmin_j d_ij, min_j d_ij (first minimum)   -> np.argmin(md.rmsd(leaders, frame, 0))
//                                                                   (two-pass-clustering.ipynb cell 13)
//   condense         upper triangle, row-major                   -> scipy squareform(d, checks=False) (clustering.ipynb:101)
//
// Each reads the matrix once (4 bytes per entry); float32 partial sums over at most 16 entries, float64 above that.
#include <cstdint>

#include "../../include/b200rmsd.h"
#include "kernels.cuh"

namespace b200 {

__device__ __forceinline__ double warp_sum_f64(double x)
{
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) x += __shfl_xor_sync(0xffffffffu, x, h);
    return x;
}

// grid-stride over rows; a warp walks one row in 128-element steps (float4 per lane when the row is 16-byte aligned)
__global__ void __launch_bounds__(256) matrix_moments_kernel(const float* __restrict__ D, int64_t rows, int64_t cols, int64_t ld,
                                                             double* __restrict__ out2)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = warp; i < rows; i += n_warps) {
        const float* row = D + i * ld;
        const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15u) == 0);
        int64_t j = 0;
        if (vec) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            const int64_t n4 = cols >> 2;
            for (int64_t q = lane; q < n4; q += 128) {  // up to 4 float4 = 16 entries per lane in float32
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int64_t qq = q + 32 * k;
                    if (qq < n4) {
                        const float4 v = __ldcs(row4 + qq);
                        a += (v.x + v.y) + (v.z + v.w);
                        b = fmaf(v.x, v.x, b); b = fmaf(v.y, v.y, b); b = fmaf(v.z, v.z, b); b = fmaf(v.w, v.w, b);
                    }
                }
                s1 += (double)a; s2 += (double)b;
            }
            j = n4 << 2;
        }
        for (int64_t k = j + lane; k < cols; k += 32) {
            const float v = row[k];
            s1 += (double)v; s2 += (double)v * (double)v;
        }
    }
    s1 = warp_sum_f64(s1); s2 = warp_sum_f64(s2);
    __shared__ double part[2][8];
    if (lane == 0) { part[0][threadIdx.x >> 5] = s1; part[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += part[0][w]; b += part[1][w]; }
        atomicAdd(out2, a);
        atomicAdd(out2 + 1, b);
    }
}

// one warp per row: rowsum[i] (+)= sum_j exp(scale * d_ij)
__global__ void __launch_bounds__(256) exp_rowsum_kernel(const float* __restrict__ D, int64_t rows, int64_t cols, int64_t ld,
                                                         float scale, int accumulate, double* __restrict__ rowsum)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < rows; i += n_warps) {
        const float* row = D + i * ld;
        double s = 0.0;
        const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15u) == 0);
        int64_t j = 0;
        if (vec) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            const int64_t n4 = cols >> 2;
            for (int64_t q = lane; q < n4; q += 128) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int64_t qq = q + 32 * k;
                    if (qq < n4) {
                        const float4 v = __ldcs(row4 + qq);
                        a += (expf(scale * v.x) + expf(scale * v.y)) + (expf(scale * v.z) + expf(scale * v.w));
                    }
                }
                s += (double)a;
            }
            j = n4 << 2;
        }
        for (int64_t k = j + lane; k < cols; k += 32) s += (double)expf(scale * row[k]);
        s = warp_sum_f64(s);
        if (lane == 0) rowsum[i] = accumulate ? rowsum[i] + s : s;
    }
}

// one warp per row: first minimum (numpy argmin semantics; NaN never wins)
__global__ void __launch_bounds__(256) row_argmin_kernel(const float* __restrict__ D, int64_t rows, int64_t cols, int64_t ld,
                                                         int32_t* __restrict__ arg, float* __restrict__ val)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < rows; i += n_warps) {
        const float* row = D + i * ld;
        float best = __int_as_float(0x7f800000);  // +inf
        int64_t bj = cols;
        for (int64_t k = lane; k < cols; k += 32) {
            const float v = __ldcs(row + k);
            if (v < best) { best = v; bj = k; }  // strict: keeps the first occurrence inside a lane
        }
#pragma unroll
        for (int h = 16; h >= 1; h >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, h);
            const int64_t oj = __shfl_xor_sync(0xffffffffu, bj, h);
            if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
        if (lane == 0) {
            arg[i] = bj < cols ? (int32_t)bj : 0;
            if (val) val[i] = best;
        }
    }
}

// one CTA per row i: out[i*n - i*(i+1)/2 + (j-i-1)] = D[i*ld + j], j > i
__global__ void __launch_bounds__(256) condense_kernel(const float* __restrict__ D, int64_t n, int64_t ld, float* __restrict__ out)
{
    for (int64_t i = blockIdx.x; i + 1 < n; i += gridDim.x) {
        const float* row = D + i * ld;
        float* dst = out + (i * n - (i * (i + 1)) / 2) - (i + 1);
        for (int64_t j = i + 1 + threadIdx.x; j < n; j += blockDim.x) dst[j] = __ldcs(row + j);
    }
}

}  // namespace b200

#define fail b200::set_error

static int launch_grid(int64_t rows_as_warps, int* sm_out)
{
    int dev = 0, sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm <= 0)
        return -1;
    *sm_out = sm;
    int64_t ctas = (rows_as_warps + 7) / 8;
    const int64_t cap = (int64_t)sm * 8;  // 8 x 256 threads per SM
    return (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
}

extern "C" {

int b200rmsd_matrix_moments_dev(const float* D, int64_t rows, int64_t cols, int64_t ld, double* out2, void* stream)
{
    if (!D || !out2 || rows < 0 || cols < 0 || ld < cols) return fail(B200RMSD_EINVAL, "matrix_moments: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) return fail(B200RMSD_ECUDA, "matrix_moments: %s", cudaGetErrorString(e));
    if (rows == 0 || cols == 0) return 0;
    int sm = 0;
    const int grid = launch_grid(rows, &sm);
    if (grid < 0) return fail(B200RMSD_ENODEVICE, "matrix_moments: no CUDA device");
    b200::matrix_moments_kernel<<<grid, 256, 0, st>>>(D, rows, cols, ld, out2);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail(B200RMSD_ECUDA, "matrix_moments: %s", cudaGetErrorString(e));
}

int b200rmsd_exp_rowsum_dev(const float* D, int64_t rows, int64_t cols, int64_t ld, float scale, int accumulate,
                            double* rowsum, void* stream)
{
    if (!D || !rowsum || rows < 0 || cols < 0 || ld < cols) return fail(B200RMSD_EINVAL, "exp_rowsum: bad arguments");
    if (rows == 0) return 0;
    int sm = 0;
    const int grid = launch_grid(rows, &sm);
    if (grid < 0) return fail(B200RMSD_ENODEVICE, "exp_rowsum: no CUDA device");
    b200::exp_rowsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(D, rows, cols, ld, scale, accumulate, rowsum);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail(B200RMSD_ECUDA, "exp_rowsum: %s", cudaGetErrorString(e));
}

int b200rmsd_row_argmin_dev(const float* D, int64_t rows, int64_t cols, int64_t ld, int32_t* arg, float* val, void* stream)
{
    if (!D || !arg || rows < 0 || cols <= 0 || ld < cols || cols > 0x7fffffff)
        return fail(B200RMSD_EINVAL, "row_argmin: bad arguments");
    if (rows == 0) return 0;
    int sm = 0;
    const int grid = launch_grid(rows, &sm);
    if (grid < 0) return fail(B200RMSD_ENODEVICE, "row_argmin: no CUDA device");
    b200::row_argmin_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(D, rows, cols, ld, arg, val);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail(B200RMSD_ECUDA, "row_argmin: %s", cudaGetErrorString(e));
}

int b200rmsd_condense_dev(const float* D, int64_t n, int64_t ld, float* out, void* stream)
{
    if (!D || (!out && n > 1) || n < 0 || ld < n) return fail(B200RMSD_EINVAL, "condense: bad arguments");
    if (n < 2) return 0;
    int sm = 0;
    if (launch_grid(n, &sm) < 0) return fail(B200RMSD_ENODEVICE, "condense: no CUDA device");
    const int64_t want = n - 1, cap = (int64_t)sm * 8;
    b200::condense_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(D, n, ld, out);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail(B200RMSD_ECUDA, "condense: %s", cudaGetErrorString(e));
}

}  // extern "C"
