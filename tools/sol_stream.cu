// Speed-of-light calibration for the write-dominated traffic mix of the fused
// per-sample kernel: plain grid-stride streaming kernels with the same
// read:write byte ratio (no arithmetic, no shared memory).  Prints GB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/sol_stream tools/sol_stream.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// every thread reads `r` of each `r + w` double2 it handles and writes `w`
template <int VEC>
__global__ void mix_kernel(const double2* __restrict__ in, double2* __restrict__ out,
                           size_t n_out, int ratio)
{
    // n_out double2 written; one double2 read per `ratio` written
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n_out; i += stride) {
        double2 v = make_double2(1.0, 2.0);
        if (ratio > 0 && (i % ratio) == 0) v = __ldcs(in + i / ratio);
        __stcs(out + i, v);
    }
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) __stcs(out + i, __ldcs(in + i));
}

int main()
{
    const size_t out_bytes = 504ull << 20, in_bytes = 512ull << 20;
    double2 *in, *out, *flush;
    CK(cudaMalloc(&in, in_bytes));
    CK(cudaMalloc(&out, in_bytes));
    CK(cudaMalloc(&flush, 256 << 20));
    CK(cudaMemset(in, 0, in_bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grids[] = {148 * 4, 148 * 8, 148 * 16, 148 * 64};
    for (int mode = 0; mode < 3; ++mode) {
        for (int g : grids) {
            float best = 1e9f;
            for (int it = 0; it < 8; ++it) {
                CK(cudaMemsetAsync(flush, 0, 256 << 20));
                CK(cudaEventRecord(e0));
                if (mode == 0) mix_kernel<2><<<g, 256>>>(in, out, out_bytes / 16, 0);
                else if (mode == 1) mix_kernel<2><<<g, 256>>>(in, out, out_bytes / 16, 6);
                else copy_kernel<<<g, 256>>>(in, out, in_bytes / 16);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (it >= 2 && ms < best) best = ms;
            }
            double bytes = mode == 0 ? (double)out_bytes
                         : mode == 1 ? out_bytes * (1.0 + 1.0 / 6) : 2.0 * in_bytes;
            printf("%s grid=%5d  %.4f ms  %.1f GB/s\n",
                   mode == 0 ? "write-only      " : mode == 1 ? "read1:write6 mix" : "copy            ",
                   g, best, bytes / best / 1e6);
        }
    }
    return 0;
}
