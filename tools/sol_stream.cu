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

// Replica of the per-sample kernel's STORE pattern at (nx,nu,ny) = (2,1,2):
// 21 output blocks of C doubles per sample (63 doubles = 504 B per sample);
// per block a warp writes the 32*C contiguous doubles of its 32 samples,
// consecutive lanes -> consecutive addresses.  No loads, no shared memory, no
// arithmetic: what the store pattern alone can reach, with 8-byte (VEC = 1)
// and 16-byte (VEC = 2) stores.
__constant__ int kBlockC[21] = {2, 2, 2, 2, 4, 4, 4, 2, 4, 3, 4, 4, 2, 2, 3, 2, 2, 4, 4, 3, 4};

template <int VEC>
__global__ void __launch_bounds__(128) pattern_kernel(double* __restrict__ out, long long n_samples,
                                                      long long ntiles)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long kw = tile * 128 + warp * 32;
        long long base = 0;
#pragma unroll
        for (int b = 0; b < 21; ++b) {
            const int C = kBlockC[b];
            double* dst = out + base + kw * C;
            if (VEC == 1) {
                for (int it = 0; it < C; ++it) __stcs(dst + lane + it * 32, 1.0 + b);
            } else {
                double2* d2 = reinterpret_cast<double2*>(dst);
                for (int e = lane; e < 16 * C; e += 32) __stcs(d2 + e, make_double2(1.0 + b, 2.0));
            }
            base += n_samples * C;
        }
    }
}

static void run_pattern()
{
    const long long N = 1000000, ntiles = (N + 127) / 128;
    double* out; double* flush;
    CK(cudaMalloc(&out, (size_t)ntiles * 128 * 63 * 8));
    CK(cudaMalloc(&flush, 256 << 20));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const long long grids[] = {ntiles, 148 * 8, 148 * 16};
    for (int vec = 1; vec <= 2; ++vec)
        for (long long g : grids) {
            float best = 1e9f, sum = 0.f;
            for (int it = 0; it < 10; ++it) {
                CK(cudaMemsetAsync(flush, 0, 256 << 20));
                CK(cudaEventRecord(e0));
                if (vec == 1) pattern_kernel<1><<<(unsigned)g, 128>>>(out, ntiles * 128, ntiles);
                else pattern_kernel<2><<<(unsigned)g, 128>>>(out, ntiles * 128, ntiles);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (it >= 2) { if (ms < best) best = ms; sum += ms; }
            }
            printf("store pattern (2,1,2) %2d-byte stores grid=%6lld  best %.4f ms  mean %.4f ms  %.1f GB/s (504 B/sample)\n",
                   8 * vec, g, best, sum / 8, 504.0 * N / best / 1e6);
        }
    // empty kernel between two events: the fixed cost every event-timed kernel carries
    float best = 1e9f;
    for (int it = 0; it < 20; ++it) {
        CK(cudaMemsetAsync(flush, 0, 256 << 20));
        CK(cudaEventRecord(e0));
        pattern_kernel<1><<<1, 128>>>(out, 128, 0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    printf("empty kernel between two events: %.2f us\n", best * 1e3);
    cudaFree(out); cudaFree(flush);
}

int main()
{
    run_pattern();
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
