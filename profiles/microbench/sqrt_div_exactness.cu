// sqrt_fast / div_fast (pik_device.cuh: the nvcc fast-path sequences used two-wide in the pair cost) against
// sqrt() and / on random bit patterns and on the magnitudes the path sees; prints the mismatch counts (0).
//   nvcc -O3 -std=c++17 --fmad=false -gencode arch=compute_100a,code=sm_100a -Ipick_ik_b200/csrc \
//        -o _scratch/sqrt_div_exactness profiles/microbench/sqrt_div_exactness.cu && ./_scratch/sqrt_div_exactness
#include <cstdio>
#include "pik_device.cuh"
using namespace pik;
__device__ unsigned long long splitmix(unsigned long long& s) { unsigned long long z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
__global__ void check_kernel(unsigned long long* bad, int iters, int mode) {
    unsigned long long s = (blockIdx.x * 1024ull + threadIdx.x) * 7919ull + 12345ull + mode * 1000003ull;
    unsigned long long nb_s = 0, nb_d = 0, fast_s = 0, fast_d = 0;
    for (int i = 0; i < iters; ++i) {
        double x, a, b;
        if (mode == 0) {  // arbitrary bit patterns
            x = __longlong_as_double((long long)splitmix(s)); a = __longlong_as_double((long long)splitmix(s)); b = __longlong_as_double((long long)splitmix(s));
        } else {  // magnitudes the path sees: [1e-20, 1e3]
            const double u = (double)(splitmix(s) >> 11) * 0x1.0p-53, v = (double)(splitmix(s) >> 11) * 0x1.0p-53, w = (double)(splitmix(s) >> 11) * 0x1.0p-53;
            x = exp2(u * 76.0 - 66.0) * (1.0 + v); a = exp2(v * 76.0 - 66.0) * (1.0 + w) * ((splitmix(s) & 1) ? -1.0 : 1.0); b = exp2(w * 76.0 - 66.0) * (1.0 + u);
        }
        bool ok;
        const double r = sqrt_fast(x, ok);
        if (ok) { ++fast_s; if (__double_as_longlong(r) != __double_as_longlong(sqrt(x))) ++nb_s; }
        const double q = div_fast(a, b, ok);
        if (ok) { ++fast_d; if (__double_as_longlong(q) != __double_as_longlong(a / b)) ++nb_d; }
    }
    atomicAdd(&bad[0], nb_s); atomicAdd(&bad[1], nb_d); atomicAdd(&bad[2], fast_s); atomicAdd(&bad[3], fast_d);
}
int main() {
    unsigned long long* bad; cudaMallocManaged(&bad, 64);
    for (int mode = 0; mode < 2; ++mode) {
        for (int i = 0; i < 4; ++i) bad[i] = 0;
        check_kernel<<<148 * 8, 256>>>(bad, 4000, mode);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d: sqrt mismatches %llu of %llu fast-path; div mismatches %llu of %llu fast-path (%s)\n", mode, bad[0], bad[2], bad[1], bad[3], cudaGetErrorString(e));
    }
    return 0;
}
