// Microbenchmark behind profiles/r01_notes.md ("what bounds the kernels"): dependent / independent DFMA issue
// rates and the latency of one eval_chain / one gd_step_compact on 1..16 warps of ONE SM.
//   nvcc -O3 -std=c++17 --fmad=false -gencode arch=compute_100a,code=sm_100a -Ipick_ik_b200/csrc \
//        -o _scratch/eval_latency profiles/microbench/eval_latency.cu && ./_scratch/eval_latency
#include <cstdio>
#include <cstring>
#include "pik_device.cuh"
using namespace pik;
using Z7 = StaticSpec<7, 0x2222222ull, true, false>;
__global__ void lat_kernel(double* out, long long* cyc, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9;
    const double m = 0.9999999, c = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = fma(a, m, c); a = fma(a, m, c); a = fma(a, m, c); a = fma(a, m, c); }
    long long t1 = clock64();
    double b0=a,b1=a+1,b2=a+2,b3=a+3,b4=a+4,b5=a+5,b6=a+6,b7=a+7;
    for (int i = 0; i < iters; ++i) { b0=fma(b0,m,c); b1=fma(b1,m,c); b2=fma(b2,m,c); b3=fma(b3,m,c); b4=fma(b4,m,c); b5=fma(b5,m,c); b6=fma(b6,m,c); b7=fma(b7,m,c); }
    long long t2 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = a + b0+b1+b2+b3+b4+b5+b6+b7;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}
__global__ void eval_kernel(double* buf, long long* cyc, int iters, int mode) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    double* q = sm + lane + warp * 40 * 32;
    double* g = q + 7 * kS;
    double* sc = g + 7 * kS;
    double* g7 = sm + 40 * 32 * (blockDim.x / 32) + warp * 16;
    for (int j = 0; j < 7; ++j) { q[j * kS] = 0.1 * j - 0.3 + lane * 0.01; g[j * kS] = 1e-5 * j; }
    if (lane < 7) g7[lane] = lane == 3 ? 1.0 : 0.1 * lane;
    __syncwarp();
    double acc = eval_chain<Z7>(q, nullptr, kViewPlain, -1, 0.0, nullptr, sc, g7, buf, nullptr);
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
        if (mode == 0) acc += eval_chain<Z7>(q, nullptr, kViewFd, lane % 7, q[(lane % 7) * kS] + 1e-4, sc, nullptr, g7, buf, nullptr);
        else if (mode == 1) acc += eval_chain<Z7>(q, g, kViewMinus, -1, 0.0, nullptr, nullptr, g7, buf, nullptr);
        else acc += gd_step_compact<Z7, false>(q, g, sc, nullptr, g7, buf, nullptr);
    }
    long long t1 = clock64();
    buf[100 + threadIdx.x + blockIdx.x * blockDim.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (t1 - t0) / iters;
}
int main() {
    double* buf; long long* cyc;
    cudaMalloc(&buf, 1 << 24); cudaMemset(buf, 0, 1 << 24);
    cudaMallocManaged(&cyc, 64);
    DevRobot rb; memset(&rb, 0, sizeof(rb)); rb.n = 7; rb.has_tip = 1;
    for (int j = 0; j < 7; ++j) { rb.kind[j] = kRevZ; rb.bounded[j] = 1; rb.sign[j] = 1; for (int i = 0; i < 9; ++i) rb.R[j][i] = (i % 4 == 0) ? 0.8 : 0.1 * (i - 4); rb.t[j][0] = 0.1; rb.t[j][2] = 0.3; rb.vmin[j] = -2.8; rb.vmax[j] = 2.8; rb.vhalf[j] = 2.8; }
    for (int i = 0; i < 9; ++i) rb.tip_R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    DevParams pr; memset(&pr, 0, sizeof(pr)); pr.step_size = 1e-4; pr.position_scale = 1; pr.rotation_scale = 0.5; pr.position_threshold = 1e-3; pr.orientation_threshold = 1e-3;
    cudaMemcpyToSymbol(c_rb, &rb, sizeof(rb)); cudaMemcpyToSymbol(c_pr, &pr, sizeof(pr));
    lat_kernel<<<1, 32>>>(buf, cyc, 10000); cudaDeviceSynchronize();
    printf("DFMA dependent: %.2f cycles/instr; 8-way independent (1 warp): %.2f cycles/instr\n", cyc[0] / 40000.0, cyc[1] / 80000.0);
    lat_kernel<<<1, 128>>>(buf, cyc, 10000); cudaDeviceSynchronize();
    printf("4 warps (1/SMSP): dependent %.2f, 8-way %.2f\n", cyc[0] / 40000.0, cyc[1] / 80000.0);
    lat_kernel<<<1, 512>>>(buf, cyc, 10000); cudaDeviceSynchronize();
    printf("16 warps (4/SMSP): dependent %.2f, 8-way %.2f\n", cyc[0] / 40000.0, cyc[1] / 80000.0);
    cudaFuncSetAttribute(eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[3] = {"eval cached-FD", "eval inline-sincos", "gd_step_compact"};
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            size_t smem = (size_t)(40 * 32 * warps + 16 * warps) * 8;
            eval_kernel<<<1, 32 * warps, smem>>>(buf, cyc, 200, mode); cudaDeviceSynchronize();
            eval_kernel<<<1, 32 * warps, smem>>>(buf, cyc, 200, mode);
            cudaError_t e = cudaDeviceSynchronize();
            printf("%s: %d warps on one SM: %lld cycles per call (%s)\n", names[mode], warps, cyc[0], cudaGetErrorString(e));
        }
    return 0;
}
