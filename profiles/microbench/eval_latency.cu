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
// pieces of one evaluation, out of line: the chain walk alone (cached sin/cos) and the pose cost alone
__device__ __noinline__ double chain_only(const double* q, const double* sc) {
    Frame F;
    frame_load_origin(F, 0);
#pragma unroll 1
    for (int j = 0; j <= 7; ++j) {
        if (j > 0) frame_mul_origin<Z7>(F, j);
        if (j == 7) break;
        joint_one_kind<kRevZ, false>(F, j, kRevZ, q[j * kS], sc[(2 * j) * kS], sc[(2 * j + 1) * kS]);
    }
    return ((F.r[0] + F.r[4]) + F.r[8]) + ((F.t[0] + F.t[1]) + F.t[2]) + ((F.r[1] + F.r[2]) + (F.r[3] + F.r[5])) + (F.r[6] + F.r[7]);
}
__device__ __noinline__ double cost_only(const double* g7, double a, double b) {
    Frame F;
    const double ca = a, sa = b;  // some rotation-like frame built from two inputs
    F.r[0] = ca; F.r[1] = -sa; F.r[2] = 0.1 * b; F.r[3] = sa; F.r[4] = ca; F.r[5] = -0.1 * a; F.r[6] = -0.05; F.r[7] = 0.08; F.r[8] = 0.99;
    F.t[0] = a; F.t[1] = b; F.t[2] = a * b;
    double dist, ang;
    return pose_cost_one(g7, F, dist, ang);
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
        q[0] = 0.1 * 0 - 0.3 + lane * 0.01 + acc * 1e-300;  // every call depends on the previous one
        sc[0] = sc[0] + acc * 1e-300;
        if (mode == 0) acc += eval_chain<Z7>(q, nullptr, kViewFd, lane % 7, q[(lane % 7) * kS] + 1e-4, sc, nullptr, g7, buf, nullptr);
        else if (mode == 1) acc += eval_chain<Z7>(q, g, kViewMinus, -1, 0.0, nullptr, nullptr, g7, buf, nullptr);
        else if (mode == 2) acc += gd_step_compact<Z7, false>(q, g, sc, nullptr, g7, buf, nullptr);
        else if (mode == 3) acc += chain_only(q, sc) * 1e-3;
        else if (mode == 4) acc += cost_only(g7, 0.8 + 1e-9 * acc, 0.6);
        else { const CostPair cp = pair_costs_from_origin<Z7>(kPairFd, lane % 8 == 7 ? -1 : lane % 8, q, nullptr, sc, g7, buf); acc += cp.m + cp.p; }
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
    const char* names[6] = {"eval cached-FD", "eval inline-sincos", "gd_step_compact", "chain walk only (cached sin/cos)", "pose cost only", "pair from origin (wide round A)"};
    for (int mode = 0; mode < 6; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            size_t smem = (size_t)(40 * 32 * warps + 16 * warps) * 8;
            eval_kernel<<<1, 32 * warps, smem>>>(buf, cyc, 200, mode); cudaDeviceSynchronize();
            eval_kernel<<<1, 32 * warps, smem>>>(buf, cyc, 200, mode);
            cudaError_t e = cudaDeviceSynchronize();
            printf("%s: %d warps on one SM: %lld cycles per call (%s)\n", names[mode], warps, cyc[0], cudaGetErrorString(e));
        }
    return 0;
}
