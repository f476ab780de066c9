// latency / throughput probes for the ops on the Jacobi dependency chain (development aid)
#include <cstdio>
#define N 512
template <int OP> __device__ double chain(double x, double y) {
    float f = (float)x; 
    #pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, y, y);                       // DFMA
        if (OP == 1) x = x * y;                              // DMUL
        if (OP == 2) x = x + y;                              // DADD
        if (OP == 3) { f = (float)x; x = (double)f + y; }    // F2F down + F2F up + DADD
        if (OP == 4) { x = __shfl_xor_sync(0xffffffffu, x, 1) + y; }   // SHFL64 + DADD
        if (OP == 5) { asm("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f)); }  // MUFU
        if (OP == 6) { f = fmaf(f, f, 1.0f); }               // FFMA
        if (OP == 7) { x = (double)__double2float_rn(x); }    // F2F both
    }
    return x + f;
}
template <int OP> __global__ void k(double* out, long long* cyc, double y) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
    x = chain<OP>(x, y);
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP> void run(const char* name, double* d, long long* c) {
    long long h;
    for (int threads : {32, 128, 256, 512, 1024}) {
        k<OP><<<1, threads>>>(d, c, 1.0000001); cudaDeviceSynchronize();
        k<OP><<<1, threads>>>(d, c, 1.0000001); cudaDeviceSynchronize();
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%-28s threads=%4d: %.1f cycles/op/warp-chain, %.2f warp-ops/clk/SM\n", name, threads, (double)h / N, (threads / 32.0) * N / (double)h);
    }
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 8 * 2048); cudaMemset(d, 0, 8 * 2048); cudaMalloc(&c, 8);
    run<0>("DFMA dependent", d, c); run<1>("DMUL dependent", d, c); run<2>("DADD dependent", d, c);
    run<3>("F2F down+up+DADD", d, c); run<4>("SHFL64+DADD", d, c); run<5>("MUFU.RSQ", d, c); run<6>("FFMA", d, c); run<7>("F2F down+up", d, c);
    return 0;
}
