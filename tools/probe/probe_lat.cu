// Latency probe for the primitives on the LDL^T tile kernel's critical path (test infrastructure).
#include <cuda_runtime.h>
#include <cstdio>
#define N 256
__global__ void lat(long long* out, double* sink, int nthreads_active) {
    __shared__ double sm[1024];
    __shared__ int smi[64];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 1024; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    if (tid < 64) smi[tid] = (tid + 1) & 63;
    __syncthreads();
    double x = 1.0 + tid * 1e-12, y = 1.0000001;
    long long t0, t1;
    // dependent DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, y, 1e-9);
    t1 = clock64();
    if (tid == 0) out[0] = (t1 - t0);
    // dependent DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x * y;
    t1 = clock64();
    if (tid == 0) out[1] = (t1 - t0);
    // dependent LDS chain (pointer chasing)
    int p = lane & 63;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) p = smi[p];
    t1 = clock64();
    if (tid == 0) out[2] = (t1 - t0);
    // dependent shfl chain
    int q = p;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) q = __shfl_sync(0xffffffffu, q, (q + 1) & 31);
    t1 = clock64();
    if (tid == 0) out[3] = (t1 - t0);
    // dependent vote chain
    unsigned v = q;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __any_sync(0xffffffffu, (v + lane) & 1) + v;
    t1 = clock64();
    if (tid == 0) out[4] = (t1 - t0);
    // dependent redux chain
    unsigned r = v + lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) r = __reduce_max_sync(0xffffffffu, r + lane) >> 1;
    t1 = clock64();
    if (tid == 0) out[5] = (t1 - t0);
    // dependent rcp.approx.f64 chain
    double z = x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { double rr; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(z)); z = rr + 1.5; }
    t1 = clock64();
    if (tid == 0) out[6] = (t1 - t0);
    // exact double division chain
    double w = z;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) w = 1.0 / w + 1.5;
    t1 = clock64();
    if (tid == 0) out[7] = (t1 - t0);
    // barrier round trips
    __syncthreads();
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[8] = (t1 - t0);
    // barrier + sts + lds handoff (one warp writes, all read)
    t0 = clock64();
    double acc = 0;
    for (int i = 0; i < N; i++) {
        if (tid == (i & 7) * 32) sm[i & 63] = acc + i;
        __syncthreads();
        acc += sm[i & 63];
    }
    t1 = clock64();
    if (tid == 0) out[9] = (t1 - t0);
    // DSETP + select chain
    double c = acc;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) c = (c > 1.0) ? c * 0.5 : c + 3.0;
    t1 = clock64();
    if (tid == 0) out[10] = (t1 - t0);
    sink[blockIdx.x * blockDim.x + tid] = x + z + w + q + v + r + p + acc + c;
}
int main() {
    long long* out; double* sink;
    cudaMallocManaged(&out, 16 * sizeof(long long)); cudaMalloc(&sink, 4096 * sizeof(double));
    const char* names[] = {"DFMA dep", "DMUL dep", "LDS dep (ptr chase)", "SHFL dep", "VOTE.any dep", "REDUX.max dep", "rcp.approx.f64+DADD", "1.0/x exact + DADD", "__syncthreads", "STS->bar->LDS handoff", "DSETP+select+DMUL/DADD"};
    for (int threads : {32, 256, 1024}) {
        lat<<<1, threads>>>(out, sink, threads); cudaDeviceSynchronize();
        lat<<<1, threads>>>(out, sink, threads); cudaDeviceSynchronize();
        printf("--- %d threads/CTA (cycles per op, warp 0)\n", threads);
        for (int i = 0; i < 11; i++) printf("%-26s %7.1f\n", names[i], (double)out[i] / N);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
