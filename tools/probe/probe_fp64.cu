// Throughput probe: raw DMMA.8x8x4 and DFMA issue rates on sm_100a (test infrastructure, not product).
#include <cuda_runtime.h>
#include <cstdio>
__global__ void dmma_loop(double* out, int iters){
  double a0=1.0+threadIdx.x*1e-9, b0=1.0-threadIdx.x*1e-9;
  double c[16]; 
  #pragma unroll
  for(int i=0;i<16;i++)c[i]=0;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<8;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c[2*i]),"+d"(c[2*i+1]):"d"(a0),"d"(b0));
  }
  double s=0; 
  #pragma unroll
  for(int i=0;i<16;i++)s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void dfma_loop(double* out, int iters){
  double a0=1.0+threadIdx.x*1e-9, b0=1e-9;
  double c[16];
  #pragma unroll
  for(int i=0;i<16;i++)c[i]=i;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<16;i++) c[i]=fma(c[i],a0,b0);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<16;i++)s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  int nsm=0; cudaDeviceGetAttribute(&nsm,cudaDevAttrMultiProcessorCount,0);
  double* out; cudaMalloc(&out, sizeof(double)*nsm*8*1024);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int warps=4; warps<=32; warps*=2){
    for(int bps=1;bps<=2;bps++){
      int iters=20000; float ms;
      dmma_loop<<<nsm*bps,warps*32>>>(out,100); cudaDeviceSynchronize();
      cudaEventRecord(e0); dmma_loop<<<nsm*bps,warps*32>>>(out,iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms,e0,e1);
      double flops=2.0*8*8*4*8*(double)iters*warps*nsm*bps;
      printf("DMMA warps/cta=%d cta/sm=%d: %.2f TFLOP/s (%.3f ms)\n",warps,bps,flops/ms*1e-9,ms);
      dfma_loop<<<nsm*bps,warps*32>>>(out,100); cudaDeviceSynchronize();
      cudaEventRecord(e0); dfma_loop<<<nsm*bps,warps*32>>>(out,iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms,e0,e1);
      flops=2.0*16*(double)iters*warps*32*nsm*bps;
      printf("DFMA warps/cta=%d cta/sm=%d: %.2f TFLOP/s (%.3f ms)\n",warps,bps,flops/ms*1e-9,ms);
    }
  }
  printf("cuda err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
