// microbenchmark: DFMA rate, IMAD.WIDE rate, and both interleaved in the same warp
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, double m, int iters){
  double a0=threadIdx.x,a1=a0+1,a2=a0+2,a3=a0+3,a4=a0+4,a5=a0+5,a6=a0+6,a7=a0+7;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int u=0;u<8;u++){
      a0=__fma_rz(a0,m,a0);a1=__fma_rz(a1,m,a1);a2=__fma_rz(a2,m,a2);a3=__fma_rz(a3,m,a3);
      a4=__fma_rz(a4,m,a4);a5=__fma_rz(a5,m,a5);a6=__fma_rz(a6,m,a6);a7=__fma_rz(a7,m,a7);
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=a0+a1+a2+a3+a4+a5+a6+a7;
}
__global__ void k_imad(uint64_t* out, uint32_t m, int iters){
  uint64_t a0=threadIdx.x,a1=a0+1,a2=a0+2,a3=a0+3,a4=a0+4,a5=a0+5,a6=a0+6,a7=a0+7;
  uint32_t b0=m^threadIdx.x,b1=b0+11,b2=b0+22,b3=b0+33;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int u=0;u<8;u++){
      asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a0):"r"(b0),"r"(m));asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a1):"r"(b1),"r"(m));
      asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a2):"r"(b2),"r"(m));asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a3):"r"(b3),"r"(m));
      asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a4):"r"(b0),"r"(b1));asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a5):"r"(b1),"r"(b2));
      asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a6):"r"(b2),"r"(b3));asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a7):"r"(b3),"r"(b0));
    }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=a0^a1^a2^a3^a4^a5^a6^a7;
}
__global__ void k_both(double* outd, uint64_t* outi, double md, uint32_t m, int iters){
  double d0=threadIdx.x,d1=d0+1,d2=d0+2,d3=d0+3;
  uint64_t a0=threadIdx.x,a1=a0+1,a2=a0+2,a3=a0+3;
  uint32_t b0=m^threadIdx.x,b1=b0+11,b2=b0+22,b3=b0+33;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int u=0;u<8;u++){
      d0=__fma_rz(d0,md,d0); asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a0):"r"(b0),"r"(m));
      d1=__fma_rz(d1,md,d1); asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a1):"r"(b1),"r"(m));
      d2=__fma_rz(d2,md,d2); asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a2):"r"(b2),"r"(m));
      d3=__fma_rz(d3,md,d3); asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(a3):"r"(b3),"r"(m));
    }
  }
  outd[blockIdx.x*blockDim.x+threadIdx.x]=d0+d1+d2+d3; outi[blockIdx.x*blockDim.x+threadIdx.x]=a0^a1^a2^a3;
}
int main(){
  int blocks=148*8, threads=256, iters=2048; void *b1,*b2; cudaMalloc(&b1,blocks*threads*8); cudaMalloc(&b2,blocks*threads*8);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  for(int rep=0;rep<3;rep++){
    cudaEventRecord(e0); k_dfma<<<blocks,threads>>>((double*)b1,1.0000001,iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("dfma  %.1f G lane-ops/s\n",(double)blocks*threads*iters*64/ms*1e-6);
    cudaEventRecord(e0); k_imad<<<blocks,threads>>>((uint64_t*)b2,0x9e3779b1u,iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("imadw %.1f G lane-ops/s\n",(double)blocks*threads*iters*64/ms*1e-6);
    cudaEventRecord(e0); k_both<<<blocks,threads>>>((double*)b1,(uint64_t*)b2,1.0000001,0x9e3779b1u,iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("both  %.1f G dfma + %.1f G imadw lane-ops/s\n",(double)blocks*threads*iters*32/ms*1e-6,(double)blocks*threads*iters*32/ms*1e-6);
  }
  return 0;
}
