#include <cstdio>
#include <cuda_runtime.h>
#include "../../cap_b200/csrc/fp.cuh"
using namespace capgpu;
template<int CH, bool RR> __global__ void k(Fq* out, const Fq* yin, int iters){
  Fq y = yin[threadIdx.x & 1];
  Fq x[CH];
  for(int c=0;c<CH;c++){ x[c]=Fq::one(); x[c].v[0]+=threadIdx.x+c; }
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int c=0;c<CH;c++) x[c]= RR ? fp_mul_rr(x[c],y) : fp_mul(x[c],y);
  }
  Fq s=x[0]; for(int c=1;c<CH;c++) s=fp_add(s,x[c]);
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<class F> double run(F f, double muls){ cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best=1e9;
  for(int r=0;r<3;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  return muls/best*1e-6; }
int main(){
  int threads=256, iters=256; Fq* out; Fq* yin; cudaMalloc(&out,148*16*256*sizeof(Fq)); cudaMalloc(&yin,2*sizeof(Fq));
  Fq h[2]; h[0]=Fq::r2(); h[1]=Fq::one(); cudaMemcpy(yin,h,sizeof h,cudaMemcpyHostToDevice);
  // correctness cross-check on device
  for(int bps : {2,4,8}){
    int blocks=148*bps; double n=(double)blocks*threads*iters;
    printf("warps/SM=%d  cios ch1 %.1f ch2 %.1f | rr ch1 %.1f ch2 %.1f ch4 %.1f G mul/s\n", bps*8,
      run([&]{k<1,false><<<blocks,threads>>>(out,yin,iters);},n*1), run([&]{k<2,false><<<blocks,threads>>>(out,yin,iters);},n*2),
      run([&]{k<1,true><<<blocks,threads>>>(out,yin,iters);},n*1), run([&]{k<2,true><<<blocks,threads>>>(out,yin,iters);},n*2), run([&]{k<4,true><<<blocks,threads>>>(out,yin,iters);},n*4));
  }
  return 0; }
