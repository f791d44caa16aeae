// Packed-FP32 microbenchmark (B200): FFMA vs FFMA2 throughput per SM, with constant / register operands and with ALU
// work mixed in.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu; result: profiles/r02_microbench_ffma2.log
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b){unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(unsigned long long v, float&a, float&b){asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v));}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
template<int MODE> __global__ void k(float* out, float a, float b, int iters){
  float x[16];
  for(int i=0;i<16;i++) x[i]=threadIdx.x*0.001f+i;
  if (MODE==0){
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<16;i++) x[i]=fmaf(x[i],a,b);
    }
  } else if (MODE==1) {
    unsigned long long p[8]; unsigned long long pa=pk(a,a), pb=pk(b,b);
    for(int i=0;i<8;i++) p[i]=pk(x[2*i],x[2*i+1]);
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<8;i++) p[i]=fma2(p[i],pa,pb);
    }
    for(int i=0;i<8;i++) upk(p[i],x[2*i],x[2*i+1]);
  } else if (MODE==2) { // FFMA with 3 distinct regs each (x = x*y + z) pattern
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<16;i++) x[i]=fmaf(x[i],x[(i+1)&15],x[(i+2)&15]);
    }
  } else if (MODE==3) { // FFMA2 with distinct regs
    unsigned long long p[8];
    for(int i=0;i<8;i++) p[i]=pk(x[2*i],x[2*i+1]);
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<8;i++) p[i]=fma2(p[i],p[(i+1)&7],p[(i+2)&7]);
    }
    for(int i=0;i<8;i++) upk(p[i],x[2*i],x[2*i+1]);
  } else if (MODE==4) { // FFMA + IADD3/LOP mix 1:1
    unsigned y[16]; for(int i=0;i<16;i++) y[i]=threadIdx.x+i;
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<16;i++){ x[i]=fmaf(x[i],a,b); y[i]=(y[i]^(y[(i+1)&15]))+it; }
    }
    for(int i=0;i<16;i++) x[i]+=__uint_as_float(y[i]&0x3fffff);
  } else if (MODE==5) { // FFMA2 + alu mix (same FMA count and alu count as mode 4)
    unsigned y[16]; for(int i=0;i<16;i++) y[i]=threadIdx.x+i;
    unsigned long long p[8]; unsigned long long pa=pk(a,a), pb=pk(b,b);
    for(int i=0;i<8;i++) p[i]=pk(x[2*i],x[2*i+1]);
    for(int it=0;it<iters;it++){
#pragma unroll
      for(int i=0;i<8;i++){ p[i]=fma2(p[i],pa,pb); y[2*i]=(y[2*i]^(y[(2*i+1)&15]))+it; y[2*i+1]=(y[2*i+1]^(y[(2*i+2)&15]))+it; }
    }
    for(int i=0;i<8;i++) upk(p[i],x[2*i],x[2*i+1]);
    for(int i=0;i<16;i++) x[i]+=__uint_as_float(y[i]&0x3fffff);
  }
  float s=0; for(int i=0;i<16;i++) s+=x[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name, int warps_per_sm){
  int blocks=148*warps_per_sm/8; int iters=4096; float* out; cudaMalloc(&out, blocks*256*4);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks,256>>>(out,1.0001f,0.5f,iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<blocks,256>>>(out,1.0001f,0.5f,iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double fma=(double)blocks*256*16*iters;
  printf("%-28s warps/SM %2d: %.3f ms  %.1f GFMA/s  = %.1f FMA/clk/SM @1.965GHz\n",name,warps_per_sm,ms,fma/ms*1e-6,fma/ms*1e-6/148/1.965);
  cudaFree(out);
}
int main(){
  for(int w: {16,32,64}){
    if(w==16){run<0>("FFMA imm/const",16);run<1>("FFMA2 const",16);run<2>("FFMA 3reg",16);run<3>("FFMA2 3reg",16);run<4>("FFMA+2alu",16);run<5>("FFMA2+2alu",16);}
    if(w==32){run<0>("FFMA imm/const",32);run<1>("FFMA2 const",32);run<2>("FFMA 3reg",32);run<3>("FFMA2 3reg",32);run<4>("FFMA+2alu",32);run<5>("FFMA2+2alu",32);}
    if(w==64){run<0>("FFMA imm/const",64);run<1>("FFMA2 const",64);run<2>("FFMA 3reg",64);run<3>("FFMA2 3reg",64);run<4>("FFMA+2alu",64);run<5>("FFMA2+2alu",64);}
  }
}
