#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a,u64 b,u64 c){ u64 d; asm("fma.rn.f32x2 %0,%1,%2,%3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
template<int MODE> __global__ void k(float* out, int iters, float x0){
  float a0=x0+threadIdx.x, a1=a0+1, a2=a0+2,a3=a0+3,a4=a0+4,a5=a0+5,a6=a0+6,a7=a0+7;
  const float m=0.999f, c=0.001f;
  if (MODE==0){
    for(int i=0;i<iters;++i){
      a0=fmaf(a0,m,c);a1=fmaf(a1,m,c);a2=fmaf(a2,m,c);a3=fmaf(a3,m,c);a4=fmaf(a4,m,c);a5=fmaf(a5,m,c);a6=fmaf(a6,m,c);a7=fmaf(a7,m,c);
    }
  } else if (MODE==1) {
    u64 p0=pk(a0,a1),p1=pk(a2,a3),p2=pk(a4,a5),p3=pk(a6,a7); u64 M=pk(m,m),C=pk(c,c);
    for(int i=0;i<iters;++i){ p0=fma2(p0,M,C);p1=fma2(p1,M,C);p2=fma2(p2,M,C);p3=fma2(p3,M,C);}
    upk(p0,a0,a1);upk(p1,a2,a3);upk(p2,a4,a5);upk(p3,a6,a7);
  } else { // mixed: 4 fma2 + 4 alu (lop3) per iter
    u64 p0=pk(a0,a1),p1=pk(a2,a3),p2=pk(a4,a5),p3=pk(a6,a7); u64 M=pk(m,m),C=pk(c,c);
    unsigned q0=threadIdx.x,q1=q0*3,q2=q0*5,q3=q0*7;
    for(int i=0;i<iters;++i){ p0=fma2(p0,M,C);q0=(q0^(q1<<3))+0x9e3779b9u;p1=fma2(p1,M,C);q1=(q1^(q2<<5))+0x7f4a7c15u;p2=fma2(p2,M,C);q2=(q2^(q3<<7))+0x85ebca6bu;p3=fma2(p3,M,C);q3=(q3^(q0<<9))+0xc2b2ae35u;}
    upk(p0,a0,a1);upk(p1,a2,a3);upk(p2,a4,a5);upk(p3,a6,a7); a0+=__uint_as_float((q0^q1^q2^q3)&0x3fffff);
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=a0+a1+a2+a3+a4+a5+a6+a7;
}
template<int MODE> void run(const char* name, float* d, int flops_per_iter){
  int iters=20000; cudaEvent_t e0,e1; cudaEventCreate(&e0);cudaEventCreate(&e1);
  k<MODE><<<148*8,256>>>(d,100,1.f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148*8,256>>>(d,iters,1.f); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double fmas=(double)148*8*256*iters*8; 
  printf("%s: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s), per-SM-per-clk@1.965GHz: %.1f FMA lanes\n",name,ms,fmas/ms/1e9,2*fmas/ms/1e9,fmas/(ms*1e-3)/148/1.965e9);
}
__device__ __forceinline__ u64 mul2(u64 a,u64 b){ u64 d; asm("mul.rn.f32x2 %0,%1,%2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ u64 add2(u64 a,u64 b){ u64 d; asm("add.rn.f32x2 %0,%1,%2;":"=l"(d):"l"(a),"l"(b)); return d;}
// MODE 0: FMUL2 throughput, 1: FADD2 throughput, 2: FFMA2 dependent chain (latency), 3: FFMA scalar dependent chain
template<int MODE> __global__ void k2(float* out, int iters, float x0){
  float a0=x0+threadIdx.x*1e-3f;
  u64 p0=pk(a0,a0+1),p1=pk(a0+2,a0+3),p2=pk(a0+4,a0+5),p3=pk(a0+6,a0+7); const u64 M=pk(0.9999f,1.0001f), C=pk(1e-4f,-1e-4f);
  float s0=a0;
  for(int i=0;i<iters;++i){
    if (MODE==0){p0=mul2(p0,M);p1=mul2(p1,M);p2=mul2(p2,M);p3=mul2(p3,M);}
    if (MODE==1){p0=add2(p0,C);p1=add2(p1,C);p2=add2(p2,C);p3=add2(p3,C);}
    if (MODE==2){p0=fma2(p0,M,C);p0=fma2(p0,M,C);p0=fma2(p0,M,C);p0=fma2(p0,M,C);}
    if (MODE==3){s0=fmaf(s0,0.9999f,1e-4f);s0=fmaf(s0,0.9999f,1e-4f);s0=fmaf(s0,0.9999f,1e-4f);s0=fmaf(s0,0.9999f,1e-4f);}
  }
  float b0,b1,b2,b3,b4,b5,b6,b7; upk(p0,b0,b1);upk(p1,b2,b3);upk(p2,b4,b5);upk(p3,b6,b7);
  out[blockIdx.x*blockDim.x+threadIdx.x]=b0+b1+b2+b3+b4+b5+b6+b7+s0;
}
// 3 distinct register operands per FFMA2 / FFMA (register-file bandwidth): 8 independent accumulators, operands rotate
template<int MODE> __global__ void k3(float* out, int iters, float x0){
  float a0=x0+threadIdx.x*1e-3f;
  if (MODE==0){
    u64 p[8]; for(int j=0;j<8;++j) p[j]=pk(a0+j,a0-j);
    for(int i=0;i<iters;++i){
      #pragma unroll
      for(int j=0;j<8;++j) p[j]=fma2(p[(j+1)&7],p[(j+3)&7],p[j]);
    }
    float s=0; for(int j=0;j<8;++j){float b0,b1; upk(p[j],b0,b1); s+=b0+b1;} out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  } else {
    float p[16]; for(int j=0;j<16;++j) p[j]=a0+j;
    for(int i=0;i<iters;++i){
      #pragma unroll
      for(int j=0;j<16;++j) p[j]=fmaf(p[(j+1)&15],p[(j+3)&15],p[j]);
    }
    float s=0; for(int j=0;j<16;++j) s+=p[j]; out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  }
}
template<int MODE> void run3(const char* name, float* d){
  int iters=5000; cudaEvent_t e0,e1; cudaEventCreate(&e0);cudaEventCreate(&e1);
  k3<MODE><<<148*4,256>>>(d,100,1e-3f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k3<MODE><<<148*4,256>>>(d,iters,1e-3f); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double lane_fmas=(double)148*4*256*iters*16;
  printf("%s: %.3f ms -> %.1f FMA lanes per SM per clk @1.965GHz\n",name,ms,lane_fmas/(ms*1e-3)/148/1.965e9);
}
template<int MODE> void run2(const char* name, float* d, int blocks, int threads){
  int iters=20000; cudaEvent_t e0,e1; cudaEventCreate(&e0);cudaEventCreate(&e1);
  k2<MODE><<<blocks,threads>>>(d,100,1.f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k2<MODE><<<blocks,threads>>>(d,iters,1.f); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  printf("%s: %.3f ms -> %.2f cycles per loop iteration of 4 instr per warp @1.965GHz (blocks=%d threads=%d)\n",name,ms,ms*1e-3*1.965e9/iters,blocks,threads);
}
// ---- math peaks the rooflines refer to: FP64 FMA, MUFU (rcp / sin), I2F, measured with 8 independent chains
// MODE 0: DFMA, 1: MUFU.RCP, 2: MUFU.SIN (via __sinf's sin.approx), 3: I2F.F32.U32
template<int MODE> __global__ void k4(float* out, int iters, float x0){
  double d[8]; float f[8]; unsigned u[8];
  for(int j=0;j<8;++j){ d[j]=x0+threadIdx.x+j; f[j]=1.0f+1e-3f*(threadIdx.x+j); u[j]=threadIdx.x*7+j; }
  for(int i=0;i<iters;++i){
    #pragma unroll
    for(int j=0;j<8;++j){
      if (MODE==0) d[j]=fma(d[j],0.999,0.001);
      if (MODE==1) { asm volatile("rcp.approx.ftz.f32 %0, %0;":"+f"(f[j])); f[j]+=0.5f; }  // (rcp(rcp(x)) alone is folded away)
      if (MODE==2) asm volatile("sin.approx.ftz.f32 %0, %0;":"+f"(f[j]));
      if (MODE==3){ float t; asm volatile("cvt.rn.f32.u32 %0, %1;":"=f"(t):"r"(u[j])); u[j]=__float_as_uint(t)>>3; }
    }
  }
  float s=0; for(int j=0;j<8;++j) s+=(float)d[j]+f[j]+(float)u[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> double run4(const char* name, float* d){
  int iters=4000; cudaEvent_t e0,e1; cudaEventCreate(&e0);cudaEventCreate(&e1);
  k4<MODE><<<148*8,256>>>(d,50,1.f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k4<MODE><<<148*8,256>>>(d,iters,1.f); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double ops=(double)148*8*256*iters*8, rate=ops/(ms*1e-3);
  printf("%s: %.3f ms -> %.3f T lane-ops/s, %.1f lanes per SM per clk @1.965GHz\n",name,ms,rate/1e12,rate/148/1.965e9);
  return rate;
}
int main(int argc, char** argv){ { float* d; cudaMalloc(&d,148*8*256*4);
  run3<0>("FFMA2, 3 distinct reg operands, 8 warps/SMSP",d); run3<1>("FFMA, 3 distinct reg operands, 8 warps/SMSP",d);
  run2<0>("FMUL2 x4 indep, 8 CTAs/SM",d,148*8,256); run2<1>("FADD2 x4 indep, 8 CTAs/SM",d,148*8,256);
  run2<2>("FFMA2 chain, 1 warp/SMSP",d,148,128); run2<3>("FFMA chain, 1 warp/SMSP",d,148,128);
  run2<2>("FFMA2 chain, 4 warps/SMSP",d,148,512); run2<3>("FFMA chain, 4 warps/SMSP",d,148,512);
  run2<2>("FFMA2 chain, 8 warps/SMSP",d,148,1024); } float* d; cudaMalloc(&d,148*8*256*4); run<0>("FFMA scalar",d,8); run<1>("FFMA2 packed",d,8); run<2>("FFMA2 + 1:1 ALU",d,8);
  double dfma=run4<0>("DFMA (FP64), 8 chains",d), rcp=run4<1>("MUFU.RCP, 8 chains",d), sn=run4<2>("MUFU.SIN, 8 chains",d), i2f=run4<3>("I2F.F32.U32, 8 chains",d);
  if (argc > 1) { FILE* f=fopen(argv[1],"w"); if (f) { fprintf(f,"{\n \"fp64_fma_lane_ops_per_s\": %.4e,\n \"fp64_tflops\": %.2f,\n \"mufu_rcp_per_s\": %.4e,\n \"mufu_sin_per_s\": %.4e,\n \"i2f_per_s\": %.4e,\n \"how\": \"tools/f2bench/ffma2_bench: 148*8 CTAs x 256 threads, 8 independent chains per thread, CUDA events, natural clocks\"\n}\n", dfma, 2*dfma/1e12, rcp, sn, i2f); fclose(f);} }
  return 0; }
