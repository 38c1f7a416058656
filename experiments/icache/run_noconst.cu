#include <cuda.h>
#include <cstdio>
#include <vector>
#define CK(x) do{CUresult r=(x); if(r){const char*s;cuGetErrorString(r,&s);printf("err %s line %d\n",s,__LINE__);return 1;}}while(0)
int main(){ CK(cuInit(0)); CUdevice d; CK(cuDeviceGet(&d,0)); CUcontext c; CK(cuCtxCreate(&c,0,d)); CUmodule m; CK(cuModuleLoad(&m,"ic_noconst.cubin"));
 CUdeviceptr out; CK(cuMemAlloc(&out, 148*8*256*8)); CUdeviceptr cc; size_t sz; CK(cuModuleGetGlobal(&cc,&sz,m,"CC")); std::vector<double> h(1024); for(int i=0;i<1024;i++)h[i]=1.0+1e-9*i; CK(cuMemcpyHtoD(cc,h.data(),8192));
 int sizes[]={64,128,256,384,512,768,1024,2048,4096,8192};
 CUevent e0,e1; cuEventCreate(&e0,0); cuEventCreate(&e1,0);
 for(int ws=1; ws<=2; ++ws) for(int N:sizes){ char nm[32]; sprintf(nm,"body%d",N); CUfunction f; CK(cuModuleGetFunction(&f,m,nm)); unsigned iters=(1<<22)/N; void*args[]={&out,&iters};
  int threads = ws==1?256:128;  // 8 warps or 4 warps per SM
  for(int rep=0;rep<2;++rep){ cuEventRecord(e0,0); CK(cuLaunchKernel(f,148,1,1,threads,1,1,0,0,args,0)); cuEventRecord(e1,0); CK(cuEventSynchronize(e1)); float ms; cuEventElapsedTime(&ms,e0,e1);
   if(rep) printf("body %5d DFMA (%6d B code) threads/SM=%d: %.3f ms  %.2f DFMA/clk/SM(@1965)  %.2f TFLOP/s\n",N,N*24,threads,ms,(double)iters*N*threads/ (ms*1e-3*1.965e9), 2.0*iters*N*threads*148/ms*1e-9);}}
 return 0; }
