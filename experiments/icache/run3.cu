#include <cuda.h>
#include <cstdio>
#include <vector>
#define CK(x) do{CUresult r=(x); if(r){const char*s;cuGetErrorString(r,&s);printf("err %s line %d\n",s,__LINE__);return 1;}}while(0)
int main(){ CK(cuInit(0)); CUdevice d; CK(cuDeviceGet(&d,0)); CUcontext c; CK(cuCtxCreate(&c,0,d)); CUmodule m; CK(cuModuleLoad(&m,"ic3.cubin"));
 CUdeviceptr out; CK(cuMemAlloc(&out, 148*8*256*8)); CUdeviceptr cc; size_t sz; CK(cuModuleGetGlobal(&cc,&sz,m,"CC")); std::vector<double> h(7168); for(int i=0;i<7168;i++)h[i]=1.0+1e-9*i; CK(cuMemcpyHtoD(cc,h.data(),7168*8));
 int sizes[]={64,128,256}; const char* modes[]={"c64","c128"};
 CUevent e0,e1; cuEventCreate(&e0,0); cuEventCreate(&e1,0);
 for(const char* mo:modes) for(int threads: {256,128}) for(int N:sizes){ char nm[32]; sprintf(nm,"%s_body%d",mo,N); CUfunction f; CK(cuModuleGetFunction(&f,m,nm)); unsigned iters=(1<<22)/N; void*args[]={&out,&iters};
  for(int rep=0;rep<2;++rep){ cuEventRecord(e0,0); CK(cuLaunchKernel(f,148,1,1,threads,1,1,0,0,args,0)); cuEventRecord(e1,0); CK(cuEventSynchronize(e1)); float ms; cuEventElapsedTime(&ms,e0,e1);
   if(rep) printf("%-5s body %5d DFMA threads/SM=%d: %.3f ms  %.2f DFMA/clk/SM(@1965)  %.2f TFLOP/s\n",mo,N,threads,ms,(double)iters*N*threads/ (ms*1e-3*1.965e9), 2.0*iters*N*threads*148/ms*1e-9);}}
 return 0; }
