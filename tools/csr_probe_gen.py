"""Generates tools/bin/csr_probe.cu: csr_build_kernel (csrc/gather.cu) with clock64 stamps after every phase, plus a main() that runs
it on B=35, n=1024, mk=10240 random indices.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/csr_probe
tools/bin/csr_probe.cu.  tools/ only."""
import re
s=open('pdgn_b200/csrc/gather.cu').read()
a=s.index('constexpr int CSR_T = 1024;')
b=s.index('// Shared skeleton of the three pull kernels.')
body=s[a:b]
# instrument
body=body.replace('int* __restrict__ pos, int stage_pos) {','int* __restrict__ pos, int stage_pos, long long* __restrict__ stamps) {\n    long long t0 = clock64(); int ph = 0;\n#define STAMP() do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 0) stamps[ph++] = clock64() - t0; } while (0)')
body=body.replace('    __syncthreads();\n    // slice of warp w','    STAMP();\n    // slice of warp w')
body=body.replace('    __syncthreads();\n    // column prefix','    STAMP();\n    // column prefix')
body=body.replace('    __syncthreads();\n    // exclusive scan of the totals','    STAMP();\n    // exclusive scan of the totals')
body=body.replace('    if (t == CSR_T - 1) ob[n] = mk;\n    __syncthreads();','    if (t == CSR_T - 1) ob[n] = mk;\n    STAMP();')
body=body.rstrip()
assert body.endswith('}')
body=body[:-1]+'    STAMP();\n}\n'
body=body.replace('    if (stage_pos) {\n        __syncthreads();','    STAMP();\n    if (stage_pos) {\n        __syncthreads();')
src='''#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
namespace pdgn { constexpr unsigned kFull = 0xffffffffu;
'''+body+'''}
using namespace pdgn;
int main(){ int b=35,n=1024,mk=10240; int *idx,*offs,*pos; long long* st; cudaMalloc(&idx,b*mk*4); cudaMalloc(&offs,b*(n+1)*4); cudaMalloc(&pos,b*mk*4); cudaMalloc(&st,64*8);
 int* h=(int*)malloc(b*mk*4); for(int i=0;i<b*mk;i++) h[i]=rand()%n; cudaMemcpy(idx,h,b*mk*4,cudaMemcpyHostToDevice);
 size_t sm=csr_smem_total(n,mk); cudaFuncSetAttribute(csr_build_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sm);
 for(int it=0;it<3;it++){ cudaMemset(st,0,64*8); cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0); csr_build_kernel<int><<<b,CSR_T,sm>>>(idx,n,mk,offs,pos,1,st); cudaEventRecord(e1); cudaDeviceSynchronize(); float ms; cudaEventElapsedTime(&ms,e0,e1);
 long long hs[8]; cudaMemcpy(hs,st,64,cudaMemcpyDeviceToHost); printf("event %.1f us | cycles after: zero %lld count %lld colprefix %lld scan %lld fill %lld writeout %lld  (%s)\\n", ms*1e3, hs[0],hs[1],hs[2],hs[3],hs[4],hs[5], cudaGetErrorString(cudaGetLastError())); }
 return 0; }
'''
open('tools/bin/csr_probe.cu','w').write(src)
