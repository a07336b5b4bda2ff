// Latency micro-benchmarks (one warp): dependent DADD, DMUL, 64-bit SHFL, DSETP+select, FP64 divide, ldexp, LDS, LDG(L2).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_dadd(double *o, double a){ double x = threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i=0;i<N;++i) x = __dadd_rn(x, a);
  long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("DADD dependent: %.1f cyc\n", (double)(t1-t0)/N); }
__global__ void k_dmul(double *o, double a){ double x = threadIdx.x+1; long long t0 = clock64();
#pragma unroll 16
  for (int i=0;i<N;++i) x = __dmul_rn(x, a);
  long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("DMUL dependent: %.1f cyc\n", (double)(t1-t0)/N); }
__global__ void k_shfl(double *o){ double x = threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i=0;i<N;++i) x = __shfl_sync(0xffffffffu, x, (i+1)&31);
  long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("SHFL64 dependent: %.1f cyc\n", (double)(t1-t0)/N); }
__global__ void k_shfl32(int *o){ int x = threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i=0;i<N;++i) x = __shfl_sync(0xffffffffu, x, (i+1)&31);
  long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("SHFL32 dependent: %.1f cyc\n", (double)(t1-t0)/N); }
__global__ void k_chain(double *o, const double *tri, double pen){ // the DP chain step
  int lane = threadIdx.x; double best = lane * 0.5; int arg = 0; long long t0 = clock64();
  for (int rep=0; rep<N/32; ++rep) {
#pragma unroll 8
  for (int k=0;k<32;++k){ double pf = __dadd_rn(best, pen); double pk = __shfl_sync(0xffffffffu, pf, k);
     if (lane > k) { double t = __dadd_rn(tri[k*32+lane], pk); if (t > best) { best = t; arg = k; } } }
  }
  long long t1 = clock64(); o[threadIdx.x] = best + arg; if (!threadIdx.x) printf("chain step: %.1f cyc\n", (double)(t1-t0)/N); }
// the chain step of exact_pruned.cu: triangle + the next step's rows (second accumulator), tiles [63][32] in shared memory
__global__ void k_chain2(double *o, const double *g, double pen){
  __shared__ double tri[63*32], nxt[63*32];
  for (int i = threadIdx.x; i < 63*32; i += 32) { tri[i] = g[i & 1023]; nxt[i] = g[(i*7) & 1023]; }
  __syncwarp();
  int lane = threadIdx.x; double best = lane * 0.5, best2 = -1e300; int arg = 0, arg2 = 0; long long t0 = clock64();
  for (int rep=0; rep<N/32; ++rep) {
#pragma unroll 8
  for (int k=0;k<32;++k){ double pf = __dadd_rn(best, pen); double pk = __shfl_sync(0xffffffffu, pf, k);
     if (lane > k) { double t = __dadd_rn(tri[(lane-k-1)*32+lane], pk); if (t > best) { best = t; arg = k; } }
     { double t2 = __dadd_rn(nxt[(31+lane-k)*32+lane], pk); if (t2 > best2) { best2 = t2; arg2 = k; } } }
  best = best2 * 0.5; best2 = -1e300;
  }
  long long t1 = clock64(); o[threadIdx.x] = best + arg + best2 + arg2; if (!threadIdx.x) printf("chain step + next-rows accumulator: %.1f cyc/row\n", (double)(t1-t0)/N); }
// micro-steps of M rows: the M current maxima are broadcast, every lane resolves the M-row chain redundantly
// (values only), then folds the M finished P into its own row with a tree of compares
template <int M>
__global__ void k_chain_micro(double *o, const double *g, double pen){
  __shared__ double tri[63*32], nxt[63*32];
  for (int i = threadIdx.x; i < 63*32; i += 32) { tri[i] = g[i & 1023]; nxt[i] = g[(i*7) & 1023]; }
  __syncwarp();
  int lane = threadIdx.x; double best = lane * 0.5, best2 = -1e300; int arg = 0, arg2 = 0; long long t0 = clock64();
  for (int rep=0; rep<N/32; ++rep) {
#pragma unroll 2
  for (int k=0;k<32;k+=M){
     double b[M], pq[M];
#pragma unroll
     for (int q=0;q<M;++q) b[q] = __shfl_sync(0xffffffffu, best, k+q);
#pragma unroll
     for (int q=0;q<M;++q) {            // row k+q against the rows k..k+q-1 of this micro-step (uniform addresses: broadcast)
        double m = b[q];
#pragma unroll
        for (int r=0;r<q;++r) { double t = __dadd_rn(tri[(q-r-1)*32 + k+q], pq[r]); m = t > m ? t : m; }
        pq[q] = __dadd_rn(m, pen);
     }
     // fold the M finished columns into this lane's row (first maximum: lower column wins ties)
     double m = -INFINITY; int a = 0;
#pragma unroll
     for (int q=0;q<M;++q) { if (lane > k+q) { double t = __dadd_rn(tri[(lane-k-q-1)*32+lane], pq[q]); if (t > m) { m = t; a = k+q; } } }
     if (m > best) { best = m; arg = a; }
     double m2 = -INFINITY; int a2 = 0;
#pragma unroll
     for (int q=0;q<M;++q) { double t = __dadd_rn(nxt[(31+lane-k-q)*32+lane], pq[q]); if (t > m2) { m2 = t; a2 = k+q; } }
     if (m2 > best2) { best2 = m2; arg2 = a2; }
  }
  best = best2 * 0.5; best2 = -1e300;
  }
  long long t1 = clock64(); o[threadIdx.x] = best + arg + best2 + arg2; if (!threadIdx.x) printf("chain micro-step M=%d (+ next-rows accumulator): %.1f cyc/row\n", M, (double)(t1-t0)/N); }
__global__ void k_div(double *o, double a){ double x = threadIdx.x+3; long long t0 = clock64();
#pragma unroll 4
  for (int i=0;i<N;++i) x = a / x + 1.5;
  long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("DDIV+DADD dependent: %.1f cyc\n", (double)(t1-t0)/N); }
__global__ void k_ldexp(double *o, int e){ double x = threadIdx.x+3; long long t0 = clock64();
#pragma unroll 4
  for (int i=0;i<N;++i) x = ldexp(x, e) + 1.0;
  long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) printf("ldexp+DADD dependent: %.1f cyc\n", (double)(t1-t0)/N); }
__global__ void k_setp(double *o, double a){ double x = threadIdx.x; double b = 0; long long t0 = clock64();
#pragma unroll 16
  for (int i=0;i<N;++i) { if (x > b) b = x + a; else x = b + a; }
  long long t1 = clock64(); o[threadIdx.x] = x + b; if (!threadIdx.x) printf("DSETP+sel+DADD: %.1f cyc\n", (double)(t1-t0)/N); }
__global__ void k_ldg(double *o, const int *next, int n){ int p = threadIdx.x; long long t0 = clock64();
  for (int i=0;i<1024;++i) p = __ldg(next + p);
  long long t1 = clock64(); o[threadIdx.x] = p; if (!threadIdx.x) printf("LDG pointer chase (%d MB footprint): %.1f cyc\n", n/262144, (double)(t1-t0)/1024); }
__global__ void k_lds(double *o){ __shared__ int s[1024]; for (int i=threadIdx.x;i<1024;i+=32) s[i]=(i*7+1)&1023; __syncwarp(); int p = threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i=0;i<N;++i) p = s[p];
  long long t1 = clock64(); o[threadIdx.x] = p; if (!threadIdx.x) printf("LDS dependent: %.1f cyc\n", (double)(t1-t0)/N); }
__global__ void k_bar(double *o){ long long t0 = clock64();
#pragma unroll 16
  for (int i=0;i<N;++i) __syncthreads();
  long long t1 = clock64(); if (!threadIdx.x) { o[0] = 1; printf("__syncthreads (256 thr): %.1f cyc\n", (double)(t1-t0)/N); } }
int main(){ double *o; cudaMalloc(&o, 4096*8); double *tri; cudaMalloc(&tri, 1024*8); cudaMemset(tri, 0, 1024*8);
  k_dadd<<<1,32>>>(o, 1e-9); k_dmul<<<1,32>>>(o, 1.0000001); k_shfl<<<1,32>>>(o); k_shfl32<<<1,32>>>((int*)o);
  k_chain<<<1,32>>>(o, tri, -0.1); k_chain2<<<1,32>>>(o, tri, -0.1); k_chain_micro<2><<<1,32>>>(o, tri, -0.1); k_chain_micro<4><<<1,32>>>(o, tri, -0.1); k_chain_micro<8><<<1,32>>>(o, tri, -0.1); k_div<<<1,32>>>(o, 3.0); k_ldexp<<<1,32>>>(o, -44); k_setp<<<1,32>>>(o, 0.25); k_lds<<<1,32>>>(o); k_bar<<<1,256>>>(o);
  for (int mb : {1, 16, 64}) { int n = mb*262144; int *h = new int[n]; for (int i=0;i<n;++i) h[i] = (int)(((long long)i*40503 + 12345) % n);
    int *d; cudaMalloc(&d, n*4); cudaMemcpy(d, h, n*4, cudaMemcpyHostToDevice); k_ldg<<<1,32>>>(o, d, n); k_ldg<<<1,32>>>(o, d, n); cudaDeviceSynchronize(); cudaFree(d); delete[] h; }
  cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError())); return 0; }
