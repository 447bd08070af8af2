// Where do the CTAs of 8-CTA clusters land when two CTAs fit per SM?  Prints, per SM, the cluster ids resident on it.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out, int spin) {
  unsigned smid, cid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
  extern __shared__ float sm[];
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = smid; out[2 * blockIdx.x + 1] = cid; }
  // stay resident so that the whole grid is co-resident when sampled
  long long t0 = clock64();
  while (clock64() - t0 < spin) { sm[threadIdx.x] += 1.f; }
}
int main() {
  const int P = 8, smem = 115136, threads = 256;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = P; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
  cfg.gridDim = dim3(P * 148);
  int ncl = 0;
  cudaOccupancyMaxActiveClusters(&ncl, (const void*)k, &cfg);
  printf("max active clusters: %d\n", ncl);
  cfg.gridDim = dim3(P * ncl);
  int* out; cudaMalloc(&out, 2 * P * ncl * sizeof(int));
  int spin = 2000000;
  void* args[] = {&out, &spin};
  cudaLaunchKernelExC(&cfg, (const void*)k, args);
  cudaDeviceSynchronize();
  int h[2 * 8 * 64];
  cudaMemcpy(h, out, 2 * P * ncl * sizeof(int), cudaMemcpyDeviceToHost);
  int sm_cl[160][4], cnt[160] = {0};
  for (int b = 0; b < P * ncl; ++b) { int s = h[2 * b]; if (cnt[s] < 4) sm_cl[s][cnt[s]++] = h[2 * b + 1]; }
  int same = 0, diff = 0, single = 0, empty = 0;
  for (int s = 0; s < 148; ++s) {
    if (cnt[s] == 0) ++empty; else if (cnt[s] == 1) ++single; else if (sm_cl[s][0] == sm_cl[s][1]) ++same; else ++diff;
  }
  printf("SMs with 2 CTAs of the SAME cluster: %d, of DIFFERENT clusters: %d, with one CTA: %d, empty: %d (%s)\n", same, diff, single, empty, cudaGetErrorString(cudaGetLastError()));
  for (int c = 0; c < 3; ++c) { printf("cluster %d on SMs:", c); for (int b = 0; b < P * ncl; ++b) if (h[2*b+1] == c) printf(" %d", h[2*b]); printf("\n"); }
  return 0;
}
