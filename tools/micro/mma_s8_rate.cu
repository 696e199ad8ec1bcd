// Microbenchmark: issue rate of mma.sync.m16n8k32 u8 x s8 -> s32 on sm_100a (legacy tensor
// path), to decide whether a warp-private Toeplitz GEMM can replace the IDP.2A first-stage FIR.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(int iters, int *out, long long *cycles) {
  int c[8][4] = {};
  uint32_t a[4] = {threadIdx.x * 2654435761u, 0x01020304u, 0x05060708u, threadIdx.x};
  uint32_t b[2] = {0x01010101u, 0x02020202u * threadIdx.x};
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  int *out; long long *cyc;
  cudaMalloc(&out, 148 * 8 * 1024 * 4); cudaMalloc(&cyc, 148 * 8 * 8);
  for (int warps = 1; warps <= 16; warps *= 2) {
    const int iters = 4096;
    k<<<148, 32 * warps>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<148, 32 * warps>>>(iters, out, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double mmas = (double)iters * 8 * warps;           // per SM
    printf("warps/SM %2d: %.1f cycles per MMA per SM (%.2f MMA/clk/SM), %.1f TMAC/s chip, err %s\n", warps,
           (double)h / mmas, mmas / (double)h, mmas * 148 * 16 * 8 * 32 / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
