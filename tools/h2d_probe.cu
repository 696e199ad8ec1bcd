// h2d_probe -- what the host side of one box can feed its GPUs: pinned host -> device copies alone,
// on 1, 2, 4, ... GPUs at once (one stream per GPU, one host thread). Separates the PCIe / host-DRAM
// ceiling from anything the demodulation engine does. Prints one JSON line per GPU count.
//   nvcc -O2 -o h2d_probe tools/h2d_probe.cu ; ./h2d_probe [MiB per copy] [copies]
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char **argv) {
  const size_t mib = argc > 1 ? (size_t)atol(argv[1]) : 512;
  const int copies = argc > 2 ? atoi(argv[2]) : 20;
  int n_dev = 0;
  CK(cudaGetDeviceCount(&n_dev));
  const size_t bytes = mib << 20;
  std::vector<void *> h(n_dev), d(n_dev), hb(n_dev), db(n_dev);
  std::vector<cudaStream_t> st(n_dev), st2(n_dev);
  for (int g = 0; g < n_dev; ++g) {
    CK(cudaSetDevice(g));
    CK(cudaHostAlloc(&h[g], bytes, cudaHostAllocPortable));
    CK(cudaHostAlloc(&hb[g], bytes / 32, cudaHostAllocPortable));
    memset(h[g], g + 1, bytes);
    CK(cudaMalloc(&d[g], bytes));
    CK(cudaMalloc(&db[g], bytes / 32));
    CK(cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&st2[g], cudaStreamNonBlocking));
  }
  for (int n = 1; n <= n_dev; n *= 2) {
    for (int pass = 0; pass < 2; ++pass) {  // 0: H2D alone; 1: H2D with the PCM-sized D2H beside it
      for (int g = 0; g < n; ++g) { CK(cudaSetDevice(g)); CK(cudaMemcpyAsync(d[g], h[g], bytes, cudaMemcpyHostToDevice, st[g])); }
      for (int g = 0; g < n; ++g) { CK(cudaSetDevice(g)); CK(cudaStreamSynchronize(st[g])); }
      const auto t0 = std::chrono::steady_clock::now();
      for (int k = 0; k < copies; ++k)
        for (int g = 0; g < n; ++g) {
          CK(cudaSetDevice(g));
          CK(cudaMemcpyAsync(d[g], h[g], bytes, cudaMemcpyHostToDevice, st[g]));
          if (pass) CK(cudaMemcpyAsync(hb[g], db[g], bytes / 32, cudaMemcpyDeviceToHost, st2[g]));
        }
      for (int g = 0; g < n; ++g) { CK(cudaSetDevice(g)); CK(cudaStreamSynchronize(st[g])); CK(cudaStreamSynchronize(st2[g])); }
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      const double gbs = (double)bytes * copies * n / s / 1e9;
      printf("{\"tool\": \"h2d_probe\", \"gpus\": %d, \"with_d2h\": %d, \"mib_per_copy\": %zu, \"copies\": %d, \"aggregate_h2d_gb_per_s\": %.2f, "
             "\"per_gpu_gb_per_s\": %.2f, \"iq_msamples_per_s_ceiling\": %.1f}\n",
             n, pass, mib, copies, gbs, gbs / n, gbs * 1e9 / 2.0 / 1e6);
    }
  }
  return 0;
}
