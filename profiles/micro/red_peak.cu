// RED.F64 throughput microbenchmark: the bound of k_score.  Every thread issues N
// fire-and-forget red.global.add.f64 to pseudo-random addresses in a footprint of
// `n_addr` doubles (k_score's mesh tally on config C2: 30 000 doubles), in runs of
// `run` adjacent bins x `per_bin` consecutive doubles (the [bin][score] layout).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_peak red_peak.cu && ./red_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_red(double* t, int64_t n_addr, int per_thread, int per_bin, int f32)
{
  uint64_t s = (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 6364136223846793005ULL + 1442695040888963407ULL;
  for (int j = 0; j < per_thread; j += per_bin) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    const int64_t bin = (int64_t)((s >> 33) % (uint64_t)(n_addr / per_bin));
    for (int k = 0; k < per_bin; ++k) {
      double* p = t + bin * per_bin + k;
      if (f32)
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f));
      else
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(1.0));
    }
  }
}

int main()
{
  const int64_t max_addr = 1 << 26;
  double* t;
  cudaMalloc(&t, max_addr * sizeof(double));
  cudaMemset(t, 0, max_addr * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int per_thread = 60;
  const int threads = 1 << 20;
  printf("%10s %8s %6s %12s %10s\n", "n_addr", "per_bin", "type", "REDs/s", "us");
  for (int f32 = 0; f32 < 2; ++f32)
    for (int per_bin : {1, 3})
      for (int64_t n_addr : {(int64_t)30000, (int64_t)300000, (int64_t)3000000, (int64_t)60000000}) {
        float best = 1e30f;
        for (int it = 0; it < 5; ++it) {
          cudaEventRecord(e0);
          k_red<<<threads / 256, 256>>>(t, n_addr, per_thread, per_bin, f32);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (it > 0 && ms < best)
            best = ms;
        }
        printf("%10ld %8d %6s %12.4g %10.1f\n", (long)n_addr, per_bin, f32 ? "f32" : "f64",
          (double)threads * per_thread / (best * 1e-3), best * 1e3);
      }
  return 0;
}
