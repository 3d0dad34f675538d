// H2D bandwidth probe (run under gpurun): pinned staging buffers written by host threads right before the copy.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  const size_t sizes[3] = {307200, 9830400, 19660800};
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  void *d;
  cudaMalloc(&d, sizes[2]);
  std::vector<unsigned char> src(sizes[2], 7);
  for (int kind = 0; kind < 3; kind++) {
    void *h = nullptr;
    if (kind == 0) cudaMallocHost(&h, sizes[2]);
    if (kind == 1) cudaHostAlloc(&h, sizes[2], cudaHostAllocWriteCombined);
    if (kind == 2) { h = aligned_alloc(4096, sizes[2]); cudaHostRegister(h, sizes[2], cudaHostRegisterDefault); }
    for (size_t bytes : sizes)
      for (int writers = 0; writers <= 8; writers += 4) {
        double best = 1e9, stage = 0;
        for (int rep = 0; rep < 6; rep++) {
          const double t0 = now();
          if (writers) {
            std::vector<std::thread> th;
            for (int t = 0; t < writers; t++)
              th.emplace_back([&, t]() { const size_t c = bytes / writers; std::memcpy((char *)h + c * t, src.data() + c * t, c); });
            for (auto &x : th) x.join();
          }
          const double t1 = now();
          cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
          cudaStreamSynchronize(st);
          const double t2 = now();
          if (t2 - t1 < best) { best = t2 - t1; stage = t1 - t0; }
        }
        printf("kind %d (0 cudaMallocHost, 1 write-combined, 2 registered) bytes %zu writers %d: stage %.1f us, h2d %.1f us = %.1f GB/s\n", kind, bytes, writers,
               1e6 * stage, 1e6 * best, bytes / best / 1e9);
      }
  }
  return 0;
}
