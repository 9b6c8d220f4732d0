// feed_probe.cu -- how fast can PAGEABLE host memory reach the GPU on this box?  (the staging path of aleo_b200_msm_g1 /
// aleo_b200_ntt_fr: helper threads memcpy slices into pinned double buffers and DMA from there.)  Prints GB/s for
// (a) the plain memcpy pageable -> pinned alone, (b) the staged copy, per thread count and slice size, (c) one
// cudaMemcpyAsync from pinned memory, (d) cudaHostRegister + copy + unregister.  Development tool, not product.
//   nvcc -O2 -o build/feed_probe tools/feed_probe.cu -lpthread && build/feed_probe
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  const size_t bytes = (size_t)2 << 30;
  unsigned char* src = (unsigned char*)malloc(bytes);
  memset(src, 1, bytes);
  unsigned char* dev;
  cudaMalloc(&dev, bytes);
  printf("host threads available: %u\n", std::thread::hardware_concurrency());
  // (c) pinned reference
  {
    unsigned char* pin;
    cudaMallocHost(&pin, bytes);
    memcpy(pin, src, bytes);
    cudaMemcpy(dev, pin, bytes, cudaMemcpyHostToDevice);
    double t = now();
    cudaMemcpy(dev, pin, bytes, cudaMemcpyHostToDevice);
    printf("pinned cudaMemcpy                          : %6.1f GB/s\n", bytes / (now() - t) / 1e9);
    cudaFreeHost(pin);
  }
  {
    double t = now();
    cudaMemcpy(dev, src, bytes, cudaMemcpyHostToDevice);
    printf("pageable cudaMemcpy (driver bounce)        : %6.1f GB/s\n", bytes / (now() - t) / 1e9);
  }
  {
    double t = now();
    cudaHostRegister(src, bytes, cudaHostRegisterDefault);
    double t1 = now();
    cudaMemcpy(dev, src, bytes, cudaMemcpyHostToDevice);
    double t2 = now();
    cudaHostUnregister(src);
    double t3 = now();
    printf("cudaHostRegister %.1f ms + copy %.1f ms + unregister %.1f ms: %6.1f GB/s overall\n", (t1 - t) * 1e3, (t2 - t1) * 1e3,
           (t3 - t2) * 1e3, bytes / (t3 - t) / 1e9);
  }
  for (size_t slice : {(size_t)1 << 20, (size_t)4 << 20, (size_t)16 << 20}) {
    for (int nt : {2, 4, 6, 8, 12, 16}) {
      std::vector<unsigned char*> buf(nt * 2);
      std::vector<cudaEvent_t> ev(nt * 2);
      std::vector<cudaStream_t> st(nt);
      for (int t = 0; t < nt; t++) {
        cudaStreamCreateWithFlags(&st[t], cudaStreamNonBlocking);
        for (int b = 0; b < 2; b++) {
          cudaMallocHost(&buf[t * 2 + b], slice);
          cudaEventCreateWithFlags(&ev[t * 2 + b], cudaEventDisableTiming);
        }
      }
      const size_t nslices = bytes / slice;
      for (int mode = 0; mode < 2; mode++) {  // 0: memcpy only, 1: staged copy
        double t0 = now();
        std::vector<std::thread> w;
        for (int t = 0; t < nt; t++)
          w.emplace_back([&, t]() {
            for (size_t i = t, round = 0; i < nslices; i += nt, round++) {
              const int b = round & 1;
              if (mode) cudaEventSynchronize(ev[t * 2 + b]);
              memcpy(buf[t * 2 + b], src + i * slice, slice);
              if (mode) {
                cudaMemcpyAsync(dev + i * slice, buf[t * 2 + b], slice, cudaMemcpyHostToDevice, st[t]);
                cudaEventRecord(ev[t * 2 + b], st[t]);
              }
            }
          });
        for (auto& x : w) x.join();
        cudaDeviceSynchronize();
        printf("slice %2zu MB, %2d threads, %-12s: %6.1f GB/s\n", slice >> 20, nt, mode ? "staged copy" : "memcpy only", bytes / (now() - t0) / 1e9);
      }
      for (int t = 0; t < nt; t++) {
        cudaStreamDestroy(st[t]);
        for (int b = 0; b < 2; b++) {
          cudaFreeHost(buf[t * 2 + b]);
          cudaEventDestroy(ev[t * 2 + b]);
        }
      }
    }
  }
  return 0;
}
