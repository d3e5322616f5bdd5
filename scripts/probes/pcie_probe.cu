// GPU-box probe: what a small host->device table upload costs while the frame uploads of other batches occupy the H2D path, and
// how fast frames can be pulled.  Answers (printed):
//   a. bandwidth of N x cudaMemcpyAsync(307 KB) issued from one C++ thread (no Python in the loop)
//   b. latency of a 4 KB H2D copy + stream sync: idle, and while another thread streams 20 MB bursts of frame copies on another stream
//   c. the same small table fetched by a kernel that reads mapped pinned memory (no copy engine involved)
//   d. bandwidth of an SM gather of 64 pinned frames (zero-copy ingest) as a function of CTAs per frame
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__global__ void fetch_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n16) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void gather_kernel(const uint8_t* const* __restrict__ src, uint8_t* __restrict__ dst, size_t bytes) {
  const uint4* s = reinterpret_cast<const uint4*>(src[blockIdx.y]);
  uint4* d = reinterpret_cast<uint4*>(dst + (size_t)blockIdx.y * bytes);
  const size_t n16 = bytes / 16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}

static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  const size_t FB = 640 * 480;
  const int S = 512, NF = 4;
  uint8_t* host = nullptr;
  uint8_t* dev = nullptr;
  CK(cudaHostAlloc((void**)&host, (size_t)S * NF * FB, cudaHostAllocMapped));
  memset(host, 7, (size_t)S * NF * FB);
  CK(cudaMalloc((void**)&dev, (size_t)S * FB));
  cudaStream_t st_big, st_small;
  CK(cudaStreamCreateWithFlags(&st_big, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&st_small, cudaStreamNonBlocking));
  // a. per-frame copies from one thread
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaStreamSynchronize(st_big));
    const double t0 = now_us();
    for (int it = 0; it < 4; ++it)
      for (int s = 0; s < S; ++s) CK(cudaMemcpyAsync(dev + (size_t)s * FB, host + ((size_t)s * NF + it) * FB, FB, cudaMemcpyHostToDevice, st_big));
    const double t_issue = now_us() - t0;
    CK(cudaStreamSynchronize(st_big));
    const double dt = now_us() - t0;
    if (rep) printf("a. 4 x 512 x cudaMemcpyAsync(307 KB), one thread: %.1f GB/s (issue %.2f us per call)\n", 4.0 * S * FB / dt / 1e3, t_issue / (4.0 * S));
  }
  // b / c. small-table latency, idle and under frame-copy load
  uint8_t* tab_h = nullptr;
  uint8_t* tab_d = nullptr;
  uint8_t* tab_hd = nullptr;
  CK(cudaHostAlloc((void**)&tab_h, 4096, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer((void**)&tab_hd, tab_h, 0));
  CK(cudaMalloc((void**)&tab_d, 4096));
  for (int load = 0; load < 3; ++load) {  // 0 idle, 1 per-frame copies in bursts of 64, 2 the same bursts in chunks of 8 frames with a yield between chunks
    std::atomic<bool> stop{false};
    std::thread bg;
    if (load) bg = std::thread([&, load] {
      int it = 0;
      while (!stop.load()) {
        for (int s = 0; s < 64; ++s) {
          cudaMemcpyAsync(dev + (size_t)s * FB, host + ((size_t)s * NF + (it & 3)) * FB, FB, cudaMemcpyHostToDevice, st_big);
          if (load == 2 && (s & 7) == 7) cudaStreamSynchronize(st_big);
        }
        cudaStreamSynchronize(st_big);
        ++it;
      }
    });
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
    for (int mode = 0; mode < 2; ++mode) {
      std::vector<double> lat;
      for (int i = 0; i < 400; ++i) {
        const double t0 = now_us();
        if (mode == 0) cudaMemcpyAsync(tab_d, tab_h, 4096, cudaMemcpyHostToDevice, st_small);
        else fetch_kernel<<<1, 256, 0, st_small>>>(reinterpret_cast<const uint4*>(tab_hd), reinterpret_cast<uint4*>(tab_d), 256);
        cudaStreamSynchronize(st_small);
        lat.push_back(now_us() - t0);
        std::this_thread::sleep_for(std::chrono::microseconds(50));
      }
      std::sort(lat.begin(), lat.end());
      printf("%s 4 KB table %s, %s: median %.1f us, p90 %.1f us, p99 %.1f us\n", mode ? "c." : "b.", mode ? "read in place by a kernel (mapped pinned)" : "cudaMemcpyAsync",
             load == 0 ? "idle" : load == 1 ? "under 64-frame copy bursts" : "under bursts chunked by 8 frames", lat[200], lat[360], lat[396]);
    }
    if (load) { stop = true; bg.join(); }
  }
  // d. zero-copy gather bandwidth vs CTAs per frame
  const uint8_t** ptr_h = nullptr;
  const uint8_t** ptr_d = nullptr;
  CK(cudaMallocHost((void**)&ptr_h, sizeof(void*) * 64));
  CK(cudaMalloc((void**)&ptr_d, sizeof(void*) * 64));
  uint8_t* host_d = nullptr;
  CK(cudaHostGetDevicePointer((void**)&host_d, host, 0));
  for (int s = 0; s < 64; ++s) ptr_h[s] = host_d + ((size_t)s * NF + 1) * FB;
  CK(cudaMemcpy(ptr_d, ptr_h, sizeof(void*) * 64, cudaMemcpyHostToDevice));
  for (int ctas : {8, 16, 32, 64, 128, 296}) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_kernel<<<dim3(ctas, 64), 256, 0, st_big>>>(ptr_d, dev, FB);
    CK(cudaStreamSynchronize(st_big));
    cudaEventRecord(e0, st_big);
    for (int i = 0; i < 4; ++i) gather_kernel<<<dim3(ctas, 64), 256, 0, st_big>>>(ptr_d, dev, FB);
    cudaEventRecord(e1, st_big);
    CK(cudaStreamSynchronize(st_big));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("d. SM gather of 64 pinned frames, %3d CTAs per frame: %.1f GB/s\n", ctas, 4.0 * 64 * FB / ms / 1e6);
  }
  return 0;
}
