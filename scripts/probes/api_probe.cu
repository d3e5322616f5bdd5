// GPU-box probe: does the CUDA runtime serialise API calls of concurrent host threads?  T threads, each with its own stream, issue
// the call mix of one frame-batch (small H2D copy, a few tiny kernels, small D2H copy, event record) back to back; prints calls/s per
// thread count, with and without other threads spinning on cudaEventQuery (what Batch::wait does while it has nothing to help with).
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

__global__ void tiny(int* p) { if (threadIdx.x == 0 && p) p[blockIdx.x] += 1; }
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  cudaFree(0);
  for (int spin = 0; spin < 2; ++spin)
    for (int T : {1, 2, 4, 8, 12}) {
      std::atomic<bool> go{false}, stop{false};
      std::atomic<long> calls{0};
      std::vector<std::thread> th;
      const int iters = 400;
      std::vector<double> per_thread_us(T, 0);
      for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] {
          cudaStream_t st;
          cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
          cudaEvent_t ev;
          cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
          int *d = nullptr, *h = nullptr;
          cudaMalloc((void**)&d, 4096);
          cudaMallocHost((void**)&h, 4096);
          while (!go.load()) {}
          const double t0 = now_us();
          for (int i = 0; i < iters; ++i) {
            cudaMemcpyAsync(d, h, 2048, cudaMemcpyHostToDevice, st);
            for (int k = 0; k < 6; ++k) tiny<<<8, 64, 0, st>>>(d);
            cudaMemcpyAsync(h, d, 2048, cudaMemcpyDeviceToHost, st);
            cudaEventRecord(ev, st);
            calls += 9;
            if ((i & 7) == 7) cudaEventSynchronize(ev);  // keep the launch queue from filling up
          }
          per_thread_us[t] = now_us() - t0;
          cudaStreamSynchronize(st);
        });
      std::vector<std::thread> spinners;
      if (spin)
        for (int s = 0; s < 4; ++s)
          spinners.emplace_back([&] {
            cudaEvent_t ev;
            cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            cudaStream_t st;
            cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
            cudaEventRecord(ev, st);
            while (!stop.load()) cudaEventQuery(ev);
          });
      go = true;
      for (auto& t : th) t.join();
      stop = true;
      for (auto& s : spinners) s.join();
      double mx = 0;
      for (double u : per_thread_us) mx = u > mx ? u : mx;
      printf("%2d issuing threads%s: %.2f us per call per thread, %.2f M calls/s in total\n", T, spin ? " + 4 threads spinning on cudaEventQuery" : "", mx / (iters * 9.0),
             T * iters * 9.0 / mx);
    }
  return 0;
}
