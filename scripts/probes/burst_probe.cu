// GPU-box probe: how long a small table round trip (165 KB H2D + tiny kernel + 100 KB D2H + ticket) takes when it is issued right after the
// frame burst of a whole step (8 batches x 64 frames of 307 KB = 157 MB) has been queued, for the ways the burst can travel:
//   dma: 32 pitched cudaMemcpy2DAsync of 16 frames on 8 streams;  sm: 8 gather kernels (2 CTAs x 128 threads per frame, 4 loads in flight);
//   sm_slim: the same gathers with 1 CTA x 64 threads per frame and one load in flight;  none: idle link.
// The small transfer is done either by the copy engine (cudaMemcpyAsync) or by a kernel that reads the mapped table in place.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

__global__ void gather(const uint8_t* const* __restrict__ src, uint8_t* __restrict__ dst, size_t bytes, int unroll4) {
  const uint4* s = reinterpret_cast<const uint4*>(src[blockIdx.y]);
  uint4* d = reinterpret_cast<uint4*>(dst + (size_t)blockIdx.y * bytes);
  const size_t n = bytes / 16, step = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (unroll4)
    for (; i + 3 * step < n; i += 4 * step) {
      const uint4 a = s[i], b = s[i + step], c = s[i + 2 * step], e = s[i + 3 * step];
      d[i] = a; d[i + step] = b; d[i + 2 * step] = c; d[i + 3 * step] = e;
    }
  for (; i < n; i += step) d[i] = s[i];
}
__global__ void consume(const uint4* __restrict__ tab, uint4* __restrict__ out, int n16) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) out[i] = tab[i];
}
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  const size_t FB = 640 * 480;
  const int NB = 8, S = 64, NF = 2;
  uint8_t *host, *dev, *host_d;
  cudaHostAlloc((void**)&host, (size_t)NB * S * NF * FB, cudaHostAllocMapped);
  memset(host, 3, (size_t)NB * S * NF * FB);
  cudaHostGetDevicePointer((void**)&host_d, host, 0);
  cudaMalloc((void**)&dev, (size_t)NB * S * FB);
  cudaStream_t sb[NB], ss;
  for (auto& s : sb) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ss, cudaStreamNonBlocking);
  const uint8_t** ptr_h;
  const uint8_t** ptr_d;
  cudaMallocHost((void**)&ptr_h, sizeof(void*) * NB * S);
  cudaMalloc((void**)&ptr_d, sizeof(void*) * NB * S);
  for (int i = 0; i < NB * S; ++i) ptr_h[i] = host_d + ((size_t)i * NF + 1) * FB;
  cudaMemcpy(ptr_d, ptr_h, sizeof(void*) * NB * S, cudaMemcpyHostToDevice);
  uint8_t *tab_h, *tab_hd, *tab_d, *out_d, *out_h;
  unsigned *flag_h;
  const int TB = 165 * 1024, OB = 100 * 1024;
  cudaHostAlloc((void**)&tab_h, TB, cudaHostAllocMapped);
  cudaHostGetDevicePointer((void**)&tab_hd, tab_h, 0);
  cudaMalloc((void**)&tab_d, TB);
  cudaMalloc((void**)&out_d, TB);
  cudaMallocHost((void**)&out_h, OB);
  cudaMallocHost((void**)&flag_h, 64);
  const char* burst_name[] = {"none", "dma (32 pitched copies, 8 streams)", "sm gather 2x128 thr/frame, 4 loads in flight", "sm gather 1x64 thr/frame, 1 load in flight"};
  for (int burst = 0; burst < 4; ++burst)
    for (int zc = 0; zc < 2; ++zc) {
      std::vector<double> lat, tot;
      for (int rep = 0; rep < 6; ++rep) {
        cudaDeviceSynchronize();
        const double t0 = now_us();
        for (int b = 0; b < NB && burst; ++b) {
          if (burst == 1)
            for (int q = 0; q < 4; ++q)
              cudaMemcpy2DAsync(dev + ((size_t)b * S + q * 16) * FB, FB, host + (((size_t)b * S + q * 16) * NF + (rep & 1)) * FB, NF * FB, FB, 16, cudaMemcpyHostToDevice, sb[b]);
          else if (burst == 2) gather<<<dim3(2, S), 128, 0, sb[b]>>>(ptr_d + b * S, dev + (size_t)b * S * FB, FB, 1);
          else gather<<<dim3(1, S), 64, 0, sb[b]>>>(ptr_d + b * S, dev + (size_t)b * S * FB, FB, 0);
        }
        const double t1 = now_us();
        if (!zc) {
          cudaMemcpyAsync(tab_d, tab_h, TB, cudaMemcpyHostToDevice, ss);
          consume<<<8, 256, 0, ss>>>(reinterpret_cast<const uint4*>(tab_d), reinterpret_cast<uint4*>(out_d), TB / 16);
        } else {
          consume<<<8, 256, 0, ss>>>(reinterpret_cast<const uint4*>(tab_hd), reinterpret_cast<uint4*>(out_d), TB / 16);
        }
        cudaMemcpyAsync(out_h, out_d, OB, cudaMemcpyDeviceToHost, ss);
        cudaStreamSynchronize(ss);
        const double t2 = now_us();
        cudaDeviceSynchronize();
        const double t3 = now_us();
        if (rep) { lat.push_back(t2 - t1); tot.push_back(t3 - t0); }
      }
      double l = 0, t = 0;
      for (double x : lat) l += x;
      for (double x : tot) t += x;
      printf("burst: %-48s table by %-22s round trip %8.1f us   burst done after %8.1f us\n", burst_name[burst], zc ? "in-place kernel read" : "copy engine", l / lat.size(), t / tot.size());
    }
  return 0;
}
