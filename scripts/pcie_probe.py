"""H2D bandwidth of the frame-ingest shapes (GPU box): what bounds the e2e pass of bench.py.
  a. one contiguous 157 MB copy (512 frames of 640x480), b. 512 separate 307 KB cudaMemcpyAsync from a strided pinned buffer,
  c. one cudaMemcpy2DAsync (512 rows of 307 KB, source pitch = frames_per_stream * 307 KB), d. the same in 8 chunks of 64 rows on 2 streams."""
import time
import torch
from cuda import cudart

FB, S, NF = 640 * 480, 512, 8
host = torch.empty((S, NF, FB), dtype=torch.uint8).pin_memory()
host.random_(0, 255)
dev = torch.empty((S, FB), dtype=torch.uint8, device="cuda")
contig = torch.empty((S, FB), dtype=torch.uint8).pin_memory()
st = torch.cuda.Stream()
st2 = torch.cuda.Stream()


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    return S * FB / dt / 1e9, dt * 1e3


def a():
    with torch.cuda.stream(st):
        dev.copy_(contig, non_blocking=True)


def b(f=[0]):
    f[0] = (f[0] + 1) % NF
    with torch.cuda.stream(st):
        for s in range(S):
            dev[s].copy_(host[s, f[0]], non_blocking=True)


K = cudart.cudaMemcpyKind.cudaMemcpyHostToDevice


def c(f=[0]):
    f[0] = (f[0] + 1) % NF
    cudart.cudaMemcpy2DAsync(dev.data_ptr(), FB, host.data_ptr() + f[0] * FB, NF * FB, FB, S, K, st.cuda_stream)


def d(f=[0]):
    f[0] = (f[0] + 1) % NF
    for i in range(8):
        cudart.cudaMemcpy2DAsync(dev.data_ptr() + i * 64 * FB, FB, host.data_ptr() + (i * 64 * NF + f[0]) * FB, NF * FB, FB, 64, K, (st if i % 2 else st2).cuda_stream)


for name, fn in (("contiguous 157MB", a), ("512 x memcpyAsync 307KB", b), ("one memcpy2D 512 rows", c), ("8 x memcpy2D 64 rows, 2 streams", d)):
    gbs, ms = timeit(fn)
    print(f"{name:36s} {gbs:7.2f} GB/s  {ms:7.3f} ms per 512 frames  -> ceiling {S / ms * 1e3:9.0f} frames/s", flush=True)
