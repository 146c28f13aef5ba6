import time, torch, numpy as np
torch.cuda.set_device(0)
n = 384 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("pinned D2H 384 MB: %.1f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("pinned H2D 384 MB: %.1f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
# host memcpy pinned -> fresh pageable with 16 threads
import ctypes, threading
src = h.numpy()
for rep in range(3):
    dst = np.empty(n, np.uint8)
    t0 = time.perf_counter()
    ths = [threading.Thread(target=lambda i=i: np.copyto(dst[i * (n // 16):(i + 1) * (n // 16)], src[i * (n // 16):(i + 1) * (n // 16)])) for i in range(16)]
    [t.start() for t in ths]; [t.join() for t in ths]
    dt = time.perf_counter() - t0
    print("16-thread memcpy pinned -> fresh pageable 384 MB: %.1f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
dst = np.empty(n, np.uint8); dst[:] = 0
t0 = time.perf_counter()
ths = [threading.Thread(target=lambda i=i: np.copyto(dst[i * (n // 16):(i + 1) * (n // 16)], src[i * (n // 16):(i + 1) * (n // 16)])) for i in range(16)]
[t.start() for t in ths]; [t.join() for t in ths]
dt = time.perf_counter() - t0
print("16-thread memcpy pinned -> touched pageable 384 MB: %.1f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
