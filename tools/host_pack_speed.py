import sys, time
sys.path.insert(0, '.')
import numpy as np
import scrooge_b200
l = scrooge_b200.lib()
print("isa", l.sg_host_pack_isa())
n = 2_000_000_000
a = np.frombuffer(b"ACGT" * (n // 4), dtype=np.uint8).copy()
out = np.zeros(n // 16 + 8, dtype=np.uint32)
l.sg_host_pack_2bit(a.ctypes.data, n, out.ctypes.data, 16)
for th in (1, 2, 4, 8, 12, 16):
    best = 1e9
    for _ in range(3):
        t = time.time(); l.sg_host_pack_2bit(a.ctypes.data, n, out.ctypes.data, th); best = min(best, time.time() - t)
    print(th, "threads", round(n / best / 1e9, 1), "GB/s of ASCII")
