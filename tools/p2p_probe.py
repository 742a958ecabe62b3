"""Peer-to-peer facts of the box: topology, peer-access flags, and device-to-device copy bandwidth for a few pairs."""
import subprocess, time, torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
n = torch.cuda.device_count()
print("can_access_peer:")
for i in range(n):
    print(i, [int(torch.cuda.can_device_access_peer(i, j)) if i != j else -1 for j in range(n)])
x = [torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % i) for i in range(n)]
for (a, b) in [(0, 1), (0, 2), (0, 3), (0, n - 1), (2, 3), (1, n - 1)]:
    if a >= n or b >= n or a == b:
        continue
    for _ in range(2):
        x[b].copy_(x[a])
    torch.cuda.synchronize(a); torch.cuda.synchronize(b)
    t = time.perf_counter()
    for _ in range(5):
        x[b].copy_(x[a])
    torch.cuda.synchronize(a); torch.cuda.synchronize(b)
    dt = (time.perf_counter() - t) / 5
    print("copy %d -> %d : %.1f GB/s" % (a, b, (256 << 20) / dt / 1e9))
