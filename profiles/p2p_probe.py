import torch, time, subprocess
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
print("can access peer 0->1:", torch.cuda.can_device_access_peer(0, 1))
for mb in (2.3, 16, 256):
    n = int(mb * 1e6)
    a = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(n, dtype=torch.uint8, device="cuda:1")
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(0):
        e0.record()
        for _ in range(10):
            b.copy_(a)
        e1.record()
        e1.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"{mb} MB cuda:0 -> cuda:1 : {us:.1f} us per copy, {n / us / 1e3:.1f} GB/s")
