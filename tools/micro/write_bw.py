import torch
dev = torch.device("cuda")
for mb in (256, 736, 2048):
    t = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    src = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for name, fn in (("memset", lambda: t.zero_()), ("copy", lambda: t.copy_(src))):
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        traffic = mb * (1 if name == "memset" else 2) / 1024
        print(f"{name} {mb} MiB: {best*1e3:.1f} us -> {traffic / (best/1e3):.0f} GiB/s of DRAM traffic")
