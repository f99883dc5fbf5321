import torch, time
x = torch.randn(16, 32, 3, 224, 224, device="cuda")
v = x.permute(0, 2, 1, 3, 4)
h = torch.empty(v.shape).pin_memory()
hc = torch.empty(x.shape).pin_memory()
def t(f, n=5):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    f(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print("d2h permuted view -> pinned: %.2f ms" % t(lambda: h.copy_(v, non_blocking=True)))
print("d2h contiguous -> pinned:    %.2f ms" % t(lambda: hc.copy_(x, non_blocking=True)))
print("device permute copy:         %.2f ms" % t(lambda: v.contiguous()))
print("h2d pinned -> device:        %.2f ms" % t(lambda: x.copy_(hc, non_blocking=True)))
