"""Control for coresidency_probe.py: the same tiny high-priority kernel next to a long-running one-CTA spin kernel
(torch.cuda._sleep) -- validates that the probe method sees kernel concurrency at all on this box."""
import torch
dev = torch.device("cuda", 0)
lo = torch.cuda.Stream(priority=0)
hi = torch.cuda.Stream(priority=-1)
small = torch.zeros(256, device=dev)
torch.cuda._sleep(1000); small.add_(1.0); torch.cuda.synchronize()      # load both kernels first
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
with torch.cuda.stream(lo):
    e[0].record(); torch.cuda._sleep(20_000_000); e[1].record()
with torch.cuda.stream(hi):
    e[2].record(); small.add_(1.0); e[3].record()
torch.cuda.synchronize()
print("spin kernel %.2f ms; probe span %.3f ms; probe end relative to spin start %.2f ms" % (e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3]), e[0].elapsed_time(e[3])))
