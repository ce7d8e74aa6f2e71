"""H2D rate of a 28 MB pinned buffer: untouched, just written by this thread, just written by 8 threads (tuning aid for the
plan blob of the streaming fit)."""
import threading, time
import numpy as np, torch
n = 28 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
big = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, pin_memory=True)
dbig = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
def timed(label, prep=None, reps=5):
    ts = []
    for _ in range(reps):
        if prep: prep()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(); d.copy_(h, non_blocking=True); e1.record()
        e1.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{label:40s} {min(ts):.3f} ms  ({n / min(ts) / 1e6:.1f} GB/s)  all: {[round(t, 3) for t in ts]}")
a = h.numpy()
timed("untouched")
timed("written by this thread", lambda: a.fill(7))
def par():
    th = [threading.Thread(target=lambda i=i: a[i * n // 8:(i + 1) * n // 8].fill(i)) for i in range(8)]
    [t.start() for t in th]; [t.join() for t in th]
timed("written by 8 threads", par)
def behind():
    with torch.cuda.stream(s):
        dbig[:92 * 1024 * 1024].copy_(big[:92 * 1024 * 1024], non_blocking=True)
    a.fill(3)
timed("written, queued behind a 92 MB copy", behind)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s):
    e0.record(); dbig.copy_(big, non_blocking=True); e1.record()
e1.synchronize(); print(f"512 MB: {e0.elapsed_time(e1):.2f} ms ({512 * 1.048576 / e0.elapsed_time(e1):.1f} GB/s)")
