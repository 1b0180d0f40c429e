"""Host<->device copy bandwidth of the box (pinned memory): H2D alone, D2H alone, both directions at once.
The e2e number of bench.py moves 1.22 GB in and 1.60 GB out per step; this says what the link allows."""
import json, time, torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return (n * (int(h2d) + int(d2h))) / dt / 1e9
run(True, True, 2)
out = {"h2d_GBps": run(True, False), "d2h_GBps": run(False, True), "both_total_GBps": run(True, True)}
print(json.dumps(out))
