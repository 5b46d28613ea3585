"""What limits a stream of 100 MB pinned H2D copies: issue pattern experiments (wall clock per copy)."""
import time, torch
dev = torch.device("cuda", 0)
n = 20 * 768 * 22 * 76
hs = [torch.randn(n).pin_memory() for _ in range(2)]
ds = [torch.empty(n, device=dev) for _ in range(2)]
cs = torch.cuda.Stream()
def run(name, body, reps=30):
    body(3); torch.cuda.synchronize(); t0 = time.perf_counter(); body(reps); torch.cuda.synchronize()
    print("%-46s %8.1f us/copy" % (name, (time.perf_counter() - t0) / reps * 1e6))
def a(k):
    for i in range(k): ds[i % 2].copy_(hs[i % 2], non_blocking=True)
def b(k):
    with torch.cuda.stream(cs):
        for i in range(k):
            ds[i % 2].copy_(hs[i % 2], non_blocking=True); e = torch.cuda.Event(); e.record(cs)
def c(k):
    evs = [None, None]
    with torch.cuda.stream(cs):
        for i in range(k):
            ds[i % 2].copy_(hs[i % 2], non_blocking=True); e = torch.cuda.Event(); e.record(cs)
            if evs[(i + 1) % 2] is not None: evs[(i + 1) % 2].synchronize()
            evs[i % 2] = e
def d(k):   # one-ahead, but wait by polling query() instead of a blocking synchronize
    evs = [None, None]
    with torch.cuda.stream(cs):
        for i in range(k):
            ds[i % 2].copy_(hs[i % 2], non_blocking=True); e = torch.cuda.Event(); e.record(cs)
            p = evs[(i + 1) % 2]
            if p is not None:
                while not p.query(): pass
            evs[i % 2] = e
def e4(k):  # 4 chunk copies per step, back to back
    m = n // 4
    for i in range(k):
        for c_ in range(4): ds[i % 2][c_ * m:(c_ + 1) * m].copy_(hs[i % 2][c_ * m:(c_ + 1) * m], non_blocking=True)
def f(k):   # two-ahead
    evs = []
    with torch.cuda.stream(cs):
        for i in range(k):
            ds[i % 2].copy_(hs[i % 2], non_blocking=True); e = torch.cuda.Event(); e.record(cs); evs.append(e)
            if len(evs) > 2: evs.pop(0).synchronize()
run("back to back, no events", a)
run("back to back, event record after each", b)
run("one ahead, event.synchronize on i-1", c)
run("one ahead, event.query spin on i-1", d)
run("4 chunk copies per step, back to back", e4)
run("two ahead, event.synchronize on i-2", f)
print("-- again, reverse order")
run("two ahead, event.synchronize on i-2", f)
run("one ahead, event.synchronize on i-1", c)
run("back to back, no events", a)
run("back to back, no events (200 copies)", a, 200)
run("one ahead, event.synchronize on i-1 (200)", c, 200)
