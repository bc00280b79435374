"""Where the host time of bench.py's timed e2e run goes, call by call (prefetcher next / draw_plan / feed_data / read-back /
final wait), for the three e2e variants alternated as bench.py runs them (fp32 feed, uint8 feed on one stream, uint8 feed
on four).  Used to tell environmental noise (shared host: PCIe / CPU contention from other tenants) from start-up effects:
on a quiet box every repetition lands within 2 % (68.6 k / 234 k / 256 k pairs/s at K = 20).    python profiles/e2e_calls.py"""
import sys, time
sys.path.insert(0, "/root/repo")
import torch, bench
import trainner_redux_b200.prefetch as P
wl = bench.Workload("c2", 1)
arm = bench.Arm(wl, torch.device("cuda:0"), 0, 1, None)
arm.feed.graphs.credits = arm.feed.graphs.max_credits = 1000.0
arm.feed.graphs.capacity = 64
for i in range(12): arm.step(i)
arm.barrier()
log = []
orig_fd = arm.feed.feed_data
def fd(*a, **k):
    t0 = time.perf_counter(); r = orig_fd(*a, **k); log.append(("feed", time.perf_counter() - t0)); return r
arm.feed.feed_data = fd
orig_next = P.CUDAPrefetcher.next
def nx(self):
    t0 = time.perf_counter(); r = orig_next(self); log.append(("next", time.perf_counter() - t0)); return r
P.CUDAPrefetcher.next = nx
orig_read = P.CUDAReadback.read
def rd(self, t):
    t0 = time.perf_counter(); r = orig_read(self, t); log.append(("read", time.perf_counter() - t0)); return r
P.CUDAReadback.read = rd
orig_wait = P.CUDAReadback.wait
def wt(self):
    t0 = time.perf_counter(); r = orig_wait(self); log.append(("wait", time.perf_counter() - t0)); return r
P.CUDAReadback.wait = wt
orig_plan = arm.plan
def pl():
    t0 = time.perf_counter(); r = orig_plan(); log.append(("plan", time.perf_counter() - t0)); return r
arm.plan = pl
seq = [(False, 1), (True, 1), (True, 4)] * 4
for rep, (u8, lanes) in enumerate(seq):
    log.clear()
    cap0 = arm.feed.graphs.captures
    c, _, _ = arm.e2e(20, 5, u8, lanes)
    # the timed run = the entries after the first 'wait' (end of warm-up)
    i = [j for j, (k, _) in enumerate(log) if k == "wait"][0]
    timed = log[i + 1:]
    tot = {}
    for k, v in timed: tot[k] = tot.get(k, 0) + v
    big = [(k, round(v * 1e3, 3)) for k, v in timed if v > 0.4e-3]
    print(f"rep {rep} u8={u8} lanes={lanes} captures+{arm.feed.graphs.captures - cap0}: {c/1e3:.1f} k; timed-run host ms by call: { {k: round(v*1e3,2) for k,v in tot.items()} }; calls > 0.4 ms: {big}", flush=True)
