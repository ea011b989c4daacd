"""SM clock / power under a sustained loop of C2 shoots vs short bursts after idle (nvidia-smi at 20 ms),
and the shoot's time in both regimes: is the in-situ kernel time (event-timed) above the ncu durations
because of the power cap?"""
import sys, os, subprocess, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lagomorph_b200 as lm
dev = torch.device("cuda")
metric = lm.FluidMetric([0.1, 0.0, 0.01])
g = torch.Generator().manual_seed(1)
m0 = torch.randn((16, 3, 128, 128, 128), generator=g).to(dev)
m0.mul_(4.0 / metric.sharp(m0).abs().max().item())
for _ in range(3): lm.expmap(metric, m0, num_steps=10)
torch.cuda.synchronize()
def shoot_ms(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): lm.expmap(metric, m0, num_steps=10)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
p = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active",
                      "--format=csv,noheader", "-lms", "20"], stdout=subprocess.PIPE, text=True)
time.sleep(1.0)
marks = []
for i in range(5):          # bursts: one shoot after 0.5 s idle
    time.sleep(0.5)
    marks.append(("burst", time.time(), shoot_ms(1)))
t0 = time.time()
marks.append(("sustained-start", t0, 0))
res = []
while time.time() - t0 < 6.0:
    res.append(shoot_ms(10))
marks.append(("sustained-end", time.time(), 0))
time.sleep(0.3)
p.terminate()
lines = p.stdout.read().strip().splitlines()
print("burst shoots (ms):", ["%.2f" % m[2] for m in marks if m[0] == "burst"])
print("sustained shoots (ms per shoot, groups of 10):", ["%.2f" % r for r in res[:3]], "...", ["%.2f" % r for r in res[-3:]], "n=%d" % len(res))
import datetime
def ts(s):
    return datetime.datetime.strptime(s.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
a, b = marks[-2][1], marks[-1][1]
inside = [l for l in lines if a + 0.5 <= ts(l.split(",")[0]) <= b]
outside = [l for l in lines if ts(l.split(",")[0]) < a]
def summ(ls, name):
    if not ls: print(name, "no samples"); return
    sm = sorted(int(l.split(",")[1].split()[0]) for l in ls)
    pw = sorted(float(l.split(",")[3].split()[0]) for l in ls)
    print("%s: %d samples, SM MHz min %d median %d max %d; power W median %.0f max %.0f; reasons %s" % (
        name, len(ls), sm[0], sm[len(sm) // 2], sm[-1], pw[len(pw) // 2], pw[-1], sorted(set(l.split(",")[5].strip() for l in ls))))
summ(outside, "idle/bursts")
summ(inside, "sustained")
print("\n".join(inside[::25][:12]))
