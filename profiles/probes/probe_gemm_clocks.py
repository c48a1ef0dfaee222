"""SM clock and board power while the dense-B pipeline, the packed dequant-GEMM and cuBLAS run
back to back for ~3 s each (is the packed kernel's deficit a power-cap clock effect?)."""
import os, subprocess, sys, threading, time, torch
sys.path.insert(0, os.getcwd())
from mxq_b200 import ops
dev = torch.device("cuda:0")
M, OC, IC = 2048, 28672, 8192
W = (torch.randn(OC, IC, device=dev) * 0.02).half()
x = torch.randn(M, IC, device=dev).half()
p = ops.pack(W)
y = torch.empty(M, OC, device=dev, dtype=torch.float16)
ws = torch.zeros(1024, dtype=torch.uint8, device=dev)
samples = []
stop = False
def sampler():
    while not stop:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True)
        samples.append((time.time(), r.stdout.strip()))
        time.sleep(0.05)
def run(name, fn, secs=3.0):
    global samples, stop
    samples, stop = [], False
    th = threading.Thread(target=sampler); th.start()
    torch.cuda.synchronize(); t0 = time.time(); n = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.time() - t0 < secs:
        for _ in range(20): fn()
        n += 20
        torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    stop = True; th.join()
    us = a.elapsed_time(b) * 1e3 / n
    mid = [s for _, s in samples[len(samples) // 3:]]
    clk = sorted(float(s.split(",")[0]) for s in mid)
    pw = sorted(float(s.split(",")[1]) for s in mid)
    print(f"{name}: {us:.1f} us/iter = {2.0 * M * OC * IC / us / 1e6:.0f} TF | SM clock median {clk[len(clk)//2]:.0f} MHz, power median {pw[len(pw)//2]:.0f} W, reasons {mid[-1].split(',')[2].strip()}", flush=True)
run("cuBLAS", lambda: torch.matmul(x, W.t(), out=y))
run("dense-B pipeline", lambda: ops.gemm_dense(x, W))
run("packed", lambda: ops.gemm(x, p, out=y, workspace=ws, validate=False))
os.environ["MXQ_GEMM_DBG"] = "480"
run("packed, producers no-op", lambda: ops.gemm(x, p, out=y, workspace=ws, validate=False))
os.environ["MXQ_GEMM_DBG"] = "0"
os.environ["MXQ_GEMM_SINGLE"] = "1"
run("packed, single-CTA kernel", lambda: ops.gemm(x, p, out=y, workspace=ws, validate=False))
