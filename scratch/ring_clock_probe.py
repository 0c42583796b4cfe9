"""1-GPU leg of the ring bench (40 000 x 4^9, one triangular launch of ~100-200 ms) with nvidia-smi sampling
clocks / power / throttle reasons every 20 ms beside it: is the spread between boxes (102 vs 191 ms) a clock cap?"""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import ring_bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
log = open("gpurun_out/ring_clock_probe.csv", "w")
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active",
                        "--format=csv,noheader", "-lms", "20", "-i", "0"], stdout=log)
time.sleep(0.5)
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
for rep in range(3):
    r = ring_bench.run_gram_ring(None, dev, 0, 1, 1653.1, reps=2, n_total=n_total)
    print(json.dumps({"rep": rep, "ms": r["ms"], "block_ms": r["block_ms"], "prepare_ms": r["prepare_ms"]}), flush=True)
    time.sleep(1.0)
smi.terminate()
log.close()
rows = [l.strip().split(", ") for l in open("gpurun_out/ring_clock_probe.csv") if l.strip()]
sm = sorted(int(r[1].split()[0]) for r in rows)
pw = sorted(float(r[3].split()[0]) for r in rows)
print(json.dumps({"samples": len(rows), "sm_mhz_min": sm[0], "sm_mhz_median": sm[len(sm) // 2], "sm_mhz_max": sm[-1],
                  "power_w_max": pw[-1], "reasons": sorted(set(r[5] for r in rows))}))
