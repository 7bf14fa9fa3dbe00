import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, touchgs_b200 as T
cfg = dict(T.synth.CONFIGS["c3"]); N = cfg["N"]; dev = torch.device("cuda:0")
t0 = time.perf_counter()
scene, params, batches, bg = bench.make_workload(cfg, N, 8, dev, 0, None)
print("setup s", time.perf_counter() - t0, "threads", torch.get_num_threads(), flush=True)
st = bench.Stepper(cfg, params, bg, dev, None, None)
def loop(tag, n=100):
    ws = []
    torch.cuda.synchronize()
    for i in range(n):
        t = time.perf_counter()
        st.device_step(batches[i % 8])
        ws.append(time.perf_counter() - t)
    torch.cuda.synchronize()
    w = sorted(ws)
    print(f"{tag}: mean {1e3*sum(ws)/n:.2f} ms  p50 {1e3*w[n//2]:.2f}  p90 {1e3*w[int(n*.9)]:.2f}  max {1e3*w[-1]:.2f}  first5 {[round(1e3*x,1) for x in ws[:5]]} mem {torch.cuda.memory_reserved()/2**30:.1f} GiB", flush=True)
loop("loop1 (cold)")
loop("loop2")
loop("loop3")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,memory.used", "--format=csv"], capture_output=True, text=True).stdout)
time.sleep(2.0)
loop("after 2s idle")
torch.set_num_threads(1)
loop("threads=1")
