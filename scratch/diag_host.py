import sys, os, time, torch, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, touchgs_b200 as T
cfg = dict(T.synth.CONFIGS["c3"]); dev = torch.device("cuda:0")
print("start", flush=True)
scene, params, batches, bg = bench.make_workload(cfg, 20000, 2, dev, 0, None)
print("workload ok", flush=True)
st = bench.Stepper(cfg, params, bg, dev, None, None)
for i in range(20):
    st.device_step(batches[i % 2]); torch.cuda.synchronize(); print("step", i, flush=True)
t0 = time.perf_counter()
for i in range(200): st.device_step(batches[i % 2])
torch.cuda.synchronize()
print("host floor per step (N=20k): %.3f ms" % (1e3 * (time.perf_counter() - t0) / 200))
pr = cProfile.Profile(); pr.enable()
for i in range(200): st.device_step(batches[i % 2])
torch.cuda.synchronize(); pr.disable()
ps = pstats.Stats(pr); ps.sort_stats("cumulative").print_stats(28)
