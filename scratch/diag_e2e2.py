import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, touchgs_b200 as T
cfg = dict(T.synth.CONFIGS["c3"]); N = cfg["N"]; dev = torch.device("cuda:0")
scene, params, batches, bg = bench.make_workload(cfg, N, 8, dev, 0, None)
st = bench.Stepper(cfg, params, bg, dev, None, None)
for b in batches: st.device_step(b)
def run(name, fn, n=60, fin=None):
    for i in range(8): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(n): fn(8 + i)
    if fin: fin()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/n:.3f} ms/step (wall {1e3*(time.perf_counter()-t0)/n:.3f})", flush=True)
run("device", lambda i: st.device_step(batches[i % 8]))
st.e2e_begin(batches)
run("e2e full", st.e2e_step, fin=st.e2e_finish)
# variant: sync copies on the current stream, item()
stage = {k: torch.empty_like(v, device=dev) for k, v in batches[0]["host"].items()}
def v_sync(i):
    b = batches[i % 8]
    for k, v in b["host"].items(): stage[k].copy_(v, non_blocking=True)
    loss = st._run(b["cam"], stage["view"], stage["proj"], stage["campos"], stage["gt"], stage["target"], stage["weight"])
    return loss
run("same-stream H2D, no loss read", v_sync)
run("same-stream H2D + item()", lambda i: v_sync(i).item())
def v_noh2d_item(i):
    b = batches[i % 8]; d = b["dev"]
    st._run(b["cam"], d["view"], d["proj"], d["campos"], d["gt"], d["target"], d["weight"]).item()
run("device + item()", v_noh2d_item)
# prefetch only, no loss
cs = torch.cuda.Stream(); stg = [{k: torch.empty_like(v, device=dev) for k, v in batches[0]["host"].items()} for _ in range(2)]
rdy = [torch.cuda.Event(), torch.cuda.Event()]; con = [torch.cuda.Event(), torch.cuda.Event()]
def pf(b, s):
    with torch.cuda.stream(cs):
        for k, v in b["host"].items(): stg[s][k].copy_(v, non_blocking=True)
        rdy[s].record(cs)
pf(batches[0], 0); kk = [0]
def v_pref(i):
    s = kk[0] & 1; b = batches[i % 8]; cur = torch.cuda.current_stream()
    cur.wait_event(rdy[s]); d = stg[s]
    st._run(b["cam"], d["view"], d["proj"], d["campos"], d["gt"], d["target"], d["weight"])
    con[s].record(cur); cs.wait_event(con[s ^ 1]); pf(batches[(i + 1) % 8], s ^ 1); kk[0] += 1
run("prefetch H2D, no loss read", v_pref)
# prefetch only big tensors? copy only gt/target/weight
