import sys, time, torch, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device('cuda:0')
print('cpus', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for mb in (1, 8, 25, 41):
    h = torch.rand(mb * 1024 * 1024 // 4).pin_memory()
    d = torch.empty_like(h, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1) / 10
    print(f'H2D {mb} MB pinned: {ms:.3f} ms/copy = {mb / 1024 / (ms * 1e-3):.1f} GB/s  (wall {1e3 * (t1 - t0) / 10:.3f} ms)')
    # D2H
    e0.record()
    for _ in range(10): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f'D2H {mb} MB: {e0.elapsed_time(e1) / 10:.3f} ms')
# item() latency
x = torch.ones(1, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100): x.item()
print('item() latency us', 1e6 * (time.perf_counter() - t0) / 100)
import subprocess
print(subprocess.run(['nvidia-smi', '--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max', '--format=csv'], capture_output=True, text=True).stdout)
# with nvidia-smi polling in background
p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm', '--format=csv,noheader', '-lms', '100'], stdout=subprocess.DEVNULL)
time.sleep(0.5)
h = torch.rand(41 * 1024 * 1024 // 4).pin_memory(); d = torch.empty_like(h, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    d.copy_(h, non_blocking=True); x.item()
print('with smi polling: copy+item ms', 1e3 * (time.perf_counter() - t0) / 20)
p.terminate()
time.sleep(0.3)
t0 = time.perf_counter()
for _ in range(20):
    d.copy_(h, non_blocking=True); x.item()
print('without smi polling: copy+item ms', 1e3 * (time.perf_counter() - t0) / 20)
