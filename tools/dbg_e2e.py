import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from pcrcg_b200 import pipeline
cfg, limits = bench.workload_config("3dmatch")
pairs = bench.make_pairs("3dmatch", 16, 0)
pts, lens = pipeline.stack_pairs(pairs)
ph, lh = torch.from_numpy(pts).pin_memory(), torch.from_numpy(lens).pin_memory()
path = pipeline.FeaturePath(cfg, limits, device="cuda:0")
pd, ld = ph.cuda(), lh.cuda()
for i in range(4):
    torch.cuda.synchronize(); t=time.perf_counter(); y,_ = path.run_device(pd, ld); torch.cuda.synchronize(); print("dev", i, round((time.perf_counter()-t)*1e3,1), torch.cuda.memory_reserved()>>20, "MB reserved")
out_host = torch.empty((y.shape[0]+1024, y.shape[1])).pin_memory()
for i in range(6):
    torch.cuda.synchronize(); t=time.perf_counter(); o,_ = path.run_host(ph, lh, out_host); torch.cuda.synchronize(); print("host", i, round((time.perf_counter()-t)*1e3,1), torch.cuda.memory_reserved()>>20)
for i in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); y,_ = path.run_device(pd, ld); torch.cuda.synchronize(); print("dev", i, round((time.perf_counter()-t)*1e3,1))
print(torch.cuda.memory_summary()[:1500])
