import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, torch
from drprg_b200 import lib, workload
wl = workload.Config2(); d, o = wl.reads(1000000, 0); words, _, lens = lib.pack_reads(d, o, 10)
ix = lib.Index(wl.prg_path, 11, 15); opts = lib.make_opts(illumina=True)
hw = torch.from_numpy(words.view(np.int32)).pin_memory(); hl = torch.from_numpy(lens.view(np.int32)).pin_memory()
n = len(lens); tb = int(o[-1])
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b = ix.upload_ptrs(hw.data_ptr(), hl.data_ptr(), n, 10, tb); t1 = time.perf_counter()
    ix.sample_begin(opts, 150); t2 = time.perf_counter()
    ix.map_batch(b); t3 = time.perf_counter()
    ix.genotype(wl.refs_path); t4 = time.perf_counter()
    v = ix.vcf_bytes(); b.free(); t5 = time.perf_counter()
    print("upload %.3f begin %.3f map %.3f gt %.3f vcf+free %.3f total %.3f ms" % tuple(x * 1e3 for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)), ix.last_timings())
