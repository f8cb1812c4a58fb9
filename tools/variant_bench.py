import sys, os, subprocess, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
if len(sys.argv) > 1:
    from drprg_b200 import lib, workload
    import numpy as np
    wl = workload.Config2(); d, o = wl.reads(1000000, 0); words, _, lens = lib.pack_reads(d, o, 10)
    ix = lib.Index(wl.prg_path, 11, 15); opts = lib.make_opts(illumina=True)
    b = ix.upload(words, None, lens, total_bases=int(o[-1]), stride_words=10)
    ts = []
    for i in range(8):
        ix.sample_begin(opts, 150); nh, nk = ix.map_batch(b); ts.append(ix.last_timings()["sketch_lookup"])
    print(json.dumps({"variant": os.environ.get("DRPRG_SKETCH_VARIANT"), "sketch_ms_min": min(ts[2:]), "hits": nh, "kept": nk}))
else:
    for v in "0123":
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, DRPRG_SKETCH_VARIANT=v))
