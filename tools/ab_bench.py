"""A/B timing of the sketch kernel across builds: python tools/ab_bench.py lib1.so lib2.so ..."""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for so in sys.argv[1:]:
    env = dict(os.environ, DRPRG_CUDA_LIB=os.path.abspath(so), DRPRG_SKETCH_VARIANT=os.environ.get("V", "0"))
    print(so, flush=True)
    subprocess.run([sys.executable, os.path.join(here, "variant_bench.py"), "run"], env=env)
