#!/bin/bash
# Host-side ingest code (FASTQ framer, parallel gunzip) under AddressSanitizer + UBSan on generated inputs: plain / no final
# newline / ragged FASTQ, gzip level 6, bgzip-style members, bit-flipped and truncated gzip files.  No GPU needed.
#   bash tools/asan_ingest.sh
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
python - "$T" <<'PY'
import gzip, random, sys, zlib
import numpy as np
from drprg_b200 import sim
T = sys.argv[1]
rng = np.random.default_rng(5)
n, L = 40000, 150
d = sim.BASES[rng.integers(0, 4, size=n * L)]
o = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
sim.write_fastq_fast(f"{T}/a.fq", d, o)
t = open(f"{T}/a.fq", "rb").read()
open(f"{T}/a_nonl.fq", "wb").write(t[:-1])
co = zlib.compressobj(6, zlib.DEFLATED, 31)
z = co.compress(t) + co.flush()
open(f"{T}/a6.fq.gz", "wb").write(z)
with open(f"{T}/a_bgz.fq.gz", "wb") as f:
    for i in range(0, len(t), 65280):
        f.write(gzip.compress(t[i:i + 65280], 6, mtime=0))
    f.write(gzip.compress(b"", 6, mtime=0))
recs = []
for i in range(5000):
    l = int(rng.integers(0, 900))
    s = bytes(sim.BASES[rng.integers(0, 4, size=l)])
    recs.append(b"@r%d\n" % i + s + b"\n+\n" + b"F" * l + b"\n")
open(f"{T}/ragged.fq", "wb").write(b"".join(recs))
random.seed(3)
for k in range(6):
    c = bytearray(z)
    for _ in range(1 + k):
        i = random.randrange(20, len(c) - 8)
        c[i] ^= 1 << random.randrange(8)
    open(f"{T}/corrupt{k}.fq.gz", "wb").write(bytes(c))
open(f"{T}/corrupt6.fq.gz", "wb").write(z[:len(z) // 2])
open(f"{T}/corrupt7.fq.gz", "wb").write(z[:len(z) - 3])
PY
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=c++17 -pthread -Idrprg_b200/csrc -I/usr/local/cuda/include \
    tools/asan_ingest_driver.cpp drprg_b200/csrc/fastq_frame.cpp drprg_b200/csrc/genotype_host.cpp drprg_b200/csrc/prg_graph.cpp \
    drprg_b200/csrc/gzip_inflate.cpp -o "$T/asan_driver" -lz 2>/dev/null
DRPRG_PARALLEL_GZIP_CHUNK=65536 DRPRG_PARALLEL_GZIP_GROUP=5 DRPRG_FRAME_SLICE=4096 "$T/asan_driver" "$T"/a.fq "$T"/a_nonl.fq "$T"/ragged.fq \
    "$T"/a6.fq.gz "$T"/a_bgz.fq.gz "$T"/corrupt*.fq.gz
rm -rf "$T"
echo "asan_ingest: done (no sanitizer report above = clean)"
