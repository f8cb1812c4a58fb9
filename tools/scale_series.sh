set -x
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_b11_n$n.json 2> gpurun_out/r2_b11_n$n.err
  tail -c 1500 gpurun_out/r2_b11_n$n.json | head -c 200; echo
done
python bench.py --gpus 1 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_b11_n1.json 2> gpurun_out/r2_b11_n1.err
