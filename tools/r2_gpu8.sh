#!/bin/bash
# NOTE: an 8-GPU visit is charged 8x the box time: every step carries a tight timeout (a rank dead-lock in bench.py once cost
# 3 x 300 s x 8 GPUs here).
# 8-GPU visit: concurrent host-link probe at N = 1, 2, 4, 8 and the bench at N = 8 (and 2, 4)
out=gpurun_out; tag=${1:-r2n8}
mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> $out/${tag}_topo.txt
for n in 1 2 4 8; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
     tools/pcie_concurrent.py > $out/${tag}_pcie_n$n.txt 2>&1; grep -E "^rank|^N=" $out/${tag}_pcie_n$n.txt
done
for n in 8 4 2; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
     bench.py --gpus $n --steps 100 --warmup 5 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err; echo "bench n=$n rc=$?"
  python -c "
import json,sys
d=json.load(open('$out/${tag}_bench_n$n.json')); print('N=%d value %.3e e2e %.3e lockstep %.3e ms %.4f'%(d['n_gpus'],d['value'],d['e2e']['value'],d['lockstep']['value'],d['ms_per_step']))"
done
