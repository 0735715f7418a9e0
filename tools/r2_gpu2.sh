#!/bin/bash
out=gpurun_out; tag=${1:-r2f}
mkdir -p $out
timeout 600 python tools/phase_time_probe.py > $out/${tag}_phase.txt 2>&1; cat $out/${tag}_phase.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_plain.json 2> $out/${tag}_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_step32 -s 230 -c 2 -o $out/${tag}_step32 -f \
   python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $out | grep $tag
