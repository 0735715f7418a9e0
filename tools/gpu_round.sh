#!/bin/bash
# One GPU visit: parity tests, bench, launch list, one full ncu capture of the step kernel.
# usage: tools/gpu_round.sh <tag>
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
timeout 600 python bench.py --steps 300 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cat $out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err; echo "ref rc=$?"
cat $out/${tag}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
   python bench.py --steps 40 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_step32 -s 5 -c 2 -o $out/${tag}_step32 -f \
   python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $out/${tag}_step32.ncu-rep --page raw --csv > $out/${tag}_raw.csv 2>/dev/null
ls -la $out
