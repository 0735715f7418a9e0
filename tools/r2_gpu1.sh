#!/bin/bash
# round 2, GPU visit 1: tests (no -x: every failure is information), bench configs, host-mode probe, family probe, launch list
out=gpurun_out; tag=${1:-r2a}
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -rf --durations=8 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -40 $out/${tag}_pytest.log
timeout 600 python bench.py --steps 300 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cat $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench20.json 2>> $out/${tag}_bench.err; echo "bench20 rc=$?"
for c in fp64_4096 fp64_256k tier2_1M_K10; do
  timeout 300 python bench.py --config $c --steps 50 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_$c.json 2>> $out/${tag}_bench.err; echo "bench $c rc=$?"
  cat $out/${tag}_bench_$c.json
done
timeout 300 python tools/host_mode_probe.py > $out/${tag}_host_modes.txt 2>&1; cat $out/${tag}_host_modes.txt
timeout 300 python tools/pcie_probe.py > $out/${tag}_pcie.txt 2>&1; cat $out/${tag}_pcie.txt
timeout 900 python tools/family_probe.py $out/${tag}_family_probe.npz > $out/${tag}_family_probe.txt 2>&1; cat $out/${tag}_family_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches.csv \
   python bench.py --steps 40 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ls -la $out | head -40
