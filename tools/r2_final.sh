#!/bin/bash
# round-2 evidence visit: GPU tests, bench (every config), launch list and one full ncu capture of the steady-state kernel
out=gpurun_out; tag=${1:-r2z}
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -rf > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/${tag}_pytest.log
timeout 400 python bench.py --steps 300 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err; echo "ref rc=$?"
for c in fp64_4096 fp64_256k tier2_1M_K10; do
  timeout 200 python bench.py --config $c --steps 50 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_$c.json 2>> $out/${tag}_bench.err; echo "bench $c rc=$?"
done
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench20.json 2>> $out/${tag}_bench.err &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 100 python tools/steady_run.py 10 20 > $out/${tag}_steady.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_env_step32 -s 205 -c 1 -o $out/${tag}_step32 -f \
   python tools/steady_run.py 10 20 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
cat $out/${tag}_steady.log
python - <<PY
import json
for f in ("bench","bench20","bench_fp64_4096","bench_fp64_256k","bench_tier2_1M_K10","bench_ref"):
    try:
        d=json.load(open("$out/${tag}_%s.json"%f))
        print(f, "value %.4e ms %.4f e2e %.4e lockstep %s episodes %s"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("lockstep",{}).get("ms_per_step"), d.get("episode_stats",{}).get("episodes")))
    except Exception as e: print(f, "ERR", e)
PY
