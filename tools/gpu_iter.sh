#!/bin/bash
# Quick GPU iteration: parity tests, bench without the CPU leg, optional extras.  usage: tools/gpu_iter.sh <tag> [extra cmd]
tag=${1:-it}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $out/${tag}_pytest.log
timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cat $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
shift
if [ -n "$1" ]; then eval "$@"; fi
