#!/bin/bash
# f64 parity-kernel variants: steady-state rate of the 256 Ki-env K=10 config for the in-tree library and every variant
for lib in b747_rl_ctrl_b200/lib/libb747_b200.so b747_rl_ctrl_b200/lib/variants/*.so; do
  B747_LIB_PATH=$PWD/$lib timeout 120 python bench.py --config fp64_256k --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'steady %.4f ms  %.3e env-steps/s  lockstep %.4f ms' % (d['ms_per_step'], d['value'], d['lockstep']['ms_per_step']))"
done
