#!/bin/bash
# A/B timing of alternative builds of the library (build/exp/*.so) on the bench workload.
# usage: tools/ab_bench.sh [steps]
steps=${1:-5}
show() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'ms/step %.3f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'fe', d['monte_carlo']['frame_errors'], 'be', d['monte_carlo']['bit_errors'], 'its %.3f' % d['monte_carlo']['mean_iterations'])"; }
for rep in 1 2; do
  python bench.py --steps $steps --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-mc --no-extras 2>/dev/null | show default
  for lib in build/exp/*.so; do
    LDPC_SPARC_B200_LIB=$PWD/$lib python bench.py --steps $steps --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-mc --no-extras 2>/dev/null | show $(basename $lib)
  done
done
