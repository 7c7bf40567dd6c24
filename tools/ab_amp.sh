#!/bin/bash
# A/B of alternative library builds (build/exp/*.so) on the AMP config-1 bench
show() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'frames/s %.0f' % d['frames_per_s'], 'its %.2f' % d['mean_iterations'], 'ser %.2e' % d['section_error_rate'])"; }
for rep in 1 2; do
  python tools/bench_amp.py --frames 4096 --steps 2 --warmup 1 2>/dev/null | tail -1 | show default
  for lib in build/exp/*.so; do
    LDPC_SPARC_B200_LIB=$PWD/$lib python tools/bench_amp.py --frames 4096 --steps 2 --warmup 1 2>/dev/null | tail -1 | show $(basename $lib)
  done
done
