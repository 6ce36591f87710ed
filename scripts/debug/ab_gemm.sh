#!/bin/bash
# same-box A/B of source trees (debug helper): per-contraction timings and the C2 step of each tree given on the command line,
# e.g.   git archive <commit> | tar -x -C build/old_tree && (cd build/old_tree && python -m tacotron_b200.build)
#        gpurun -- bash scripts/debug/ab_gemm.sh . build/old_tree
for tree in "$@"; do
  echo "== $tree"
  (cd $tree; python /root/repo/scripts/debug/gemm_time.py 2>&1 | tail -6; python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['sections_ms'], d['decoder_step_p50_us'])")
done
