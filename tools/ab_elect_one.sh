#!/bin/bash
# A/B of the MMA-issue variants of the tcgen05 layers (profiles/r1_sass_mma_issue.md), to be run ON THE GPU BOX:
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/ab_elect_one.sh'
# 1. bench of the default build (lane 0 by thread index: waterfall loops around every UTCHMMA)
# 2. rebuild with -DEMER_TC_ELECT_ONE=1 (elect.sync), GPU parity tests, smoke, bench
# 3. restore the default build
# Everything lands in gpurun_out/ab_*.  The switch becomes the default only if step 2 is green.
set -u
mkdir -p gpurun_out
B="--steps 120 --warmup 12 --no-cpu-baseline --profile-all"
timeout 300 python bench.py $B > gpurun_out/ab_default.log 2> gpurun_out/ab_default.err
tail -1 gpurun_out/ab_default.log | cut -c1-160

export EMER_TC_ELECT_ONE=1
python -m emernerf_b200.build || exit 1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/ab_elect_pytest.log 2>&1
tail -2 gpurun_out/ab_elect_pytest.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python bench.py $B > gpurun_out/ab_elect.log 2> gpurun_out/ab_elect.err
tail -1 gpurun_out/ab_elect.log | cut -c1-160
grep "^# emer_linear_tc" gpurun_out/ab_default.err | sort > gpurun_out/ab_default_tc.txt
grep "^# emer_linear_tc" gpurun_out/ab_elect.err | sort > gpurun_out/ab_elect_tc.txt
paste -d'|' gpurun_out/ab_default_tc.txt gpurun_out/ab_elect_tc.txt | cut -c1-230

unset EMER_TC_ELECT_ONE
python -m emernerf_b200.build
