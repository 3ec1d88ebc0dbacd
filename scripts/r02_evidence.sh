#!/bin/bash
# Round-2 evidence run on one B200 (under gpurun): bench lines, ncu launch list + full capture of one solve of the
# benched build, compute-sanitizer logs.  Everything lands in gpurun_out/ (copied to profiles/ afterwards).
set -u
O=gpurun_out
T=r02zz
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/${T}_clocks.csv 2>/dev/null &
SMI=$!
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > $O/${T}_smoke.log 2>&1
# 1. headline: cfg4 (default flags), then cfg2 / cfg5 stand-in / cfg3 stand-in / the reference's own benchmark shape
timeout 1500 python bench.py > $O/${T}_bench_cfg4.json 2> $O/${T}_bench_cfg4.err
timeout 600 python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err
timeout 600 python bench.py --workload cfg5 --no-cpu-baseline > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
CHB_CFG3_SIDE=48 timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > $O/${T}_bench_cfg3_48.json 2> $O/${T}_bench_cfg3_48.err
timeout 600 python bench.py --workload ico6 --k 128 --no-cpu-baseline > $O/${T}_bench_ico6.json 2> $O/${T}_bench_ico6.err
# 2. the round-1 execution (one persistent kernel per sweep) of the same build, for the record
CHB_LOCAL_KERNEL=0 timeout 900 python bench.py --no-cpu-baseline > $O/${T}_bench_cfg4_persistent_only.json 2> $O/${T}_bench_cfg4_persistent_only.err
# 3. ncu: launch list of a bench run, full capture of the launches of ONE solve of the benched build (cfg4)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${T}_launches_cfg4.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_bench_under_ncu.json 2>/dev/null
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:chb_(sweep|local|wide)' -s 8 -c 4 -o $O/${T}_solve_cfg4 python scripts/solve_once.py cfg4 f 128 3 > $O/${T}_ncu_cfg4.log 2>&1
kill $SMI 2>/dev/null
if [ "${CHB_EVIDENCE_SANITIZER:-1}" = 0 ]; then exit 0; fi
# 4. compute-sanitizer on a small problem: every executor (local-task kernel, persistent kernel, wide-level kernel)
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/solve_once.py ico5 f 128 2 > $O/${T}_sanitizer_${tool}_ico5_f32.log 2>&1
  CHB_WIDE_MIN_TASKS=1 timeout 900 compute-sanitizer --tool $tool python scripts/solve_once.py grid3d-10 d 33 2 > $O/${T}_sanitizer_${tool}_grid3d_f64_wide.log 2>&1
done
grep -H "ERROR SUMMARY\|RACECHECK SUMMARY" $O/${T}_sanitizer_*.log
