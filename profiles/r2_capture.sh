#!/bin/bash
# Round-2 evidence capture (run under gpurun from the repo root): GPU tests, smoke, bench lines of the five configs,
# the reference arm, the ncu launch list of the cfg2 bench and ncu --set full captures of the two kernels of the step.
set -u
TAG=${1:-f}
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_gputest_$TAG.log 2>&1; tail -3 gpurun_out/r2_gputest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_$TAG.log 2>&1; tail -1 gpurun_out/r2_smoke_$TAG.log
for c in cfg2 cfg5 cfg3 cfg4 cfg1; do
  timeout 240 python bench.py --config $c > gpurun_out/r2_bench_${c}_$TAG.json 2> gpurun_out/r2_bench_${c}_$TAG.err
done
timeout 240 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_$TAG.json 2> gpurun_out/r2_bench_ref_$TAG.err
DRB_BENCH_GATE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launches_bench_$TAG.log 2>&1
[ -n "${SKIP_NCU_FULL:-}" ] || timeout 200 ncu --set full --clock-control none --cache-control none --import-source on -k regex:score_msac_tc_kernel -s 1 -c 1 \
    -o gpurun_out/r2_tc_bf16p_s -f python profiles/ncu_one_tc.py tc_bf16p_s > gpurun_out/ncu_tc_$TAG.log 2>&1
[ -n "${SKIP_NCU_FULL:-}" ] || timeout 200 ncu --set full --clock-control none --import-source on -k regex:solve_e5_kernel -s 1 -c 1 \
    -o gpurun_out/r2_solve_e5_160x4 -f python profiles/ncu_one_tc.py tc_bf16p_s > gpurun_out/ncu_e5_$TAG.log 2>&1
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${c}_$TAG.json"))
print("$c", round(d["value"]/1e6,3), "M hyp/s", round(d["ms_per_step"],4), "ms e2e", round(d["e2e"]["value"]/1e6,3))
PY
done
