# Round 1, last GPU call: hardware check of the tensor-core scorer (see profiles/r1_notes.md)
mkdir -p gpurun_out
timeout 45 python profiles/try_score_tc.py tc_bf16 > gpurun_out/try_bf16.log 2>&1; RB=$?
if [ $RB -eq 0 ]; then S=tc_bf16; else S=tc_tf32; fi
echo "bf16 rc=$RB -> scorer $S" | tee gpurun_out/choice.txt
DRB_SERVICE_SCORER=$S timeout 50 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_$S.json 2> gpurun_out/bench_$S.err; echo "bench rc=$?" | tee -a gpurun_out/choice.txt
DRB_SERVICE_SCORER=$S timeout 50 python -m pytest tests/test_gpu_service.py -x -q > gpurun_out/svc_$S.log 2>&1; echo "svc rc=$?" | tee -a gpurun_out/choice.txt
timeout 60 ncu --set full --clock-control none --import-source on -k regex:score_msac_tc_kernel -s 1 -c 1 -f -o gpurun_out/tc_full python profiles/ncu_one_tc.py $S > gpurun_out/ncu_tc.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/choice.txt
tail -n 12 gpurun_out/try_bf16.log; cut -c1-1800 gpurun_out/bench_$S.json; tail -n 4 gpurun_out/svc_$S.log; tail -n 3 gpurun_out/ncu_tc.log
