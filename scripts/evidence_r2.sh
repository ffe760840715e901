# round-2 evidence capture (run on a B200 through gpurun): GPU tests, bench (ours + reference arm), launch list, ncu of the
# headline kernel and of the reward-net backward launches, reward-update timing, parity table, compute-sanitizer
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_final.log 2>&1; tail -3 gpurun_out/r2_gputest_final.log
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 300 gpurun_out/r2_bench_1gpu.json
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; tail -c 200 gpurun_out/r2_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-big-modes --no-cpu-baseline > gpurun_out/r2_b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_v2 -s 1 -c 1 -o gpurun_out/r2_v2_train -f python scripts/prof_rollout.py --log2-pops 20 > gpurun_out/r2_v2_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_v2 -s 1 -c 1 -o gpurun_out/r2_v2_record -f python scripts/prof_rollout.py --log2-pops 18 --mode record > gpurun_out/r2_v2_record.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rnet_kernel -s 4 -c 2 -o gpurun_out/r2_rnet_bwd_final -f python scripts/prof_irl.py 2 > gpurun_out/r2_rnet_bwd_final.log 2>&1
python scripts/time_irl_update.py 2>/dev/null | tail -1 > gpurun_out/r2_irl_time_final.log; cat gpurun_out/r2_irl_time_final.log
python tests/parity_maxerr.py > gpurun_out/r2_parity_maxerr.md 2> gpurun_out/r2_parity_maxerr.err; tail -3 gpurun_out/r2_parity_maxerr.md
scripts/sanitize_run.sh | tail -12
