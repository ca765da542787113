# End-of-round evidence: GPU tests, bench lines, refreshed ncu launch lists (inference + train step), attention-backward capture.
R=${1:-r1}
python -m pytest tests -q -m gpu --tb=line 2>&1 | tail -4 > gpurun_out/final_tests_$R.log
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_${R}_n1.json 2> gpurun_out/bench_${R}_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>/dev/null
python bench.py --workload c3 --steps 20 --warmup 3 > gpurun_out/bench_${R}_c3.json 2>/dev/null
python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/bench_${R}_c4_1gpu.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 1 --warmup 1 --skip-cpu > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_$R.csv python tools/ncu_train.py 32 2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_d -f -o gpurun_out/prof_attn_bwd_$R python tools/prof_kernels.py attn_bwd > /dev/null 2>&1
cat gpurun_out/final_tests_$R.log; tail -c 600 gpurun_out/bench_${R}_n1.json; echo; cut -c1-300 gpurun_out/bench_${R}_c3.json; cut -c1-200 gpurun_out/bench_${R}_reference.json; ls -la gpurun_out/*.csv gpurun_out/prof_attn_bwd_$R.ncu-rep
