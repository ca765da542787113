set -x
R=r1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 1 --warmup 1 --skip-cpu > gpurun_out/launches_bench.log 2>&1
for t in attn gemm wgrad attn_bwd; do
  case $t in attn) k=attention_tc;; gemm) k=gemm_tc;; wgrad) k=wgrad_tc;; attn_bwd) k=attn_bwd_d;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -f -o gpurun_out/prof_${t}_$R python tools/prof_kernels.py $t > gpurun_out/prof_$t.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
