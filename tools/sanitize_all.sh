# compute-sanitizer evidence for profiles/sanitizer_rNN.txt
R=${1:-r1}
OUT=gpurun_out/sanitizer_$R.txt
echo "# compute-sanitizer, round ${R#r} (tools/sanitize.py: inference + call eval + call training-forward + init + 2 train_steps + 1 train_step with the fused training forward / two-CTAs-per-SM GEMMs forced + mel inversion, B4 T_text 20 T_mel 70)" > $OUT
for tool in memcheck synccheck racecheck; do
  echo "" >> $OUT; echo "## --tool $tool" >> $OUT
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 | grep -E "COMPUTE-SANITIZER|^ok|ERROR SUMMARY|RACECHECK SUMMARY|=========.*(error|hazard|Invalid|Race)" | head -40 >> $OUT
done
cat $OUT
