# second session of round 2: same-box A/B of the library before / after the epilogue refactor, then compute-sanitizer over the default library
mkdir -p gpurun_out
for v in old new old new; do
  L=$PWD/plssvm_b200/libplssvm_b200.so; [ $v = old ] && L=$PWD/_ab/lib_old.so
  PLSSVM_B200_LIB=$L timeout 200 python tools/check_pair.py --skip-parity --only=C3 > gpurun_out/ab_refactor_${v}_C3.log 2>&1; echo $v C3 rc=$?; grep -E "tflops|rror" gpurun_out/ab_refactor_${v}_C3.log | cut -c1-400
  PLSSVM_B200_LIB=$L timeout 200 python tools/check_pair.py --skip-parity --only=C2 > gpurun_out/ab_refactor_${v}_C2.log 2>&1; echo $v C2 rc=$?; grep -E "tflops|rror" gpurun_out/ab_refactor_${v}_C2.log | cut -c1-400
done
for tool in memcheck synccheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/r02b_sanitizer_$tool.log 2>&1; echo $tool rc=$?; tail -3 gpurun_out/r02b_sanitizer_$tool.log
done
