# very last pass: default library (CL = 1 path touched by the CL = 2 generalisation): GPU tests + default bench; experimental library: the cluster / pair variant tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -4 > gpurun_out/r02b_tests_last2.log; tail -2 gpurun_out/r02b_tests_last2.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02b_bench_default_last2.json 2> gpurun_out/r02b_bench_default_last2.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r02b_bench_default_last2.json").read().splitlines()[-1])
    print("bench", round(l["value"], 2), round(l["e2e"]["value"], 2), (l.get("e2e_csvm") or {}).get("value"), l["roofline"]["frac"], l["clocks"], [(e.get("workload", "?")[:2], round(e.get("value", 0), 3), (e.get("roofline") or {}).get("frac_sustained")) for e in (l.get("extra_workloads") or [])])
except Exception as e:
    print("bench FAILED", e); print(open("gpurun_out/r02b_bench_default_last2.err").read()[-600:])
PY
PLSSVM_B200_LIB=$PWD/_ab/lib_exp.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster or pair or variant" 2>&1 | tail -2 | tee gpurun_out/r02b_tests_experimental2.log
