# last pass of the second session: GPU tests, smoke and the default bench with the final library; role counters of the final kernels (fine split: -DPB_TILE_STATS_FINE build)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -6 > gpurun_out/r02b_tests_last.log; tail -3 gpurun_out/r02b_tests_last.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke_last.log 2>&1; tail -3 gpurun_out/r02b_smoke_last.log
timeout 600 python bench.py > gpurun_out/r02b_bench_default_last.json 2> gpurun_out/r02b_bench_default_last.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r02b_bench_default_last.json").read().splitlines()[-1])
    print("bench", round(l["value"], 2), round(l["e2e"]["value"], 2), (l.get("e2e_csvm") or {}).get("value"), l["roofline"]["frac"], l["clocks"], [(e.get("workload", "?")[:2], round(e.get("value", 0), 3), e.get("repeat_call_seconds"), (e.get("roofline") or {}).get("frac_sustained")) for e in (l.get("extra_workloads") or [])])
except Exception as e:
    print("bench FAILED", e); print(open("gpurun_out/r02b_bench_default_last.err").read()[-600:])
PY
PLSSVM_B200_LIB=$PWD/_ab/lib_fine.so timeout 200 python tools/check_pair.py --skip-parity --stats > gpurun_out/r02b_role_stats.log 2>&1; grep -E "tile_stats|\"impl\"" gpurun_out/r02b_role_stats.log | cut -c1-400
