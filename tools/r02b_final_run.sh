# round 2, second session: the driver's sequence on one B200 with the final library + ncu evidence of the final kernels (C2: single-CTA fp64, C3: CTA-pair fp32)
mkdir -p gpurun_out
( nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/r02b_clocks.csv & echo $! > /tmp/smi.pid )
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -40 > gpurun_out/r02b_tests_final.log
t1=$(date +%s); echo "pytest seconds $((t1-t0))" >> gpurun_out/r02b_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke_final.log 2>&1
t2=$(date +%s)
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err
t3=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err
t4=$(date +%s); echo "reference arm seconds $((t3-t2)), default bench seconds $((t4-t3))" >> gpurun_out/r02b_tests_final.log
kill $(cat /tmp/smi.pid)
M="sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/ncu_launches_C2_r02b.csv python bench.py --steps 3 --warmup 1 --no-extra --no-e2e --no-cpu-baseline --no-dmma-line > gpurun_out/ncu_launches_C2_r02b.stdout 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/ncu_launches_C3_r02b.csv python bench.py --workload C3 --steps 3 --warmup 1 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_C3_r02b.stdout 2>&1
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:tile_kernel_i8 -s 1 -c 1 -f -o gpurun_out/prof_i8_C2_r02b python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu-baseline --no-dmma-line > gpurun_out/prof_i8_C2_r02b.stdout 2>&1
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:tile_kernel_i8 -s 1 -c 1 -f -o gpurun_out/prof_i8_C3_r02b python bench.py --workload C3 --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/prof_i8_C3_r02b.stdout 2>&1
tail -6 gpurun_out/r02b_tests_final.log; tail -4 gpurun_out/r02b_smoke_final.log; ls -la gpurun_out/*r02b*.ncu-rep; python - <<'PY'
import json
for f in ("r02b_bench_reference", "r02b_bench_default"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().splitlines()[-1])
        print(f, l.get("value"), l.get("e2e", {}).get("value"), (l.get("e2e_csvm") or {}).get("value"), (l.get("roofline") or {}).get("frac"), l.get("clocks"), [(e.get("workload", "?")[:2], round(e.get("value", 0), 2), (e.get("roofline") or {}).get("frac_sustained")) for e in (l.get("extra_workloads") or [])])
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-600:])
PY
