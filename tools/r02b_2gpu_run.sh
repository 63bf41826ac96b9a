# second session of round 2: the new fp32 default (CTA-pair kernel, tile ranges in super-tiles) on 2 GPUs — torchrun, one-process device group, multi-GPU tests
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 400 "$@" > gpurun_out/r02b_$tag.json 2> gpurun_out/r02b_$tag.err; python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r02b_$tag.json").read().splitlines()[-1])
    print("$tag", "value", round(l["value"],1), "ms", round(l["ms_per_step"],2), "impl", l.get("tile_impl"), "e2e", l.get("e2e") and round(l["e2e"]["value"],1), "parity", l.get("parity_vs_n1") and (l["parity_vs_n1"]["max_rel_err"], l["parity_vs_n1"]["alpha_equal_across_ranks"]), "rebal", l.get("tile_share_rebalances"), l["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/r02b_$tag.err").read()[-800:])
PY
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run C3_n1 python bench.py --gpus 1 --workload C3 --steps 20 --warmup 3 --no-extra --no-cpu-baseline
run C3_n2_torchrun $TR --nproc-per-node 2 --master-port 29562 bench.py --gpus 2 --workload C3 --steps 20 --warmup 3
run C3_n2_group python bench.py --gpus 2 --workload C3 --steps 20 --warmup 3
run C2_n2_torchrun $TR --nproc-per-node 2 --master-port 29563 bench.py --gpus 2 --steps 20 --warmup 3
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py -m gpu -q --timeout 500 2>&1 | tail -6 | tee gpurun_out/r02b_tests_2gpu.log
