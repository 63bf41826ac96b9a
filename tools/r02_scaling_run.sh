# 8-GPU box: strong scaling of C2 at 1 / 4 / 8 GPUs (one process per GPU under torchrun, the driver's launch), the same through ONE process
# driving a device group behind the C ABI, fixed vs rate-weighted tile shares, C5 with in-library predict sharding, multi-GPU tests.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_box_8gpu.txt
run() { tag=$1; shift; timeout 600 "$@" > gpurun_out/r02_scale_$tag.json 2> gpurun_out/r02_scale_$tag.err; python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r02_scale_$tag.json").read().splitlines()[-1])
    print("$tag", "value", round(l["value"],1), "ms", round(l["ms_per_step"],2), "e2e", l.get("e2e") and round(l["e2e"]["value"],1), "csvm", l.get("e2e_csvm") and l["e2e_csvm"].get("value"), "parity", l.get("parity_vs_n1") and (l["parity_vs_n1"]["max_rel_err"], l["parity_vs_n1"]["alpha_equal_across_ranks"]), "rebal", l.get("tile_share_rebalances"), "tile_ms", l["roofline"].get("avg_launch_ms"), l["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/r02_scale_$tag.err").read()[-800:])
PY
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run n1 python bench.py --gpus 1 --steps 40 --warmup 3 --no-extra --no-cpu-baseline --no-dmma-line
run n4_torchrun $TR --nproc-per-node 4 --master-port 29541 bench.py --gpus 4 --steps 40 --warmup 3
run n8_torchrun $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 40 --warmup 3
run n8_torchrun_fixed_shares $TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --steps 40 --warmup 3 --balance 0 --no-e2e
run n8_group python bench.py --gpus 8 --steps 40 --warmup 3
run n8_group_C5 python bench.py --gpus 8 --workload C5 --steps 2 --warmup 1
run n8_torchrun_C3 $TR --nproc-per-node 8 --master-port 29544 bench.py --gpus 8 --workload C3 --steps 40 --warmup 3
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py -m gpu -q --timeout 600 2>&1 | tail -15 | tee gpurun_out/r02_tests_8gpu.log
