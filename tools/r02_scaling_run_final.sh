# final 8-GPU pass with the final library: N = 1 / 8 under torchrun, the one-process device group (incl. e2e through the reference's csvm::fit),
# C5 under torchrun (sharded predict + exchange of the value ranges), multi-GPU tests
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 "$@" > gpurun_out/r02_final_$tag.json 2> gpurun_out/r02_final_$tag.err; python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r02_final_$tag.json").read().splitlines()[-1])
    print("$tag", "value", round(l["value"],1), "ms", round(l["ms_per_step"],2), "e2e", l.get("e2e") and round(l["e2e"]["value"],1), "csvm", l.get("e2e_csvm") and l["e2e_csvm"].get("value"), "parity", l.get("parity_vs_n1") and (l["parity_vs_n1"]["max_rel_err"], l["parity_vs_n1"]["alpha_equal_across_ranks"]), "rebal", l.get("tile_share_rebalances"), l["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/r02_final_$tag.err").read()[-800:])
PY
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run n1 python bench.py --gpus 1 --steps 40 --warmup 3 --no-extra --no-cpu-baseline --no-dmma-line
run n8_torchrun $TR --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --steps 40 --warmup 3
run n8_group python bench.py --gpus 8 --steps 40 --warmup 3
run n8_torchrun_C5 $TR --nproc-per-node 8 --master-port 29553 bench.py --gpus 8 --workload C5 --steps 2 --warmup 1
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py -m gpu -q --timeout 600 2>&1 | tail -8 | tee gpurun_out/r02_tests_8gpu_final.log
