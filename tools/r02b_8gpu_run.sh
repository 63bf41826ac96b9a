# second session of round 2: 8-GPU confirmation with the final library (C3 on CTA pairs, C2) under torchrun
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 "$@" > gpurun_out/r02b_$tag.json 2> gpurun_out/r02b_$tag.err; python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r02b_$tag.json").read().splitlines()[-1])
    print("$tag", "value", round(l["value"],1), "ms", round(l["ms_per_step"],2), "impl", l.get("tile_impl"), "e2e", l.get("e2e") and round(l["e2e"]["value"],1), "parity", l.get("parity_vs_n1") and (l["parity_vs_n1"]["max_rel_err"], l["parity_vs_n1"]["alpha_equal_across_ranks"]), "rebal", l.get("tile_share_rebalances"), l["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/r02b_$tag.err").read()[-800:])
PY
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run C3_n8_torchrun $TR --nproc-per-node 8 --master-port 29572 bench.py --gpus 8 --workload C3 --steps 30 --warmup 3
run C2_n8_torchrun $TR --nproc-per-node 8 --master-port 29573 bench.py --gpus 8 --steps 40 --warmup 3
