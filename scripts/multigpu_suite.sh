#!/bin/bash
# one gpurun --gpus N call: config 5 (T sweep, strong scaling), config 4 (training step, pitch-sharded), bench weak/strong
# with the fused and the copy-engine exchange.  usage: scripts/multigpu_suite.sh N   (writes gpurun_out/mg_N_*.json)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
[ -n "$SKIP_C5" ] || timeout 300 $TR --master-port 29601 scripts/config5_sweep.py --steps 20 2>gpurun_out/mg_${N}_c5.err | tail -1 > gpurun_out/mg_${N}_config5.json
timeout 300 $TR --master-port 29602 scripts/config4_train.py --steps 30 2>gpurun_out/mg_${N}_c4.err | tail -1 > gpurun_out/mg_${N}_config4.json
for mode in ${MODES:-fused push}; do
  TKB_GATHER=$mode timeout 300 $TR --master-port 29603 bench.py --gpus $N --steps 50 --warmup 5 2>gpurun_out/mg_${N}_weak_$mode.err | tail -1 > gpurun_out/mg_${N}_weak_$mode.json
done
[ -n "$SKIP_STRONG" ] || TKB_GATHER=push timeout 300 $TR --master-port 29604 bench.py --gpus $N --steps 50 --warmup 5 --scaling strong 2>gpurun_out/mg_${N}_strong.err | tail -1 > gpurun_out/mg_${N}_strong.json
python - <<PY
import json
for name in ("config5","config4","weak_fused","weak_push","strong"):
    try:
        d=json.load(open("gpurun_out/mg_${N}_%s.json"%name))
        if name=="config5": print(name, [(r["T"], "%.3g"%r["cells_per_s"], "%.2f"%r["frac_of_hbm_peak"]) for r in d["rows"]])
        elif name=="config4": print(name, d["ms_per_step"], d["parts_ms"])
        else: print(name, "%.4g"%d["value"], d["ms_per_step"], "e2e %.3g"%d["e2e"]["value"])
    except Exception as e: print(name, "FAILED", e)
PY
