#!/bin/bash
# one box with N GPUs: concurrent PCIe ceiling, then the bench end-to-end at N ranks in three host-side variants
N=${1:-8}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/r2_topo_${N}.txt 2>&1; nproc >> $OUT/r2_topo_${N}.txt; numactl -H >> $OUT/r2_topo_${N}.txt 2>&1
timeout 300 $TR --master-port 29511 scripts/pcie_probe_multi.py > $OUT/r2_pcie_probe_${N}gpu.json 2> $OUT/r2_pcie_probe_${N}gpu.err; cat $OUT/r2_pcie_probe_${N}gpu.json
timeout 300 python scripts/pcie_probe_multi.py > $OUT/r2_pcie_probe_1gpu.json 2>/dev/null; cat $OUT/r2_pcie_probe_1gpu.json
i=0
for env in "X=1" "OTT_WC=1" "OTT_PIN=1"; do
  for wlk in cfg4 cfg2; do
    i=$((i+1))
    env $env timeout 400 $TR --master-port $((29520+i)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-extra --workload $wlk > $OUT/r2_scale${N}_${wlk}_${env%%=*}.json 2> $OUT/r2_scale${N}_${wlk}_${env%%=*}.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_scale${N}_${wlk}_${env%%=*}.json"))
    print("$env $wlk: value %.0f  e2e_dev %.0f (%.2f ms)  e2e_host %.0f (%.2f ms)" % (d["value"], d["e2e_device_out"]["value"], d["e2e_device_out"]["ms_per_step"], d["e2e_host_out"]["value"], d["e2e_host_out"]["ms_per_step"]))
except Exception as e:
    print("$env $wlk failed", e); print(open("$OUT/r2_scale${N}_${wlk}_${env%%=*}.err").read()[-800:])
PY
  done
done
