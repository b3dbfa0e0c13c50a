#!/bin/bash
# e2e legs of bench.py at N GPUs (torchrun, one process per GPU): default feeding and the narrowing-off variant.
# usage: tools/e2e_n.sh TAG N
TAG=${1:-n2}; N=${2:-2}
mkdir -p gpurun_out
nproc > gpurun_out/e2e_n_$TAG.log; lscpu | grep -i "model name\|^CPU(s)\|NUMA\|L3" >> gpurun_out/e2e_n_$TAG.log
nvidia-smi topo -m >> gpurun_out/e2e_n_$TAG.log 2>&1
for mode in 1; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --pairs 512 --steps 3 --warmup 3 --no-cpu-baseline --no-latency --no-inlib > gpurun_out/e2e_n_${TAG}_$mode.json 2> gpurun_out/e2e_n_${TAG}_$mode.err
  python - <<PY | tee -a gpurun_out/e2e_n_$TAG.log
import json
try:
    j = json.loads(open("gpurun_out/e2e_n_${TAG}_$mode.json").read().strip().splitlines()[-1])
    def f(e): return "%.0f pairs/s total (host %.1f GB/s per GPU, link %.1f GB/s per GPU, rank-0 doubles %d / narrowed %d)" % (e["value"], e["host_gbs_per_gpu"], e["link_gbs_per_gpu"], e["pairs_fed_as_doubles"], e["pairs_narrowed_on_host"])
    print("N=$N headline-narrowing=$mode: value %.0f e2e" % j["value"], f(j["e2e"]))
    for k, v in j["e2e_variants"].items(): print("   ", k, f(v))
except Exception as ex:
    print("N=$N mode $mode: FAILED", ex)
PY
done
