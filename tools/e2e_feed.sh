#!/bin/bash
# Host-fed (e2e) legs of bench.py under different copy-thread counts / store kinds.
# usage: tools/e2e_feed.sh TAG "threads:nt:bothways ..."     e.g. "12:1:1 12:1:0 8:1:1"
TAG=${1:-feed}; CONFIGS=${2:-"12:1:1 12:1:0"}
mkdir -p gpurun_out
nproc > gpurun_out/e2e_feed_$TAG.log; lscpu | grep -i "model name\|^CPU(s)\|NUMA\|L3" >> gpurun_out/e2e_feed_$TAG.log
for c in $CONFIGS; do
  IFS=: read th nt bw dp bl <<< "$c"; bw=${bw:-1}; dp=${dp:-2}; bl=${bl:-1}
  AUDIOSYNC_CUDA_COPY_THREADS=$th AUDIOSYNC_CUDA_NT_STORES=$nt AUDIOSYNC_CUDA_FEED_BOTH_WAYS=$bw AUDIOSYNC_CUDA_FEED_DEPTH=$dp AUDIOSYNC_CUDA_FEED_BACKLOG=$bl E2E_ONLY_HEADLINE=${E2E_ONLY_HEADLINE:-0} timeout 600 python bench.py --pairs 256 --steps 3 --warmup 3 \
    --no-cpu-baseline --no-latency > gpurun_out/e2e_feed_${TAG}_${th}_${nt}_${bw}_${dp}_${bl}.json 2> gpurun_out/e2e_feed_${TAG}_${th}_${nt}_${bw}_${dp}_${bl}.err
  python - <<PY | tee -a gpurun_out/e2e_feed_$TAG.log
import json
try:
    j = json.loads(open("gpurun_out/e2e_feed_${TAG}_${th}_${nt}_${bw}_${dp}_${bl}.json").read().strip().splitlines()[-1])
    def f(e): return "%.0f pairs/s (host %.1f GB/s, link %.1f GB/s, doubles %d / narrowed %d)" % (e["value"], e["host_gbs_per_gpu"], e["link_gbs_per_gpu"], e["pairs_fed_as_doubles"], e["pairs_narrowed_on_host"])
    print("threads $th nt $nt both-ways $bw depth $dp backlog $bl: e2e", f(j["e2e"]))
    for k, v in j["e2e_variants"].items(): print("   ", k, f(v))
except Exception as ex:
    print("threads $th nt $nt both-ways $bw depth $dp backlog $bl: FAILED", ex)
PY
done
