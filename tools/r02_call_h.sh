#!/bin/bash
# round-2 GPU call H (N GPUs = $1): push gather with the entry barrier, skewed ranks, several CTA counts.
N=${1:-2}
O=gpurun_out/r02h_n$N; mkdir -p $O
export MWRT_QUIET=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
port=29580
CFGS=${CFGS:-"multicast:8 multicast:16 multicast:32 multicast:4 peers:8 ce:8"}
for cfg in $CFGS; do
  set -- ${cfg/:/ }
  MWRT_PUSH_BLOCKS=$2 timeout 200 $TR --master-port $port tools/check_ce_gather.py $1 > $O/check_$1_$2.log 2>&1; echo "check $1 $2 CTAs rc=$?"
  grep -E "identical|ms/step" $O/check_$1_$2.log; port=$((port+1))
done
timeout 300 $TR --master-port 29590 bench.py --gpus $N --steps 20 --warmup 3 --no-extras > $O/bench_push.json 2> $O/bench_push.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$O/bench_push.json"))
print("bench N=$N value %.4e ms/step %.4f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
