#!/bin/bash
# round-2 GPU call G (8 GPUs): the multicast push gather at N=8 (4 / 8 / 16 CTAs) vs NCCL, bench at N=8 and N=4.
O=gpurun_out/r02g; mkdir -p $O
export MWRT_QUIET=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29560
for nb in 8 4 16; do
  MWRT_PUSH_BLOCKS=$nb timeout 200 $TR --nproc-per-node 8 --master-port $port tools/check_ce_gather.py multicast > $O/check_push_n8_mc$nb.log 2>&1; echo "check mc$nb rc=$?"; tail -n 3 $O/check_push_n8_mc$nb.log; port=$((port+1))
done
timeout 300 $TR --nproc-per-node 8 --master-port 29570 bench.py --gpus 8 --steps 20 --warmup 3 --no-extras > $O/bench_n8_push.json 2> $O/bench_n8_push.err; echo "bench n8 rc=$?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 $TR --nproc-per-node 4 --master-port 29571 bench.py --gpus 4 --steps 20 --warmup 3 --no-extras > $O/bench_n4_push.json 2> $O/bench_n4_push.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
for f in ("bench_n8_push", "bench_n4_push"):
    try:
        d = json.load(open("gpurun_out/r02g/%s.json" % f))
        print(f, "value %.4e ms/step %.4f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["config"]["collective"][:60])
    except Exception as e:
        print(f, "unreadable", e)
PY
for f in $O/*.err; do echo "== $f"; tail -n 2 $f; done
