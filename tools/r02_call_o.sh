#!/bin/bash
# round-2 GPU call O (N GPUs = $1): BASELINE configs 3/5/4 with the piece-wise push gather vs NCCL all-gather
N=${1:-2}
O=gpurun_out/r02o_n$N; mkdir -p $O
export MWRT_QUIET=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
port=29600
for cfg in 5 3 4; do
  for g in push nccl; do
    timeout 300 $TR --master-port $port bench.py --gpus $N --config $cfg --gather $g --steps 5 > $O/cfg${cfg}_$g.json 2> $O/cfg${cfg}_$g.err; echo "config $cfg $g rc=$?"
    port=$((port+1))
  done
done
python - <<PY
import json
for cfg in (5, 3, 4):
    for g in ("push", "nccl"):
        try:
            d = json.load(open("$O/cfg%d_%s.json" % (cfg, g)))["detail"]
            print("config %d %-4s N=%d: %.3f ms per pass (kernels %.3f), identical to NCCL: %s, finite %s | %s" % (cfg, g, d["n_gpus"], d["ms_per_pass"], d["ms_kernels_only"], d["gather_identical_to_nccl"], d["finite"], d["gathered"][:70]))
        except Exception as e:
            print("config", cfg, g, "unreadable:", e)
PY
for f in $O/*.err; do echo "== $f"; tail -n 3 $f; done
