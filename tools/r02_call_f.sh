#!/bin/bash
# round-2 GPU call F (2 GPUs): push gather (multicast kernel / peer stores / copy engines) vs NCCL at N=2,
# GPU tests after the fp32-loop and ray-tracing changes.
O=gpurun_out/r02f; mkdir -p $O
export MWRT_QUIET=1
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
python tools/prof_target2.py c2rt 10000 8 > $O/time_c2rt.log 2>&1; python tools/prof_target2.py c2 10000 8 >> $O/time_c2rt.log 2>&1; python tools/prof_target2.py c2f32 10000 8 >> $O/time_c2rt.log 2>&1; cat $O/time_c2rt.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29540
for t in multicast peers ce; do
  timeout 300 $TR --master-port $port tools/check_ce_gather.py $t > $O/check_push_n2_$t.log 2>&1; echo "check $t rc=$?"; tail -n 3 $O/check_push_n2_$t.log; port=$((port+1))
done
timeout 600 $TR --master-port 29550 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > $O/bench_n2_push.json 2> $O/bench_n2_push.err; echo "bench push rc=$?"
tail -n 3 $O/bench_n2_push.err
ls -la $O
