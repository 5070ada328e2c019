#!/bin/bash
# round-2 GPU call E (8 GPUs): copy-engine gather vs NCCL at N=8, bench at N=8 and N=4 with the full-size
# BASELINE configs 3/4/5 split over the ranks, and the end-to-end (host entry) scaling diagnosis.
O=gpurun_out/r02e; mkdir -p $O
export MWRT_QUIET=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python tools/prof_target2.py c2rt 10000 8 > $O/time_c2rt.log 2>&1; cat $O/time_c2rt.log
nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> $O/topo.txt
timeout 240 $TR --nproc-per-node 8 --master-port 29533 tools/check_ce_gather.py > $O/check_ce_n8.log 2>&1; echo "check_ce rc=$?"; tail -3 $O/check_ce_n8.log
timeout 400 $TR --nproc-per-node 8 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8_ce.json 2> $O/bench_n8_ce.err; echo "bench n8 ce rc=$?"
timeout 240 $TR --nproc-per-node 8 --master-port 29535 bench.py --gpus 8 --steps 20 --warmup 3 --gather nccl --no-extras > $O/bench_n8_nccl.json 2> $O/bench_n8_nccl.err; echo "bench n8 nccl rc=$?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 $TR --nproc-per-node 4 --master-port 29536 bench.py --gpus 4 --steps 20 --warmup 3 > $O/bench_n4_ce.json 2> $O/bench_n4_ce.err; echo "bench n4 ce rc=$?"
MWRT_NO_BIND=1 timeout 200 $TR --nproc-per-node 8 --master-port 29537 tools/e2e_diag.py > $O/e2e_diag_n8_nobind.json 2> $O/e2e_diag_n8_nobind.err
timeout 200 $TR --nproc-per-node 8 --master-port 29538 tools/e2e_diag.py > $O/e2e_diag_n8_bind.json 2> $O/e2e_diag_n8_bind.err
MWRT_NO_BIND=1 timeout 100 python tools/e2e_diag.py > $O/e2e_diag_n1_nobind.json 2> $O/e2e_diag_n1.err
cat $O/e2e_diag_*.json
for f in $O/*.err; do echo "== $f"; tail -n 2 $f; done
ls -la $O
