cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29541 tests/dist_check.py > gpurun_out/r02_dist_check_n8.log 2>&1; echo "dist_check n8 rc=$?"
grep -c -- "-> OK" gpurun_out/r02_dist_check_n8.log
timeout 900 $TR --nproc-per-node 8 --master-port 29542 tools/c5_check.py --steps 5 --out gpurun_out/r02_c5_n8.json > gpurun_out/c5e.log 2>&1; echo "c5 n8 rc=$?"
tail -1 gpurun_out/c5e.log | cut -c1-400
timeout 900 $TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8_a.json 2> gpurun_out/r02_bench_n8_a.err; echo "bench n8 rc=$?"
cut -c1-300 gpurun_out/r02_bench_n8_a.json
