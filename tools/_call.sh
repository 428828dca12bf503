cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_final2.log
tail -4 gpurun_out/r02_gputest_final2.log
timeout 600 python bench.py > gpurun_out/r02_bench_c3_final.json 2> gpurun_out/r02_bench_c3_final.err; echo "bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
