cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_final.log
tail -4 gpurun_out/r02_gputest_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "bench rc=$?"
timeout 600 python bench.py --workload c4_repeat_hubs --steps 5 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; echo "bench c4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/probe.py c3_human 0 1 > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k4_pairs|k2_partition2|k2_deliver2|k2_resolve|k4_finalize|k4_fire_dense|k2_classify|k_fire_rounds_all|k4_fire_redo|k4_fire_init|k4_vertex_facts" -c 11 -f -o gpurun_out/prof_r02_final python tools/probe.py c3_human 0 1 > gpurun_out/ncu_f.log 2>&1; echo "ncu full rc=$?"
