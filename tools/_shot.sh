cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NEW16="GTSB_HUBS=1 GTSB_HUB_SORT=1 GTSB_SMALL_MAX=16"
export NEW32="GTSB_HUBS=1 GTSB_HUB_SORT=1 GTSB_SMALL_MAX=32"
stamp() { echo "$1 $(date +%s.%N)" >> gpurun_out/shot_times.log; }
stamp start
# 1. hub parity with the new kernels: config-4 shape against the compiled reference on both build paths, tiny adversarial graphs
env $NEW16 timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "c4_repeat_hubs-150000 or fallback_reasons or stats_report or tiny_adversarial" > gpurun_out/shot_parity_hubs.log 2>&1; echo "rc=$?" >> gpurun_out/shot_parity_hubs.log
stamp parity_hubs
# 2. config 4 at full size: new kernels, then the kernels as they were (digests must agree)
env $NEW16 timeout 25 python tools/probe.py c4_repeat_hubs 0 5 > gpurun_out/shot_c4_new16.json 2> gpurun_out/shot_c4_new16.err; echo "rc=$?" >> gpurun_out/shot_c4_new16.err
stamp c4_new16
timeout 25 python tools/probe.py c4_repeat_hubs 0 5 > gpurun_out/shot_c4_old.json 2> gpurun_out/shot_c4_old.err; echo "rc=$?" >> gpurun_out/shot_c4_old.err
stamp c4_old
# 3. config 3: L2 fetch granularity 32 B
GTSB_L2_FETCH=32 GTSB_TRACE=1 timeout 25 python tools/probe.py c3_human 0 10 > gpurun_out/shot_c3_fetch32.json 2> gpurun_out/shot_c3_fetch32.err; echo "rc=$?" >> gpurun_out/shot_c3_fetch32.err
stamp c3_fetch32
# 4. the quick parity set with every new switch on
env $NEW16 timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "not full_size and not named and not tiny_adversarial" > gpurun_out/shot_parity_quick.log 2>&1; echo "rc=$?" >> gpurun_out/shot_parity_quick.log
stamp parity_quick
# 5. the other bound of the thread-per-bucket path; config 3 as it is on this box
env $NEW32 timeout 25 python tools/probe.py c4_repeat_hubs 0 5 > gpurun_out/shot_c4_new32.json 2> gpurun_out/shot_c4_new32.err; echo "rc=$?" >> gpurun_out/shot_c4_new32.err
stamp c4_new32
GTSB_TRACE=1 timeout 25 python tools/probe.py c3_human 0 10 > gpurun_out/shot_c3_default.json 2> gpurun_out/shot_c3_default.err; echo "rc=$?" >> gpurun_out/shot_c3_default.err
stamp c3_default
