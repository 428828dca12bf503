cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_e.log
tail -5 gpurun_out/r02_gputest_e.log
for v in "new=1" "GTSB_MAIL=0" "GTSB_PAIRS=1" "GTSB_PAIRS_OCC=6"; do
  tag=$(echo $v | tr '=' '_')
  env $v timeout 300 python tools/probe.py c3_human 0 10 > gpurun_out/r02_probe_f_$tag.json 2> gpurun_out/r02_probe_f_$tag.err
  echo "probe $v rc=$?"
done
