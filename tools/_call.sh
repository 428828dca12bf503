cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -m gpu -x -q -k "proposals_than or dropin or binding or config1 or digest_equals" > gpurun_out/r02_gputest_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_h.log
tail -4 gpurun_out/r02_gputest_h.log
for v in "GTSB_NB=128" "GTSB_NB=256" "GTSB_MAIL=3 GTSB_NB=128" "GTSB_MAIL=3 GTSB_NB=256"; do
  tag=$(echo $v | tr '= ' '__')
  env $v timeout 300 python tools/probe.py c3_human 0 10 > gpurun_out/r02_probe_h_$tag.json 2> gpurun_out/r02_probe_h_$tag.err
  echo "probe $v rc=$?"
done
timeout 900 python tools/c5_check.py --steps 3 --out gpurun_out/r02_c5_n1.json > gpurun_out/c5d.log 2>&1; echo "c5 n1 rc=$?"
tail -2 gpurun_out/c5d.log | cut -c1-300
python tools/c5_check.py --compare gpurun_out/r02_c5_n1.json gpurun_out/r02_c5_n2.json | tee gpurun_out/r02_c5_n1_vs_n2.json
