cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_exhaustive.py tests/test_dropin.py -m gpu -x -q > gpurun_out/r02_gputest_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_l.log
tail -4 gpurun_out/r02_gputest_l.log
timeout 300 python tools/probe.py c3_human 0 10 > gpurun_out/r02_probe_k.json 2> gpurun_out/r02_probe_k.err; echo "probe rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_probe_k.json')); print(d['ms_per_step'], d['stats'])"
