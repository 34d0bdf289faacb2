#!/bin/bash
# bench lines: N=1 full line, reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'P'
import json
l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
def show(d,ind=0):
    for k,v in d.items():
        if isinstance(v,dict) and k not in ('config','clocks'): print(' '*ind+k+':'); show(v,ind+2)
        elif k in ('value','ms_per_step','frac','achieved','launch_ms','all_gather_ms','error','trace','kernel_path','gpu_launches','env_step_kernel_only_steps_per_sec','lob_messages_per_sec') or k=='books_curve': print(' '*ind+f'{k}: {v}')
show(l)
P
