# A/B of liblqb200 variants on the bench workload: prints search / step times per library.  usage: ab_seek.sh "lib[:lanes[:ENV=V]]" ...
for spec in "$@"; do
  lib=${spec%%:*}; rest=${spec#*:}; lanes=0; envs=""
  if [ "$rest" != "$spec" ]; then lanes=${rest%%:*}; e=${rest#*:}; [ "$e" != "$rest" ] && envs=$e; fi
  env $envs LQB_LIB=gr-liquiddsp_b200/lib/$lib timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --lanes $lanes --no-workloads --no-e2e --no-cpu-baseline 2>gpurun_out/ab_err.txt | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$spec', 'step %.2f ms' % d['ms_per_step'], [(k['name'], round(k['ms_per_step'],2)) for k in d['kernels']], d['frames_found_per_step'], d['frames_valid_per_step'])"
done
