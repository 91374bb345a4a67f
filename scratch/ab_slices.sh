# A/B of the search's time slices on the bench workload.  usage: ab_slices.sh "slice:lanes" ...   (slice "-" = default policy)
for spec in "$@"; do
  sl=${spec%%:*}; ln=${spec#*:}
  if [ "$sl" = "-" ]; then unset LQB_SEEK_SLICE; else export LQB_SEEK_SLICE=$sl; fi
  timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --lanes $ln --no-workloads --no-e2e --no-cpu-baseline 2>gpurun_out/ab_err.txt | python -c "
import sys,json
t=sys.stdin.read().strip().splitlines()
if not t: print('$spec', 'NO OUTPUT'); sys.exit(0)
d=json.loads(t[-1])
print('$spec', 'step %.2f ms' % d['ms_per_step'], 'value %.0f' % d['value'], [(k['name'], round(k['ms_per_step'],2)) for k in d['kernels']], d['frames_found_per_step'], d['frames_valid_per_step'], d['gpu_launches'])"
done
