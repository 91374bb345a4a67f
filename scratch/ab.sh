for v in liblqb200 v_b0; do
  LQB_LIB=gr-liquiddsp_b200/lib/$v.so timeout -s KILL 200 python bench.py --steps 8 --warmup 3 --no-workloads --no-e2e --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  LQB_LIB=gr-liquiddsp_b200/lib/$v.so timeout -s KILL 200 python bench.py --workload detector --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/abd_$v.json 2> gpurun_out/abd_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_$v.json'))
print('$v', round(d['ms_per_step'],2), [round(k['ms_per_step'],2) for k in d['kernels']], d['frames_found_per_step'], d['frames_valid_per_step'], d['kernels'][0].get('cfo_bins_per_exact_window'))
d=json.load(open('gpurun_out/abd_$v.json'))
print('   det', round(d['ms_per_step'],2), d['detections_per_step'], d['search_work_last_step'])
PY
done
timeout -s KILL 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
