timeout -s KILL 300 python -m pytest tests/test_gpu_tx_det.py tests/test_blocks.py tests/test_python_layer.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout -s KILL 200 python bench.py --workload tx --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/e1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tx', round(d['ms_per_step'],3), d['kernel_ms_per_step'], d['roofline']['frac'])"
timeout -s KILL 200 python bench.py --workload tx_rx_per --steps 5 --warmup 2 2>gpurun_out/e2.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('per', d['value'], d['ms_per_step'], d['ms_per_step_parts'], d['roofline']['frac'], d['failed'])"
tail -c 300 gpurun_out/e1.err gpurun_out/e2.err
