# round 2: two ranks on one box exactly as the driver launches them (NUMA slice per rank, pool with the availability counter)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02ae_topo.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_bench_workload.py tests/test_gpu_estimator.py -x -q -m gpu -k "workload or step_call or prefetch or batch" > gpurun_out/r02ae_pytest.txt 2>&1; tail -2 gpurun_out/r02ae_pytest.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/r02ae_n2.json 2> gpurun_out/r02ae_n2.err
echo "n2 rc=$?"
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline --no-single-stream > gpurun_out/r02ae_n1.json 2> gpurun_out/r02ae_n1.err
echo "n1 rc=$?"
python - <<'P'
import json
for n in ('n2','n1'):
    try:
        d=json.loads(open(f'gpurun_out/r02ae_{n}.json').read().strip().splitlines()[-1])
        print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2),'batches',d['config']['batches_per_gpu'],'threads',d['config'].get('host_threads'),'numa',d['config'].get('numa'),'cpu',d.get('host_cpu',{}).get('value_pass'))
    except Exception as e:
        print(n,'FAILED',e); print(open(f'gpurun_out/r02ae_{n}.err').read()[-1500:])
P
