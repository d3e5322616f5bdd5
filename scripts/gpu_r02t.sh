# round 2: 2-GPU replicas check exactly as the driver launches it, then the 1-GPU line on the same box
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02t_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02t_n2.json 2> gpurun_out/r02t_n2.err
echo "n2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02t_n1.json 2> gpurun_out/r02t_n1.err
echo "n1 rc=$?"
python - <<'P'
import json
for n in ('n2','n1'):
    try:
        d=json.loads(open(f'gpurun_out/r02t_{n}.json').read().strip().splitlines()[-1])
        print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2),'cpu',d.get('cpu_baseline',{}).get('value'))
    except Exception as e:
        print(n,'FAILED',e); print(open(f'gpurun_out/r02t_{n}.err').read()[-1500:])
P
