# round 2: two ranks, 8 batches of 128 per rank (12 CPUs each) against 6 of 171
mkdir -p gpurun_out
for v in b8 b6; do
  if [ $v = b8 ]; then B=8; else B=6; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 --batches $B --no-cpu-baseline --no-single-stream > gpurun_out/r02af_$v.json 2> gpurun_out/r02af_$v.err
  echo "$v rc=$?"
done
python - <<'P'
import json
for n in ('b8','b6'):
    try:
        d=json.loads(open(f'gpurun_out/r02af_{n}.json').read().strip().splitlines()[-1])
        print(n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2),'batches',d['config']['batches_per_gpu'],'threads',d['config'].get('host_threads'),'cpu',d.get('host_cpu',{}).get('value_pass'))
    except Exception as e:
        print(n,'FAILED',e); print(open(f'gpurun_out/r02af_{n}.err').read()[-1500:])
P
