# round 2, GPU call 6: which TMA start coordinates fault (negative vs not 16-byte aligned), PCIe / copy-engine probe, host-scope profiles of the pool mode
mkdir -p gpurun_out
for v in 15 16 17 18 19 20 21 22 23 24; do timeout 60 scripts/probes/tma_probe $v; done 2>&1 | tee gpurun_out/r02f_tma_probe.txt
timeout 120 scripts/probes/pcie_probe 2>&1 | tee gpurun_out/r02f_pcie_probe.txt
export XIVO_LANES=1 XIVO_ZC_TABLES=0
timeout 300 python bench.py --steps 6 --warmup 3 --batches 8 --no-cpu-baseline --no-single-stream --profile-level 2 --profile-overlapped > gpurun_out/r02f_pool8_dev.json 2> gpurun_out/r02f_pool8_dev.err
timeout 300 python bench.py --steps 6 --warmup 3 --batches 8 --no-cpu-baseline --no-single-stream --profile-level 2 --profile-overlapped --profile-e2e > gpurun_out/r02f_pool8_e2e.json 2> gpurun_out/r02f_pool8_e2e.err
python - <<'P'
import json
for n in ("dev", "e2e"):
    try:
        d = json.loads(open(f"gpurun_out/r02f_pool8_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "prof ms/step", d["roofline"]["profiled_pass_ms_per_step"])
        hp = d["host_phase_ms_per_step"]
        print(" batch-level ms/step:", {k: v for k, v in hp.items() if not k.startswith("x_")})
        print(" per-seq-frame us:", {k: round(v * 1000 / 4096, 2) for k, v in hp.items() if k.startswith("x_")})
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r02f_pool8_{n}.err").read()[-600:])
P
