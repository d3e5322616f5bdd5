# CPU micro-benchmark of the host state machine (tests/cpp/host_perf.cpp); no GPU needed
set -e
g++ -std=c++17 -O3 -march=x86-64-v3 -I /usr/local/cuda/include tests/cpp/host_perf.cpp -o /tmp/host_perf
python - <<'P'
import json
from xivo_b200 import sim
cfg = sim.load_cfg("xivo_b200/cfg/pcw_sim.json")
cfg["tracker_cfg"].update(num_features_min=120, num_features_max=150)
open("/tmp/host_perf_cfg.json", "w").write(json.dumps(cfg))
P
/tmp/host_perf /tmp/host_perf_cfg.json ${1:-4} ${2:-14} ${3:-600} ${4:-1}
