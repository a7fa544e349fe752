"""Two steady-state steps of the full ML+2PN pipeline (QWS shape, 18,944 instances) for an ncu launch list of the ML stage
(Net.score_requests) and the candidate selection: `ncu --metrics gpu__time_duration.sum python scripts/pipeline_launches.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda")
shape = sys.argv[1] if len(sys.argv) > 1 else "qws"
print(bench.pipeline_block(dev, shape, 18944, 1, 1)["stage_ms"])
