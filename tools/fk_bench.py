"""FK / rot6d kernel sweep only (BASELINE config 3): prints bench.py's `fk` object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
sizes = tuple(int(a) for a in sys.argv[1:]) or (43, 683, 10923, 174763, 699051)
print(json.dumps(bench.fk_sweep(dev, bench.peaks(), sizes=sizes)))
