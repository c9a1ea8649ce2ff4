"""Generates tests/golden/cfg4_eval.npz: the UNMODIFIED reference IST_Net (oracle/ref_harness.py, CPU) on a dense 4096-point
crop (BASELINE.json configs[4] shape for the point branch; 64x64 RGB keeps the fixture small), eval forward with non-trivial
BatchNorm running statistics / affine parameters.  Run in the build container only:  python tests/gen_golden_cfg4.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import perturb_batchnorm, sd_checksum  # noqa: E402
from istnet_b200 import model as M  # noqa: E402
from istnet_b200.synth import make_batch  # noqa: E402
from oracle import ref_harness  # noqa: E402

ns = ref_harness.load()
torch.set_num_threads(8)
torch.manual_seed(1)
mine = M.IST_Net(6, False)
perturb_batchnorm(mine, seed=41)
ref = ns.ist_net.IST_Net(6, False)
ref.load_state_dict(mine.state_dict())
ref.eval()
inp = make_batch(1, 4096, 64, seed=14, quantize=True)
with torch.no_grad():
    ep = ref({k: v.clone() for k, v in inp.items()})
np.savez_compressed(
    os.path.join(ROOT, "tests", "golden", "cfg4_eval.npz"),
    sd_checksum=sd_checksum(mine.state_dict()),
    **{"in_" + k: v.numpy() for k, v in inp.items()},
    **{"out_" + k: v.numpy() for k, v in ep.items()},
)
print({k: tuple(v.shape) for k, v in ep.items()})
