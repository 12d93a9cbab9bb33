"""Freeze the three reference sequences as compact input fixtures.

/root/reference does not exist on the GPU box, so the parsed problems
(indices, pixel measurements, initial parameters -- data, not source code) are
stored as compressed .npz next to the golden outputs.  Run in the authoring
container:  python tests/golden/make_inputs.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gbp_poplar_b200 import BALProblem  # noqa: E402

SEQ_DIR = "/root/reference/sequences"
OUT = os.path.dirname(os.path.abspath(__file__))

for name in ("fr1xyz", "fr1desk", "fr2robot2"):
    bal = BALProblem.load(os.path.join(SEQ_DIR, name + ".txt"))
    np.savez_compressed(
        os.path.join(OUT, f"seq_{name}.npz"),
        intrinsics=bal.intrinsics.copy(),
        cam_idx=bal.camera_index.astype(np.uint16),
        lmk_idx=bal.point_index.astype(np.uint16),
        observations=bal.observations.copy(),
        parameters=bal.parameters.copy(),
        dims=np.array([bal.n_keyframes, bal.n_points, bal.n_edges], dtype=np.int64),
    )
    print(name, bal.n_keyframes, bal.n_points, bal.n_edges)
