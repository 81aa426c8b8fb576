"""Pins oracle/ingest.py against the UNMODIFIED reference (utils/dataset_utils.py) and writes tests/golden/transforms.npz.
Run in the build container (needs /root/reference): python oracle/make_transform_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
np.int = int  # the reference module uses the alias numpy 2 removed (dataset_utils.py:17,240); not on the path restated here

from utils.dataset_utils import ae_trans_list  # noqa: E402  (the reference)
from oracle import ingest  # noqa: E402

rng = np.random.default_rng(7)
N, T, V = 6, 6, 17
base = rng.standard_normal((N, 3, T, V)).astype(np.float32)
base[:, 2] = 1.0                       # PoseDatasetRobust fills the third channel with ones (dataset.py:247-249)
base[:, :2][rng.random((N, 2, T, V)) < 0.1] = 0.0   # missing joints are exact zeros
mats_ref = np.stack([np.asarray(t.trans_mat) for t in ae_trans_list[:5]]).astype(np.float32)
mats = ingest.ae_trans_mats(5)
assert mats.tobytes() == mats_ref.tobytes(), "transform matrices differ from the reference"
items = []
for idx in range(5 * N):
    sample, trans = idx % N, idx // N
    ref = ae_trans_list[trans](np.array(base[sample]))[:2]
    got = ingest.dataset_item(base, idx, mats)
    assert ref.dtype == got.dtype and ref.tobytes() == got.tobytes(), (idx, np.abs(ref - got).max())
    items.append(ref)
out = os.path.join(ROOT, "tests", "golden", "transforms.npz")
np.savez_compressed(out, base=base, mats=mats_ref, items=np.stack(items).astype(np.float32))
print("oracle/ingest.py is bit-identical to the reference on", 5 * N, "dataset items; wrote", out)
