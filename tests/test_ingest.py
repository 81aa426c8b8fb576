"""Row f1: dataset items = affine transforms of base windows.  The CPU part pins the oracle restatement
(oracle/ingest.py) to the fixture written from the unmodified reference; the GPU part checks mcd_expand_transforms."""
import os

import numpy as np
import pytest
import torch

from oracle import ingest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "transforms.npz")


def test_oracle_matches_reference_fixture():
    g = np.load(GOLD)
    base, mats, items = g["base"], g["mats"], g["items"]
    assert np.array_equal(ingest.ae_trans_mats(5), mats)
    n = base.shape[0]
    for idx in range(5 * n):
        got = ingest.dataset_item(base, idx, mats)
        assert got.dtype == np.float32 and got.tobytes() == items[idx].tobytes(), idx


def test_identity_and_flip_semantics():
    rng = np.random.default_rng(0)
    w = rng.standard_normal((3, 6, 17)).astype(np.float32)
    mats = ingest.ae_trans_mats(5)
    assert np.array_equal(ingest.apply_pose_transform(w, mats[0])[:2], w[:2])
    flipped = ingest.apply_pose_transform(w, mats[1])
    assert np.array_equal(flipped[0], -w[0]) and np.array_equal(flipped[1], w[1]) and np.array_equal(flipped[2], w[2])


@pytest.mark.gpu
def test_expand_transforms_vs_reference_fixture():
    from mocodad_b200 import ScoringEngine, synthetic as synth
    from mocodad_b200.engine import pose_transform_matrices
    g = np.load(GOLD)
    base, items = g["base"], g["items"]
    n = base.shape[0]
    eng = ScoringEngine(seg_len=6, n_frames_cond=3, noise_steps=4, device="cuda:0")
    eng.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=3, T_cond=3), seed=0))
    d_base = torch.from_numpy(np.ascontiguousarray(base[:, :2])).cuda()
    mats = pose_transform_matrices(5)
    whole = eng.expand_transforms(d_base, mats, 0, 5 * n).cpu().numpy()
    assert np.abs(whole - items).max() <= 1e-6  # products and sums rounded separately like the einsum; tolerance 1 ulp-class
    assert np.array_equal(whole[:n], base[:, :2])  # identity transform is exact
    part = eng.expand_transforms(d_base, mats, 7, 11).cpu().numpy()  # ragged range straddling transforms
    assert np.array_equal(part, whole[7:18])
    with pytest.raises(Exception):
        eng.expand_transforms(d_base, mats, 5 * n - 1, 2)


@pytest.mark.gpu
def test_score_dataset_host_equals_scoring_the_materialised_dataset():
    from mocodad_b200 import ScoringEngine, synthetic as synth
    from mocodad_b200.engine import pose_transform_matrices
    eng = ScoringEngine(seg_len=6, n_frames_cond=3, noise_steps=4, device="cuda:0")
    eng.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=3, T_cond=3), seed=0))
    base = synth.synth_batch(37, 6, seed=5)[0]
    got = eng.score_dataset_host(base, 2, num_transform=5, batch=64, seed=11)
    items = eng.expand_transforms(base.cuda(), pose_transform_matrices(5), 0, 5 * 37)
    want = eng.reverse_diffusion(items, 2, seed=11, first_window=0)["best"].cpu()
    assert got.shape == (185,) and torch.equal(got, want)
