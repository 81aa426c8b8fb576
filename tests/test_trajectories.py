"""Row f1, second slice: trajectory frame rows -> dataset items.  tests/golden/trajectories.npz holds what the
UNMODIFIED reference (PoseDatasetRobust, utils/dataset.py:197-268) produced from a synthetic trajectory tree
(oracle/make_trajectory_golden.py).  CPU part: the oracle restatement and the host window table reproduce the fixture
bit for bit.  GPU part: mcd_normalize_frames / mcd_build_items reproduce it bit for bit (float32 arithmetic, every operation
rounded as numpy rounds it); transformed items within 1e-6 like test_ingest (the einsum's summation order is torch's)."""
import os

import numpy as np
import pytest
import torch

from mocodad_b200 import ingest
from oracle import trajectories as otr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "trajectories.npz")
CASES = [("L6", 6, 1, ""), ("L27", 27, 1, "_long"), ("L6s2", 6, 2, "")]


def _ts(g, sfx) -> ingest.TrajectorySet:
    return ingest.TrajectorySet(g["coords" + sfx], g["frames" + sfx], g["lengths" + sfx], g["ids" + sfx], [])


@pytest.mark.parametrize("tag,seg_len,stride,sfx", CASES)
def test_oracle_and_window_table_match_reference_fixture(tag, seg_len, stride, sfx):
    g = np.load(GOLD)
    ts = _ts(g, sfx)
    starts, meta, frames = ingest.window_table(ts, seg_len, stride)
    o_starts, o_meta, o_frames = otr.window_table(ts.lengths, ts.frames, ts.ids, seg_len, stride)
    assert np.array_equal(starts, o_starts) and np.array_equal(meta, o_meta) and np.array_equal(frames, o_frames)
    assert np.array_equal(meta, g["meta_" + tag]) and np.array_equal(frames, g["ids_" + tag])
    got = otr.base_windows(ts.coords, starts, seg_len, stride, g["center"], g["scale"], g["vid_res"])
    assert got.dtype == np.float32 and got[:, :2].tobytes() == g["base_" + tag].tobytes()
    assert np.all(got[:, 2] == 1.0)


def test_window_table_edge_cases():
    empty = ingest.TrajectorySet(np.zeros((0, 34), np.float32), np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros((0, 3), np.int64), [])
    s, m, f = ingest.window_table(empty, 6)
    assert s.shape == (0,) and m.shape == (0, 4) and f.shape == (0, 6)
    # ragged: one too-short trajectory between two usable ones; frame numbers with a gap
    lengths = np.array([7, 3, 6])
    frames = np.concatenate([np.arange(10, 17), np.arange(3), [1, 2, 4, 5, 6, 9]]).astype(np.int32)
    ts = ingest.TrajectorySet(np.ones((16, 34), np.float32), frames, lengths, np.array([[1, 2, 3], [1, 2, 4], [5, 6, 7]]), [])
    s, m, f = ingest.window_table(ts, 6)
    assert s.tolist() == [0, 1, 10]
    assert m.tolist() == [[1, 2, 3, 10], [1, 2, 3, 11], [5, 6, 7, 1]]
    assert f[2].tolist() == [1, 2, 4, 5, 6, 9]
    s2, _, f2 = ingest.window_table(ts, 3, 3)   # rows 0,3,6 of a 7-row trajectory only
    assert s2.tolist() == [0] and f2.tolist() == [[10, 13, 16]]
    with pytest.raises(ValueError):
        ingest.window_table(ts, 0)


def test_load_trajectories_reads_the_reference_layout(tmp_path):
    g = np.load(GOLD)
    root = tmp_path / "testing" / "trajectories"
    (root / "01-0200").mkdir(parents=True)
    rows = np.concatenate([g["frames"][:9, None].astype(np.float64), g["coords"][:9].astype(np.float64)], axis=1)
    np.savetxt(root / "01-0200" / "0003.csv", rows, delimiter=",", fmt=["%d"] + ["%.2f"] * 34)
    ts = ingest.load_trajectories(str(root))
    assert ts.ids.tolist() == [[1, 200, 3]] and ts.names == ["01-0200_0003"] and ts.lengths.tolist() == [9]
    assert np.array_equal(ts.coords, g["coords"][:9]) and np.array_equal(ts.frames, g["frames"][:9])
    assert ingest.split_subfolder("test") == "testing" and ingest.split_subfolder("validation") == "validating"
    (root / "01-0200" / "0004.csv").write_text("1,2,3\n")
    with pytest.raises(ValueError):
        ingest.load_trajectories(str(root))


def test_bbox_normalisation_properties():
    g = np.load(GOLD)
    out = otr.bbox_centre_normalize(g["coords"], g["vid_res"])
    assert np.all(np.abs(out) <= 0.5 + 1e-6)                     # inside the (enlarged) box, centred
    assert np.all(out[g["coords"] == 0] == 0)                    # missing joints stay exact zeros
    assert np.all(out[~g["coords"].any(axis=1)] == 0)            # missing frames too


def test_scaler_fit_matches_the_reference_pickle(tmp_path):
    """Train split: RobustScaler fitted on the normalised training rows == the estimator the reference pickled."""
    g = np.load(GOLD)
    rows = otr.bbox_centre_normalize(g["train_coords"], g["vid_res"])
    sk = ingest.fit_robust_scaler(rows, g["train_lengths"], 6, 1, exp_dir=str(tmp_path))
    assert sk.center_.dtype == np.float32 and np.array_equal(sk.center_.astype(np.float64), g["center"])
    assert np.array_equal(np.asarray(sk.scale_, dtype=np.float64), g["scale"])
    center, scale = ingest.load_robust_scaler(str(tmp_path))          # the pickle the test split loads
    assert np.array_equal(center, g["center"]) and np.array_equal(scale, g["scale"])
    assert (g["train_lengths"] < 6).any()                              # the short trajectory is excluded like upstream
    with pytest.raises(ValueError):
        ingest.fit_robust_scaler(rows[:-1], g["train_lengths"], 6)
    with pytest.raises(ValueError):
        ingest.fit_robust_scaler(rows, g["train_lengths"], 1000)


def _engine(seg_len):
    from mocodad_b200 import ScoringEngine, synthetic as synth
    eng = ScoringEngine(seg_len=seg_len, n_frames_cond=3, noise_steps=4, device="cuda:0")
    eng.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=seg_len - 3, T_cond=3), seed=0))
    return eng


@pytest.mark.gpu
def test_normalize_frames_bit_exact_vs_oracle():
    g = np.load(GOLD)
    eng = _engine(6)
    extra = otr.stress_rows(11)   # checked against the reference itself by oracle/make_trajectory_golden.py
    for coords, res in ((g["coords"], g["vid_res"]), (g["coords_long"], g["vid_res"]), (extra, np.array([640, 360], np.float32)),
                        (extra, np.array([1080, 720], np.float32))):
        want = otr.bbox_centre_normalize(coords, res)
        d = torch.from_numpy(coords).cuda()
        got = eng.normalize_frames(d, res).cpu().numpy()
        assert np.array_equal(got, want), np.abs(got - want).max()
        assert np.array_equal(eng.normalize_frames(d, res, out=d).cpu().numpy(), want)   # in place
    assert eng.normalize_frames(torch.empty(0, 34, device="cuda"), [640, 360]).shape == (0, 34)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,seg_len,stride,sfx", CASES)
def test_build_items_vs_reference_fixture(tag, seg_len, stride, sfx):
    from mocodad_b200.engine import pose_transform_matrices
    g = np.load(GOLD)
    ts = _ts(g, sfx)
    eng = _engine(seg_len)
    starts, _, _ = ingest.window_table(ts, seg_len, stride)
    n = len(starts)
    rows = eng.normalize_frames(torch.from_numpy(ts.coords).cuda(), g["vid_res"])
    d_start = torch.from_numpy(starts).cuda()
    base = eng.build_items(rows, d_start, g["center"], g["scale"], row_step=stride).cpu().numpy()
    assert base.shape == (n, 2, seg_len, 17) and base.tobytes() == g["base_" + tag].tobytes()     # the reference's segs_data_np[:, :2]
    mats = pose_transform_matrices(5)
    items = eng.build_items(rows, d_start, g["center"], g["scale"], mats=mats, row_step=stride).cpu().numpy()
    assert items.shape == (5 * n, 2, seg_len, 17)
    assert np.array_equal(items[:n], base)                                                       # identity transform is exact
    assert np.abs(items[g["item_idx_" + tag]] - g["items_" + tag]).max() <= 1e-6                 # the reference's __getitem__
    want = eng.expand_transforms(torch.from_numpy(base).cuda(), mats, 0, 5 * n).cpu().numpy()
    assert np.array_equal(items, want)                                                           # same as the two-step path
    part = eng.build_items(rows, d_start, g["center"], g["scale"], mats=mats, first_item=n - 3, n_items=n + 7, row_step=stride)
    assert np.array_equal(part.cpu().numpy(), items[n - 3:2 * n + 4])                            # ragged range across transforms
    with pytest.raises(Exception):
        eng.build_items(rows, d_start, g["center"], g["scale"], mats=mats, first_item=5 * n - 1, n_items=2, row_step=stride)
    with pytest.raises(Exception):
        eng.build_items(rows, d_start, g["center"], np.zeros(34), row_step=stride)
    # scaler fused into the normalisation (one pass over the rows) + pure gather: same bits
    scaled = eng.normalize_frames(torch.from_numpy(ts.coords).cuda(), g["vid_res"], center=g["center"], scale=g["scale"])
    assert np.array_equal(eng.build_items(scaled, d_start, row_step=stride).cpu().numpy(), base)
    assert np.array_equal(eng.build_items(scaled, d_start, mats=mats, row_step=stride).cpu().numpy(), items)
    odd = eng.build_items(scaled, d_start, mats=mats, first_item=1, n_items=5 * n - 2, row_step=stride)   # chunk tails, N not a multiple of 8
    assert np.array_equal(odd.cpu().numpy(), items[1:5 * n - 1])
    assert eng.build_items(scaled, d_start, mats=mats, first_item=3, n_items=0, row_step=stride).shape == (0, 2, seg_len, 17)


@pytest.mark.gpu
def test_score_trajectories_host_equals_scoring_the_materialised_dataset():
    from mocodad_b200.engine import pose_transform_matrices
    g = np.load(GOLD)
    ts = _ts(g, "")
    eng = _engine(6)
    starts, _, _ = ingest.window_table(ts, 6, 1)
    n = len(starts)
    got = eng.score_trajectories_host(ts.coords, starts, g["center"], g["scale"], g["vid_res"], 2, num_transform=5, batch=64, seed=11)
    base = torch.from_numpy(g["base_L6"]).cuda()
    items = eng.expand_transforms(base, pose_transform_matrices(5), 0, 5 * n)
    want = eng.reverse_diffusion(items, 2, seed=11, first_window=0)["best"].cpu()
    assert got.shape == (5 * n,) and torch.equal(got, want)
    # a rank's shard of the dataset index space gives the same scores
    lo, hi = 2 * n - 5, 4 * n + 9
    shard = eng.score_trajectories_host(ts.coords, starts, g["center"], g["scale"], g["vid_res"], 2, batch=50, seed=11, item_range=(lo, hi))
    assert torch.equal(shard, want[lo:hi])
    with pytest.raises(ValueError):
        eng.score_trajectories_host(ts.coords, starts + 10_000, g["center"], g["scale"], g["vid_res"], 2)


@pytest.mark.gpu
def test_module_scores_a_trajectory_tree_like_the_batch_path(tmp_path):
    """MoCoDAD.test_on_trajectories (data dir -> AUC, items built in HBM) == forward() over the materialised dataset."""
    import argparse
    import pickle
    from sklearn.preprocessing import RobustScaler
    from mocodad_b200 import MoCoDAD, synthetic as synth
    from test_module import BASE
    g = np.load(GOLD)
    ts = _ts(g, "")
    # the reference's on-disk layout: {data_dir}/testing/trajectories/{scene}-{clip}/{person}.csv, {ckpt_dir}/local_robust.pickle, gt masks
    data_dir, ckpt_dir, gt_dir = tmp_path / "data", tmp_path / "ckpt", tmp_path / "gt"
    ckpt_dir.mkdir()
    gt_dir.mkdir()
    row0 = 0
    rng = np.random.default_rng(2)
    for k, n in enumerate(ts.lengths):
        scene, clip, person = (int(v) for v in ts.ids[k])
        folder = data_dir / "testing" / "trajectories" / f"{scene:02d}-{clip:04d}"
        folder.mkdir(parents=True, exist_ok=True)
        rows = np.concatenate([ts.frames[row0:row0 + n, None].astype(np.float64), ts.coords[row0:row0 + n].astype(np.float64)], axis=1)
        np.savetxt(folder / f"{person:04d}.csv", rows, delimiter=",", fmt=["%d"] + ["%.2f"] * 34)
        np.save(gt_dir / f"{scene:02d}_{clip:04d}.npy", (rng.random(int(ts.frames.max()) + 8) < 0.3).astype(np.int64))
        row0 += n
    sk = RobustScaler(quantile_range=(10.0, 90.0))
    sk.center_, sk.scale_ = g["center"].astype(np.float32), g["scale"]
    with open(ckpt_dir / "local_robust.pickle", "wb") as fh:
        pickle.dump(sk, fh)
    cfg = dict(BASE, noise_steps=4, n_generated_samples=2, gt_path=str(gt_dir), ckpt_dir=str(ckpt_dir), dataset_choice="HR-STC",
               pad_size=3, filter_kernel_size=5, frames_shift=2)
    model = MoCoDAD(argparse.Namespace(**cfg))
    model.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=3, T_cond=3), seed=0))
    model = model.to("cuda:0")
    scores, trans, meta, frames = model.score_trajectories(str(data_dir), g["vid_res"], batch=100)
    # the directory walk may visit trajectories in another order than the fixture: compare through the window identity
    ts2 = ingest.load_trajectories(str(data_dir / "testing" / "trajectories"))
    starts2, meta2, frames2 = ingest.window_table(ts2, 6, 1)
    n = len(starts2)
    assert scores.shape == (5 * n,) and np.array_equal(meta[:n], meta2) and np.array_equal(frames[n:2 * n], frames2)
    assert np.array_equal(trans, np.repeat(np.arange(5), n))
    base = otr.base_windows(ts2.coords, starts2, 6, 1, g["center"], g["scale"], g["vid_res"])[:, :2]
    eng = model.engine()
    from mocodad_b200.engine import pose_transform_matrices
    items = eng.expand_transforms(torch.from_numpy(np.ascontiguousarray(base)).cuda(), pose_transform_matrices(5), 0, 5 * n)
    model.on_test_epoch_start()
    for i0 in range(0, 5 * n, 64):
        sl = slice(i0, min(i0 + 64, 5 * n))
        model.test_step([items[sl], torch.from_numpy(trans[sl]), torch.from_numpy(meta[sl]), torch.from_numpy(frames[sl])], 0)
    want = np.concatenate([o[0].cpu().numpy() for o in model._test_output_list])
    assert np.array_equal(scores, want)
    auc_batches = model.on_test_epoch_end()
    auc = model.test_on_trajectories(str(data_dir), g["vid_res"], batch=77)
    assert 0.0 <= auc <= 1.0 and auc == auc_batches


@pytest.mark.gpu
def test_scaler_fit_on_device_normalised_rows_matches_the_reference_pickle():
    g = np.load(GOLD)
    eng = _engine(6)
    sk = eng.fit_scaler_host(g["train_coords"], g["train_lengths"], g["vid_res"])
    assert np.array_equal(sk.center_.astype(np.float64), g["center"]) and np.array_equal(np.asarray(sk.scale_, np.float64), g["scale"])


@pytest.mark.gpu
def test_scaler_fit_of_a_strided_train_split_matches_the_reference_pickle(tmp_path):
    """config/STC/mocodad_train.yaml has seg_stride 6: the reference drops every trajectory too short for one STRIDED window
    before fitting (get_robust_data.py:44,58).  Fixture: the unmodified reference's pickle at seg_len 6, seg_stride 3."""
    import argparse
    import pickle
    from mocodad_b200 import MoCoDAD, synthetic as synth
    from test_module import BASE
    g = np.load(GOLD)
    eng = _engine(6)
    sk = eng.fit_scaler_host(g["train_coords"], g["train_lengths"], g["vid_res"], seg_stride=3)
    assert np.array_equal(np.asarray(sk.center_, np.float64), g["center_s3"]) and np.array_equal(np.asarray(sk.scale_, np.float64), g["scale_s3"])
    assert not np.array_equal(g["center_s3"], g["center"])
    # through the module: the YAML's seg_stride reaches the fit of the 'train' split (and only that split)
    data_dir, ckpt_dir = tmp_path / "data", tmp_path / "ckpt"
    ckpt_dir.mkdir()
    row0 = 0
    for k, n in enumerate(g["train_lengths"]):
        folder = data_dir / "training" / "trajectories" / f"01-{k + 1:04d}"
        folder.mkdir(parents=True)
        rows = np.concatenate([np.arange(1, n + 1, dtype=np.float64)[:, None], g["train_coords"][row0:row0 + n].astype(np.float64)], axis=1)
        np.savetxt(folder / "0001.csv", rows, delimiter=",", fmt=["%d"] + ["%.2f"] * 34)
        row0 += n
    cfg = dict(BASE, ckpt_dir=str(ckpt_dir), data_dir=str(data_dir), vid_res=[float(v) for v in g["vid_res"]], seg_stride=3, split="train")
    model = MoCoDAD(argparse.Namespace(**cfg))
    model.load_state_dict(synth.synth_state_dict(synth.state_dict_spec(T=3, T_cond=3), seed=0))
    model.to("cuda:0").fit_trajectory_scaler()
    with open(ckpt_dir / "local_robust.pickle", "rb") as fh:
        got = pickle.load(fh)
    assert np.array_equal(np.asarray(got.center_, np.float64), g["center_s3"]) and np.array_equal(np.asarray(got.scale_, np.float64), g["scale_s3"])


def test_window_table_equals_the_loop_restatement_on_random_trajectory_sets():
    """Vectorised host table (mocodad_b200/ingest.py) vs the loop restatement of preprocessing.py:55-86, ragged inputs."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(lengths=st.lists(st.integers(min_value=0, max_value=40), min_size=0, max_size=8),
           seg_len=st.integers(min_value=1, max_value=12), stride=st.integers(min_value=1, max_value=3), seed=st.integers(0, 2**16))
    def check(lengths, seg_len, stride, seed):
        rng = np.random.default_rng(seed)
        F = int(sum(lengths))
        frames = (np.cumsum(rng.integers(1, 3, size=F)) if F else np.zeros(0)).astype(np.int32)
        ids = rng.integers(0, 50, size=(len(lengths), 3)).astype(np.int64)
        ts = ingest.TrajectorySet(np.zeros((F, 34), np.float32), frames, np.asarray(lengths, dtype=np.int64), ids, [])
        s, m, f = ingest.window_table(ts, seg_len, stride)
        os_, om, of = otr.window_table(ts.lengths, ts.frames, ts.ids, seg_len, stride)
        assert s.dtype == np.int64 and np.array_equal(s, os_) and np.array_equal(m, om) and np.array_equal(f, of)
        if len(s):
            assert s.max() + (seg_len - 1) * stride < F and (np.diff(s) > 0).all()
    check()
