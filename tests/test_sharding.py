"""Window sharding + the single score all-gather, on CPU with gloo and world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mocodad_b200.sharding import gather_scores, shard_bounds


@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (7, 2), (1024, 8), (1000003, 8), (5, 8)])
def test_shard_bounds_partition_the_range(n, world):
    spans = [shard_bounds(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(n, world, world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_total, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float32) * 0.5  # "score" of global window i is i/2
        full = gather_scores(local, n_total)
        q.put((rank, full.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 10, 1])
def test_gather_scores_world_size_2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [0.5 * i for i in range(n_total)]
    assert got[0] == want and got[1] == want


def test_single_process_passthrough():
    x = torch.arange(5.0)
    assert gather_scores(x, 5) is x


class _FakeEngine:
    """Stands in for the CUDA engine in the host-logic test below: the 'score' of dataset item i is i / 4."""

    def score_trajectories_host(self, coords, win_start, center, scale, vid_res, n_generated_samples, *, num_transform, batch, seed,
                                item_range):
        lo, hi = item_range
        assert 0 <= lo <= hi <= num_transform * len(win_start)
        return torch.arange(lo, hi, dtype=torch.float32) * 0.25


def _traj_worker(rank, world, port, root, q):
    import argparse
    import numpy as np
    from mocodad_b200 import MoCoDAD
    from test_module import BASE
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = dict(BASE, data_dir=os.path.join(root, "data"), vid_res=[640, 360], ckpt_dir=os.path.join(root, "ckpt"), num_transform=5)
        model = MoCoDAD(argparse.Namespace(**cfg))
        model.engine = lambda: _FakeEngine()          # host logic only: sharding of the dataset index space + the one all-gather
        scores, trans, meta, frames = model.score_trajectories(batch=16)
        q.put((rank, scores.tolist(), trans.tolist(), meta.shape, frames.shape))
    finally:
        dist.destroy_process_group()


def test_score_trajectories_shards_the_dataset_index_space_world_size_2(tmp_path):
    import pickle
    import numpy as np
    from sklearn.preprocessing import RobustScaler
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "trajectories.npz"))
    folder = tmp_path / "data" / "testing" / "trajectories" / "01-0200"
    folder.mkdir(parents=True)
    (tmp_path / "ckpt").mkdir()
    n_rows = int(g["lengths"][0])
    rows = np.concatenate([g["frames"][:n_rows, None].astype(np.float64), g["coords"][:n_rows].astype(np.float64)], axis=1)
    np.savetxt(folder / "0001.csv", rows, delimiter=",", fmt=["%d"] + ["%.2f"] * 34)
    sk = RobustScaler(quantile_range=(10.0, 90.0))
    sk.center_, sk.scale_ = g["center"].astype(np.float32), g["scale"]
    with open(tmp_path / "ckpt" / "local_robust.pickle", "wb") as fh:
        pickle.dump(sk, fh)
    n_win = n_rows - 6 + 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_traj_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r: rest for r, *rest in (q.get(timeout=180) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [0.25 * i for i in range(5 * n_win)]
    for r in range(2):
        scores, trans, meta_shape, frames_shape = got[r]
        assert scores == want                                   # every rank ends with all scores, in dataset order
        assert trans == [i // n_win for i in range(5 * n_win)]
        assert meta_shape == (5 * n_win, 4) and frames_shape == (5 * n_win, 6)
