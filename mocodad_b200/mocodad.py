"""Host-side mirror of the reference's ``MoCoDAD`` LightningModule for the inference/scoring path.

Same constructor (``MoCoDAD(args)`` from the YAML ``Namespace``), same ``forward`` /
``test_step`` / ``validation_step`` / epoch hooks, same ``state_dict`` names (Lightning checkpoints
load unchanged) as models/mocodad.py:22-334 -- but ``forward`` hands the whole reverse-diffusion
loop (mocodad.py:129-184) to the sm_100a CUDA library through the C ABI.  There is no PyTorch
implementation of the arithmetic in this package: on a machine without the library or without a
CUDA device ``forward`` raises.

Scope: conditioning strategies 'inject' and 'no_condition' (every shipped
config uses 'inject'); 'concat' / 'inbetween_imp' / 'random_imp' raise NotImplementedError, as do
the training entry points.  Optional new knobs are read with ``getattr(args, ..., default)``:
  b200_rng    'philox' (default; in-kernel counter-based noise keyed by seed / window / sample /
              step) or 'torch' (noise drawn with torch.randn in the reference's call order
              mocodad.py:162,176 and injected, so a seeded run reproduces the reference's draws)
  seed        Philox key (the YAML key the reference carries but never uses at eval time)
"""
from __future__ import annotations

import argparse
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import engine as _engine
from .params import state_dict_spec

try:  # the reference's orchestration layer, when installed
    import pytorch_lightning as _pl
except Exception:  # pragma: no cover - Lightning is absent in the build container and on the B200 boxes
    from .lightning_standin import pytorch_lightning as _pl
_Base = _pl.LightningModule


class _ParamNode(nn.Module):
    """Plain container: gives parameters/buffers the dotted names of the reference modules."""


def _build_param_tree(root: nn.Module, spec) -> None:
    for name, shape in spec.items():
        node = root
        parts = name.split(".")
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, _ParamNode())
            node = node._modules[part]
        leaf = parts[-1]
        if leaf == "num_batches_tracked":
            node.register_buffer(leaf, torch.tensor(0, dtype=torch.long))
        elif leaf in ("running_mean", "running_var"):
            node.register_buffer(leaf, torch.zeros(shape) if leaf == "running_mean" else torch.ones(shape))
        else:
            node.register_parameter(leaf, nn.Parameter(_init_like_reference(name, shape, spec), requires_grad=False))


def _init_like_reference(name: str, shape, spec) -> torch.Tensor:
    """Fresh-module values with the reference's init distributions (stsgcn.py:135-140 for A/T;
    torch defaults for conv / linear / BatchNorm / PReLU).  Checkpoints overwrite all of them."""
    leaf = name.rsplit(".", 1)[-1]
    if leaf in ("A", "T"):
        bound = 1.0 / float(shape[1]) ** 0.5
        return torch.empty(shape).uniform_(-bound, bound)
    if name.endswith("prelu.weight"):
        return torch.full(shape, 0.25)
    if ".tcn.1." in name or ".residual.1." in name or ".block.1." in name:  # BatchNorm affine
        return torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
    wshape = spec[name[: -len(leaf)] + "weight"]
    fan_in = 1
    for d in wshape[1:]:
        fan_in *= d
    bound = 1.0 / float(fan_in) ** 0.5
    return torch.empty(shape).uniform_(-bound, bound)


class MoCoDAD(_Base):
    """Drop-in for ``models.mocodad.MoCoDAD`` on the test / validation / predict path."""

    losses = ("l1", "smooth_l1", "mse")  # mocodad.py:24
    conditioning_strategies = {'cat': 'concat', 'concat': 'concat',  # mocodad.py:25-29
                               'add2layers': 'inject', 'inject': 'inject',
                               'inbetween_imp': 'inbetween_imp', 'interleave': 'inbetween_imp',
                               'random_indices': 'random_imp', 'random_imp': 'random_imp',
                               'no_condition': 'no_condition', 'none': 'no_condition'}

    def __init__(self, args: argparse.Namespace) -> None:
        super().__init__()
        self.save_hyperparameters(args)
        # Data parameters (mocodad.py:46-48)
        self.n_frames = args.seg_len
        self.num_coords = args.num_coords
        self.n_joints = self._infer_number_of_joint(args)
        # Model parameters (mocodad.py:51-62)
        self.embedding_dim = args.embedding_dim
        self.dropout = args.dropout
        self.conditioning_strategy = self.conditioning_strategies[args.conditioning_strategy]
        self.conditioning_indices = args.conditioning_indices
        self.n_frames_condition, self.n_frames_corrupt, self.input_n_frames = self._set_conditioning_strategy()
        self.conditioning_architecture = args.conditioning_architecture if self.conditioning_strategy == 'inject' else None
        self.cond_h_dim = args.h_dim
        self.cond_latent_dim = args.latent_dim
        self.cond_channels = args.channels
        self.cond_dropout = args.dropout
        # Training and inference parameters (mocodad.py:65-81)
        self.learning_rate = args.opt_lr
        if args.loss_fn not in self.losses:
            raise KeyError(args.loss_fn)
        self.loss_fn_name = args.loss_fn
        self.rec_weight = args.rec_weight
        self.noise_steps = args.noise_steps
        self.aggregation_strategy = args.aggregation_strategy
        self.n_generated_samples = args.n_generated_samples
        self.model_return_value = args.model_return_value
        self.gt_path = args.gt_path
        self.split = args.split
        self.use_hr = args.use_hr
        self.ckpt_dir = args.ckpt_dir
        self.save_tensors = args.save_tensors
        self.num_transforms = args.num_transform
        self.anomaly_score_pad_size = args.pad_size
        self.anomaly_score_filter_kernel_size = args.filter_kernel_size
        self.anomaly_score_frames_shift = args.frames_shift
        self.dataset_name = args.dataset_choice
        # where the trajectories live and their video resolution: only the device-ingest entry points read them
        # (the reference hands them to get_dataset_and_loader, utils/dataset.py:270-331)
        self.data_dir = getattr(args, "data_dir", None)
        self.vid_res = getattr(args, "vid_res", None)
        self.seg_stride = int(getattr(args, "seg_stride", 1) or 1)   # train split only (utils/dataset.py:308)
        # knobs new to this implementation (optional)
        self.rng_mode = getattr(args, "b200_rng", "philox")
        if self.rng_mode not in ("philox", "torch"):
            raise ValueError(f"b200_rng must be 'philox' or 'torch', got {self.rng_mode!r}")
        self.seed = int(getattr(args, "seed", 999) or 0)
        self._window_cursor = 0  # global index of the next window this module scores (Philox counter)

        self._set_diffusion_variables()
        self.build_model()

    # ------------------------------------------------------------------ construction
    def build_model(self) -> None:
        """mocodad.py:90-126 -- here: the parameter tree (reference names) + a lazily built engine."""
        if self.conditioning_strategy == 'inject':
            if self.conditioning_architecture not in ('AE', 'E'):
                raise NotImplementedError(f'Conditioning architecture {self.conditioning_architecture} not implemented.')
            if self.cond_latent_dim != self.embedding_dim:
                raise ValueError("latent_dim must equal embedding_dim: the condition embedding is added to the "
                                 "time embedding (stsae_unet.py:425-426)")
        if self.n_joints != 17:
            raise NotImplementedError(f"{self.n_joints} joints: the denoiser's joint pyramid is fixed at 17/12/10 "
                                      "(stsae_unet.py:11); headless / kp18 layouts fail in the reference too")
        spec = state_dict_spec(T=self.input_n_frames, T_cond=self.n_frames_condition, num_coords=self.num_coords,
                               embedding_dim=self.embedding_dim, h_dim=self.cond_h_dim, latent_dim=self.cond_latent_dim,
                               channels=self.cond_channels, conditioning_architecture=self.conditioning_architecture,
                               n_joints=self.n_joints)
        self._spec = spec
        _build_param_tree(self, spec)
        self._engine: Optional[_engine.ScoringEngine] = None
        self._engine_key = None

    def _set_conditioning_strategy(self) -> Tuple[int, int, int]:
        """mocodad.py:753-796 restricted to the strategies this path implements."""
        input_n_frames = self.n_frames
        if self.conditioning_strategy == 'no_condition':
            n_frames_cond = 0
        elif self.conditioning_strategy == 'inject':
            if isinstance(self.conditioning_indices, int):
                n_frames_cond = self.n_frames // self.conditioning_indices
                self._cond_first = True
            else:
                idx = list(self.conditioning_indices)
                assert idx == list(range(min(idx), max(idx) + 1)), \
                    'Conditioning indices must be a list of consecutive integers'
                assert (min(idx) == 0) or (max(idx) == (self.n_frames - 1)), \
                    'Conditioning indices must start from 0 or end at the last frame'
                n_frames_cond = len(idx)
                self._cond_first = min(idx) == 0
            input_n_frames = self.n_frames - n_frames_cond
        elif self.conditioning_strategy in ('concat', 'inbetween_imp', 'random_imp'):
            raise NotImplementedError(
                f"conditioning strategy '{self.conditioning_strategy}' is outside the B200 scoring path "
                "(no shipped config uses it; SURVEY.md section 2)")
        else:
            raise NotImplementedError(f'Conditioning strategy {self.conditioning_strategy} not implemented')
        if not hasattr(self, "_cond_first"):
            self._cond_first = True
        return n_frames_cond, self.n_frames - n_frames_cond, input_n_frames

    def _set_diffusion_variables(self) -> None:
        """mocodad.py:799-808; the table comes from the library's host function (bit-identical)."""
        self._beta_, self._alpha_, self._alpha_hat_ = _engine.schedule(self.noise_steps)

    def _infer_number_of_joint(self, args: argparse.Namespace) -> int:
        """mocodad.py:563-580"""
        if args.headless:
            return 14
        if args.kp18_format:
            return 18
        return 17

    # ------------------------------------------------------------------ engine
    def _weights_key(self):
        dev = self.device
        ver = 0
        for t in list(self.parameters()) + list(self.buffers()):
            ver += t._version
        first = next(self.parameters())
        return (str(dev), ver, first.data_ptr())

    def engine(self) -> _engine.ScoringEngine:
        """The CUDA engine for the module's current device and weights (re-packed when they change)."""
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("MoCoDAD (B200) computes only on a CUDA device: move the module with .to('cuda'); "
                               "there is no CPU fallback")
        key = self._weights_key()
        if self._engine is None or self._engine.device != torch.device(dev.type, dev.index if dev.index is not None else torch.cuda.current_device()):
            self._engine = _engine.ScoringEngine(
                seg_len=self.n_frames, n_frames_cond=self.n_frames_condition, cond_first=self._cond_first,
                noise_steps=self.noise_steps, loss_fn=self.loss_fn_name, embedding_dim=self.embedding_dim,
                h_dim=self.cond_h_dim, channels=self.cond_channels, device=dev, n_joints=self.n_joints,
                num_coords=self.num_coords)
            self._engine_key = None
        if key != self._engine_key:
            self._engine.load_state_dict(self.state_dict())
            self._engine_key = key
        return self._engine

    # ------------------------------------------------------------------ forward
    def forward(self, input_data: List[torch.Tensor], aggr_strategy: str = None, return_: str = None) -> List[torch.Tensor]:
        """mocodad.py:129-184.  ``input_data`` = [data [B,C,seg_len,V], transformation_idx, metadata,
        actual_frames]; returns [loss and/or selected poses] + [data, transformation_idx, metadata, frames]."""
        tensor_data, meta_out = self._unpack_data(input_data)
        eng = self.engine()
        data = tensor_data.to(torch.float32).contiguous()
        B = data.shape[0]
        G = self.n_generated_samples
        aggr = self.aggregation_strategy if aggr_strategy is None else aggr_strategy
        if return_ is None:
            if self.model_return_value is None:
                raise ValueError('Either return_ or self.model_return_value must be set')
            return_ = self.model_return_value
        need_pose = return_ in ('pose', 'all')

        noise = None
        if self.rng_mode == "torch":  # the reference's draw order: per sample x_T then z_{N-1}..z_2
            slots = max(self.noise_steps - 1, 1)
            noise = torch.empty(G, slots, B, self.num_coords, self.input_n_frames, self.n_joints, device=data.device)
            for g in range(G):
                for k in range(self.noise_steps - 1 if self.noise_steps > 1 else 1):
                    noise[g, k] = torch.randn(B, self.num_coords, self.input_n_frames, self.n_joints, device=data.device)
        first_window = self._window_cursor
        self._window_cursor += B

        if aggr == 'random':  # mocodad.py:479-480 returns a bare sample
            res = eng.reverse_diffusion(data, G, noise=noise, seed=self.seed, first_window=first_window, want_samples=True)
            return res["x0"][np.random.randint(G)]
        simple = aggr in ('best', 'worst') and not need_pose
        res = eng.reverse_diffusion(data, G, noise=noise, seed=self.seed, first_window=first_window,
                                    want_losses=not simple, want_worst=(aggr == 'worst'),
                                    want_samples=need_pose or aggr in ('all', 'mean_pose', 'median_pose'))
        selected_x, loss = self._aggregate(eng, res, data, aggr, need_pose)
        return self._pack_out_data(selected_x, loss, [tensor_data] + meta_out, return_=return_)

    def _aggregate(self, eng, res, data, aggr: str, need_pose: bool):
        """mocodad.py:454-520 on the per-sample losses [G,B] (and samples) the library produced."""
        if aggr in ('best', 'worst'):
            loss = res['best'] if aggr == 'best' else res['worst']
            sel = None
            if need_pose:
                losses = res['losses']
                idx = torch.argmin(losses, dim=0) if aggr == 'best' else torch.argmax(losses, dim=0)
                sel = res['x0'][idx, torch.arange(losses.shape[1], device=losses.device)]
            return sel, loss
        losses = res['losses']
        if aggr == 'all':
            return res['x0'].permute(1, 0, 2, 3, 4), losses.permute(1, 0)
        if aggr == 'mean':
            return None, torch.mean(losses, dim=0)
        if aggr == 'median':
            return None, torch.median(losses, dim=0)[0]
        if aggr in ('mean_pose', 'median_pose'):
            sel = torch.mean(res['x0'], dim=0) if aggr == 'mean_pose' else torch.median(res['x0'], dim=0)[0]
            sel = sel.contiguous()
            return sel, eng.window_loss(sel, data, 1)['best']
        if 'quantile' in aggr:
            q = float(aggr.split(':')[-1])
            return None, torch.quantile(losses, q, dim=0)
        raise ValueError(f'Unknown aggregation strategy {aggr}')

    def _pack_out_data(self, selected_x, loss_of_selected_x, additional_out, return_: str):
        """mocodad.py:606-636"""
        if return_ == 'pose':
            out = [selected_x]
        elif return_ == 'loss':
            out = [loss_of_selected_x]
        elif return_ == 'all':
            out = [loss_of_selected_x, selected_x]
        else:
            raise ValueError(f"unknown return_ {return_!r}")
        return out + additional_out

    def _unpack_data(self, x) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        """mocodad.py:843-858"""
        return x[0].to(self.device), [x[1], x[2], x[3]]

    # ------------------------------------------------------------------ Lightning surface
    def test_step(self, batch: List[torch.Tensor], batch_idx: int) -> None:
        self._test_output_list.append(self.forward(batch))

    def on_test_epoch_start(self) -> None:
        super().on_test_epoch_start()
        self._test_output_list = []
        self._window_cursor = 0

    def on_test_epoch_end(self) -> float:
        auc = self._epoch_end(self._test_output_list)
        del self._test_output_list
        self.log('AUC', auc)
        return auc

    def validation_step(self, batch: List[torch.Tensor], batch_idx: int) -> None:
        self._validation_output_list.append(self.forward(batch))

    def on_validation_epoch_start(self) -> None:
        super().on_validation_epoch_start()
        self._validation_output_list = []
        self._window_cursor = 0

    def on_validation_epoch_end(self) -> float:
        auc = self._epoch_end(self._validation_output_list)
        del self._validation_output_list
        self.log('AUC', auc, sync_dist=True)
        return auc

    def predict_step(self, batch, batch_idx: int = 0, dataloader_idx: int = 0):
        return self.forward(batch)

    def training_step(self, batch, batch_idx):
        raise NotImplementedError("training is outside the B200 scoring path (SURVEY.md section 8 f4); "
                                  "train with the reference and load the checkpoint here")

    def configure_optimizers(self):
        raise NotImplementedError("training is outside the B200 scoring path")

    def _epoch_end(self, outputs) -> float:
        out, gt_data, trans, meta, frames = _processing_data(outputs)
        if self.save_tensors:
            self._save_tensors({'prediction': out, 'gt_data': gt_data, 'trans': trans, 'metadata': meta, 'frames': frames},
                               split_name=self.split, aggr_strategy=self.aggregation_strategy, n_gen=self.n_generated_samples)
        return self.post_processing(out, gt_data, trans, meta, frames)

    def post_processing(self, out, gt_data, trans, meta, frames) -> float:
        """Score assembly -> frame-level AUC (mocodad.py:337-430): per clip / person max over windows, padding around
        absences, log-range mix over persons, HR masks, shift + Gaussian smoothing, mean over the affine
        transformations, ``roc_auc_score`` (mocodad_b200/postproc.py, pinned against the reference).  On a CUDA device the
        first stage -- the per-(clip, person) frame maxima over the epoch's windows -- runs on the GPU (``mcd_frame_scores``);
        the saved-tensor branch of a module that was never moved to a GPU assembles them on the host."""
        from . import postproc
        accel = self.engine().frame_scores_host if self.device.type == "cuda" else None
        gt = postproc.load_ground_truth(self.gt_path)
        clip_masks = postproc.hr_ubnormal_masks(self.split) if (self.use_hr and self.dataset_name == 'UBnormal') else None
        avenue = postproc.avenue_hr_mask() if self.dataset_name == 'HR-Avenue' else None
        return float(postproc.dataset_auc(np.asarray(out), np.asarray(trans), np.asarray(meta), np.asarray(frames), gt,
                                          num_transform=self.num_transforms, pad_size=self.anomaly_score_pad_size,
                                          frames_shift=self.anomaly_score_frames_shift,
                                          filter_kernel_size=self.anomaly_score_filter_kernel_size,
                                          clip_masks=clip_masks, avenue_masks=avenue, frame_scores=accel))

    def test_on_saved_tensors(self, split_name: str) -> float:
        """mocodad.py:433-448"""
        tensors = self._load_tensors(split_name, self.aggregation_strategy, self.n_generated_samples)
        auc_score = self.post_processing(tensors['prediction'], tensors['gt_data'], tensors['trans'],
                                         tensors['metadata'], tensors['frames'])
        print(f'AUC score: {auc_score:.6f}')
        return auc_score

    def score_trajectories(self, data_dir: str = None, vid_res=None, split: str = None, batch: int = 1024):
        """The test epoch straight from the reference's on-disk trajectories, bypassing the host
        dataset: what ``get_dataset_and_loader`` (utils/dataset.py:270-331 -> PoseDatasetRobust) + ``trainer.test`` produce,
        with the frame rows uploaded once and every dataset item built in HBM (``ScoringEngine.score_trajectories_host``).
        Needs ``{ckpt_dir}/local_robust.pickle`` like the reference's test split (get_robust_data.py:115-127).  Under an
        initialised ``torch.distributed`` group every rank scores its contiguous shard of the dataset index space and one
        all-gather assembles the scores.  Returns (loss [K*N], trans [K*N], meta [K*N,4], frames [K*N,seg_len]) as numpy,
        in dataset order -- the inputs of ``post_processing``."""
        from . import ingest, sharding
        import torch.distributed as dist
        if self.aggregation_strategy != 'best':
            raise NotImplementedError("score_trajectories returns the 'best' aggregation (every shipped test config); use "
                                      "forward() on batches for the other strategies")
        split = self.split if split is None else split
        data_dir = self.data_dir if data_dir is None else data_dir
        vid_res = self.vid_res if vid_res is None else vid_res
        if data_dir is None or vid_res is None or len(vid_res) != 2:
            raise ValueError("score_trajectories needs data_dir and vid_res = [width, height] (arguments or YAML keys)")
        ts = ingest.load_trajectories(os.path.join(data_dir, ingest.split_subfolder(split), 'trajectories'))
        starts, meta, frames = ingest.window_table(ts, self.n_frames, 1)       # no strides for the test set (dataset.py:308)
        if split == 'validation' and 'UBnormal' not in data_dir:
            # get_robust_data.py:120-123: outside UBnormal the validation split is scaled by a scaler fitted on itself
            # (pickled as local_robust_val.pickle), not by the training run's
            scaler = self.engine().fit_scaler_host(ts.coords, ts.lengths, vid_res, seg_stride=1)
            import pickle
            with open(os.path.join(self.ckpt_dir, 'local_robust_val.pickle'), 'wb') as fh:
                pickle.dump(scaler, fh)
            center, scale = ingest.scaler_arrays(scaler)
        else:
            center, scale = ingest.load_robust_scaler(self.ckpt_dir)
        if int(self.num_transforms) < 1:
            raise NotImplementedError("num_transform < 1: the reference's untransformed dataset path applies a random temporal crop "
                                      "per item (utils/dataset.py:77-83, 127-130), which has no device counterpart")
        K, N = int(self.num_transforms), len(starts)
        total = K * N
        rank, world = (dist.get_rank(), dist.get_world_size()) if (dist.is_available() and dist.is_initialized()) else (0, 1)
        lo, hi = sharding.shard_bounds(total, rank, world)
        local = self.engine().score_trajectories_host(ts.coords, starts, center, scale, vid_res, self.n_generated_samples,
                                                      num_transform=K, batch=batch, seed=self.seed, item_range=(lo, hi))
        scores = sharding.gather_scores(local.to(self.device), total).cpu().numpy() if world > 1 else local.numpy()
        trans = np.repeat(np.arange(K, dtype=np.int64), N)                     # item idx -> idx // N (dataset.py:70-72)
        return scores, trans, np.tile(meta, (K, 1)), np.tile(frames, (K, 1))

    def fit_trajectory_scaler(self, data_dir: str = None, vid_res=None, split: str = 'train'):
        """The train split's scaler step (utils/get_robust_data.py:115-119): fit the RobustScaler on the training trajectories
        (rows normalised on the device) and pickle it as ``{ckpt_dir}/local_robust.pickle`` for ``score_trajectories``."""
        from . import ingest
        data_dir = self.data_dir if data_dir is None else data_dir
        vid_res = self.vid_res if vid_res is None else vid_res
        if data_dir is None or vid_res is None or len(vid_res) != 2:
            raise ValueError("fit_trajectory_scaler needs data_dir and vid_res = [width, height] (arguments or YAML keys)")
        ts = ingest.load_trajectories(os.path.join(data_dir, ingest.split_subfolder(split), 'trajectories'))
        # the train split drops trajectories too short for one STRIDED window before the fit (get_robust_data.py:44,58:
        # input_gap = seg_stride - 1; utils/dataset.py:308 hands seg_stride to the train split only)
        stride = self.seg_stride if 'train' in split else 1
        return self.engine().fit_scaler_host(ts.coords, ts.lengths, vid_res, seg_stride=stride, exp_dir=self.ckpt_dir)

    def test_on_trajectories(self, data_dir: str = None, vid_res=None, split: str = None, batch: int = 1024) -> float:
        """``score_trajectories`` -> ``post_processing`` -> AUC: the device-ingest twin of eval_MoCoDAD.py:30-38."""
        scores, trans, meta, frames = self.score_trajectories(data_dir, vid_res, split=split, batch=batch)
        auc_score = self.post_processing(scores, None, trans, meta, frames)
        self.log('AUC', auc_score)
        return auc_score

    def _load_tensors(self, split_name: str, aggr_strategy: str, n_gen: int) -> Dict[str, torch.Tensor]:
        """mocodad.py:583-603"""
        path = os.path.join(self.ckpt_dir, 'saved_tensors_{}_{}_{}'.format(split_name, aggr_strategy, n_gen))
        return {f.split('.')[0]: torch.load(os.path.join(path, f), weights_only=False) for f in os.listdir(path)}   # numpy arrays, like upstream

    def _save_tensors(self, tensors, split_name: str, aggr_strategy: str, n_gen: int) -> None:
        """mocodad.py:689-705"""
        path = os.path.join(self.ckpt_dir, 'saved_tensors_{}_{}_{}'.format(split_name, aggr_strategy, n_gen))
        os.makedirs(path, exist_ok=True)
        for t_name, tensor in tensors.items():
            torch.save(tensor, os.path.join(path, t_name + '.pt'))

    # schedule accessors with the reference's names (mocodad.py:861-873)
    @property
    def _beta(self) -> torch.Tensor:
        return self._beta_.to(self.device)

    @property
    def _alpha(self) -> torch.Tensor:
        return self._alpha_.to(self.device)

    @property
    def _alpha_hat(self) -> torch.Tensor:
        return self._alpha_hat_.to(self.device)


def _processing_data(data):
    """utils/model_utils.py:110-137 -- concatenate the per-batch outputs on the host."""
    cols = [[], [], [], [], []]
    for arr in data:
        for c, t in zip(cols, arr[:5]):
            c.append(t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t))
    return tuple(np.concatenate(c, axis=0) for c in cols)


class MoCoDADlatent(MoCoDAD):
    """Drop-in for ``models.mocodad_latent.MoCoDADlatent`` at stage 'diffusion' (the shipped
    config/UBnormal/mocodad-latent_test.yaml; ``eval_MoCoDAD.py:24`` picks it when the YAML carries ``diffusion_on_latent``):
    the diffusion runs on a ``latent_embedding_dim`` vector -- the corrupt frames go once through the down half of the denoiser
    (STSE_Unet at the constant step -1) and ``n_generated_samples`` x ``noise_steps - 1`` calls of an MLP denoiser follow
    (models/mocodad_latent.py:69-132).  Same 292-entry ``state_dict`` as the reference module.  Stage 'pretrain' is training
    only and raises."""

    def __init__(self, args: argparse.Namespace) -> None:
        # mocodad_latent.py:24-28
        self.stage = args.stage
        self.latent_embedding_dim = args.latent_embedding_dim
        self.hidden_sizes = args.hidden_sizes
        self.pretrained_model_ckpt_path = args.pretrained_model_ckpt_path
        if self.stage not in ('diffusion', 'pretrain'):
            raise ValueError(f'Unknown stage {self.stage}')
        if self.stage == 'pretrain':
            raise NotImplementedError("MoCoDADlatent stage 'pretrain' is a training stage (SURVEY.md section 8 row f4); "
                                      "train with the reference and score with stage 'diffusion' here")
        super().__init__(args)
        assert self.conditioning_strategy == 'inject', \
            'Conditioning strategy must be inject. Other strategies are not supported for the latent space'
        # mocodad_latent.py:36-38: the frozen main net comes from the pretraining checkpoint
        if self.pretrained_model_ckpt_path:
            self._freeze_main_net_and_load_ckpt()

    def build_model(self) -> None:
        """mocodad_latent.py:42-66 -- parameter tree of STSE_Unet (+ to_time_dim) and the MLP Denoiser."""
        if self.conditioning_architecture not in ('AE', 'E'):
            raise NotImplementedError(f'Conditioning architecture {self.conditioning_architecture} not implemented.')
        if self.cond_latent_dim != self.embedding_dim:
            raise ValueError("latent_dim must equal embedding_dim (the condition embedding is added to the time embedding)")
        if self.n_joints != 17:
            raise NotImplementedError(f"{self.n_joints} joints: the joint pyramid is fixed at 17/12/10 (stsae_unet.py:11)")
        if list(self.hidden_sizes)[-1] != self.latent_embedding_dim:
            raise ValueError("hidden_sizes[-1] must equal latent_embedding_dim: the denoiser predicts noise of the latent's shape "
                             "(the reference's DDPM update broadcasts otherwise, mocodad_latent.py:119)")
        self._spec = state_dict_spec(T=self.input_n_frames, T_cond=self.n_frames_condition, num_coords=self.num_coords,
                                     embedding_dim=self.embedding_dim, h_dim=self.cond_h_dim, latent_dim=self.cond_latent_dim,
                                     channels=self.cond_channels, conditioning_architecture=self.conditioning_architecture,
                                     n_joints=self.n_joints, latent_embedding_dim=self.latent_embedding_dim,
                                     hidden_sizes=self.hidden_sizes)
        _build_param_tree(self, self._spec)
        self._engine = None
        self._engine_key = None

    def engine(self) -> _engine.ScoringEngine:
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("MoCoDADlatent (B200) computes only on a CUDA device: move the module with .to('cuda'); "
                               "there is no CPU fallback")
        key = self._weights_key()
        if self._engine is None or self._engine.device != torch.device(dev.type, dev.index if dev.index is not None else torch.cuda.current_device()):
            self._engine = _engine.ScoringEngine(
                seg_len=self.n_frames, n_frames_cond=self.n_frames_condition, cond_first=self._cond_first,
                noise_steps=self.noise_steps, loss_fn=self.loss_fn_name, embedding_dim=self.embedding_dim,
                h_dim=self.cond_h_dim, channels=self.cond_channels, device=dev, n_joints=self.n_joints,
                num_coords=self.num_coords, latent_embedding_dim=self.latent_embedding_dim, hidden_sizes=self.hidden_sizes)
            self._engine_key = None
        if key != self._engine_key:
            self._engine.load_state_dict(self.state_dict())
            self._engine_key = key
        return self._engine

    def forward(self, input_data: List[torch.Tensor], condition_data: torch.Tensor = None, aggr_strategy: str = 'best',
                *, return_: str = None) -> List[torch.Tensor]:
        """mocodad_latent.py:69-132 (stage 'diffusion'): [loss and/or selected latents] + [data, transformation_idx, metadata,
        frames].  As upstream, ``aggr_strategy`` defaults to 'best' here (the YAML key is not consulted)."""
        tensor_data, meta_out = self._unpack_data(input_data)
        eng = self.engine()
        data = tensor_data.to(torch.float32).contiguous()
        B, G, L = data.shape[0], self.n_generated_samples, self.latent_embedding_dim
        if return_ is None:
            if self.model_return_value is None:
                raise ValueError('Either return_ or self.model_return_value must be set')
            return_ = self.model_return_value
        need_latent = return_ in ('pose', 'all')
        noise = None
        if self.rng_mode == "torch":   # the reference's draw order: per sample x_T (torch.randn, :104) then z (:117)
            slots = max(self.noise_steps - 1, 1)
            noise = torch.empty(G, slots, B, L, device=data.device)
            for g in range(G):
                for k in range(self.noise_steps - 1 if self.noise_steps > 1 else 1):
                    noise[g, k] = torch.randn(B, L, device=data.device)
        first_window = self._window_cursor
        self._window_cursor += B
        aggr = aggr_strategy
        simple = aggr in ('best', 'worst') and not need_latent
        res = eng.latent_reverse_diffusion(data, G, noise=noise, seed=self.seed, first_window=first_window,
                                           want_losses=not simple, want_worst=(aggr == 'worst'),
                                           want_samples=need_latent or aggr in ('all', 'mean_pose', 'median_pose', 'random'))
        if aggr == 'random':
            return res["x0"][np.random.randint(G)]
        selected, loss = self._aggregate_latent(res, aggr, need_latent)
        return self._pack_out_data(selected, loss, [tensor_data] + meta_out, return_=return_)

    def _aggregate_latent(self, res, aggr: str, need_latent: bool):
        """mocodad.py:454-520 on latent vectors (the loss target is the latent code)."""
        if aggr in ('best', 'worst'):
            loss = res['best'] if aggr == 'best' else res['worst']
            sel = None
            if need_latent:
                losses = res['losses']
                idx = torch.argmin(losses, dim=0) if aggr == 'best' else torch.argmax(losses, dim=0)
                sel = res['x0'][idx, torch.arange(losses.shape[1], device=losses.device)]
            return sel, loss
        losses = res['losses']
        if aggr == 'all':
            return res['x0'].permute(1, 0, 2), losses.permute(1, 0)
        if aggr == 'mean':
            return None, torch.mean(losses, dim=0)
        if aggr == 'median':
            return None, torch.median(losses, dim=0)[0]
        if aggr in ('mean_pose', 'median_pose'):
            sel = torch.mean(res['x0'], dim=0) if aggr == 'mean_pose' else torch.median(res['x0'], dim=0)[0]
            d = sel - res['code']
            if self.loss_fn_name == 'smooth_l1':
                per = torch.where(d.abs() < 1.0, 0.5 * d * d, d.abs() - 0.5)
            else:
                per = d.abs() if self.loss_fn_name == 'l1' else d * d
            return sel, per.mean(dim=-1)
        if 'quantile' in aggr:
            return None, torch.quantile(losses, float(aggr.split(':')[-1]), dim=0)
        raise ValueError(f'Unknown aggregation strategy {aggr}')

    def score_trajectories(self, *a, **k):
        raise NotImplementedError("device ingest (score_trajectories) is wired to the pose-space model; feed MoCoDADlatent batches")

    def _freeze_main_net_and_load_ckpt(self) -> None:
        """mocodad_latent.py:222-227"""
        self.load_state_dict(torch.load(self.pretrained_model_ckpt_path, map_location='cpu', weights_only=False)['state_dict'], strict=False)
