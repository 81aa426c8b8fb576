"""``pytorch_lightning`` stand-in: ``LightningModule`` and ``Trainer`` as far as eval_MoCoDAD.py:36-38 and
predict_MoCoDAD.py use them (``Trainer(accelerator, devices, default_root_dir, max_epochs, logger)``,
``.test(model, dataloaders=, ckpt_path=)``, ``.predict(...)``) and as far as models/mocodad.py:22-334 relies on the base class
(``save_hyperparameters``, ``log``, ``device``, the ``on_*_epoch_start`` hooks).  Orchestration only: no arithmetic.

What ``Trainer.test`` does, in Lightning's order: load ``ckpt['state_dict']`` into the module (strict), move it to the device,
``eval()``, ``torch.inference_mode()``, ``on_test_epoch_start`` -> ``test_step(batch, i)`` per batch (tensors of the batch moved
to the device) -> ``on_test_epoch_end``; returns ``[{name: value logged with self.log}]``."""
from __future__ import annotations

import argparse
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

__version__ = "0.0-standin"


class LightningModule(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.hparams = argparse.Namespace()
        self._logged: Dict[str, float] = {}

    @property
    def device(self) -> torch.device:
        for p in self.parameters():
            return p.device
        for b in self.buffers():
            return b.device
        return torch.device("cpu")

    def save_hyperparameters(self, args=None, *a, **k) -> None:
        if args is not None:
            self.hparams = args

    def log(self, name, value, *a, **k) -> None:
        self._logged[name] = float(value)

    def on_test_epoch_start(self) -> None:
        pass

    def on_validation_epoch_start(self) -> None:
        pass

    def on_test_epoch_end(self) -> None:
        pass

    def on_validation_epoch_end(self) -> None:
        pass

    def predict_step(self, batch, batch_idx: int = 0, dataloader_idx: int = 0):
        return self(batch)


def _to_device(obj: Any, device: torch.device) -> Any:
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=True)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(o, device) for o in obj)
    if isinstance(obj, dict):
        return {k: _to_device(v, device) for k, v in obj.items()}
    return obj


class Trainer:
    def __init__(self, accelerator: str = "auto", devices=None, default_root_dir: Optional[str] = None, max_epochs: Optional[int] = None,
                 logger=None, **kwargs) -> None:
        self.accelerator, self.devices, self.default_root_dir = accelerator, devices, default_root_dir
        self.callback_metrics: Dict[str, float] = {}

    def _device(self) -> torch.device:
        acc = str(self.accelerator).lower()
        if acc == "cpu":
            return torch.device("cpu")
        if acc in ("gpu", "cuda") or (acc == "auto" and torch.cuda.is_available()):
            if not torch.cuda.is_available():
                raise RuntimeError("Trainer(accelerator='gpu'): no CUDA device is available")
            dev = self.devices
            if isinstance(dev, (list, tuple)):
                index = int(dev[0]) if dev else 0
            elif isinstance(dev, int):
                index = 0          # Lightning: an int is a COUNT of devices
            else:
                index = 0
            return torch.device("cuda", index)
        return torch.device("cpu")

    @staticmethod
    def _load(model: nn.Module, ckpt_path: Optional[str]) -> None:
        if ckpt_path:
            ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
            model.load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt)

    def test(self, model: nn.Module, dataloaders=None, ckpt_path: Optional[str] = None, verbose: bool = True) -> List[Dict[str, float]]:
        self._load(model, ckpt_path)
        device = self._device()
        model.to(device)
        model.eval()
        if hasattr(model, "_logged"):
            model._logged.clear()
        with torch.inference_mode():
            model.on_test_epoch_start()
            for i, batch in enumerate(dataloaders):
                model.test_step(_to_device(batch, device), i)
            model.on_test_epoch_end()
        self.callback_metrics = dict(getattr(model, "_logged", {}))
        if verbose:
            for k, v in self.callback_metrics.items():
                print(f"{k}: {v:.6f}")
        return [dict(self.callback_metrics)]

    def predict(self, model: nn.Module, dataloaders=None, ckpt_path: Optional[str] = None) -> List[Any]:
        self._load(model, ckpt_path)
        device = self._device()
        model.to(device)
        model.eval()
        outs = []
        with torch.inference_mode():
            for i, batch in enumerate(dataloaders):
                outs.append(model.predict_step(_to_device(batch, device), i))
        return outs
