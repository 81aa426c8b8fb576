"""A caller of the reference's public entry surface, used where the reference checkout is absent (the GPU box): the same
imports and the same sequence of calls a user's evaluation script makes against aleflabo/MoCoDAD -- config file -> namespace
-> init_args -> model class chosen by the `diffusion_on_latent` key -> saved-tensor shortcut or dataset + loader ->
Trainer.test with the checkpoint.  Run through `python -m mocodad_b200.dropin tests/eval_caller.py -c cfg.yaml`, the names
below resolve to the B200 overlay; run inside the reference tree they resolve to the reference.  (In the build container the
tests also run the reference's own, unmodified eval_MoCoDAD.py through the same launcher.)"""
import argparse
import os

import pytorch_lightning as pl
import yaml
from models.mocodad import MoCoDAD
from models.mocodad_latent import MoCoDADlatent
from utils.argparser import init_args
from utils.dataset import get_dataset_and_loader

if __name__ == "__main__":
    cli = argparse.ArgumentParser()
    cli.add_argument("-c", "--config", required=True)
    with open(cli.parse_args().config) as fh:
        args = init_args(argparse.Namespace(**yaml.safe_load(fh)))
    model_cls = MoCoDADlatent if hasattr(args, "diffusion_on_latent") else MoCoDAD
    model = model_cls(args)
    if args.load_tensors:
        auc = model.test_on_saved_tensors(split_name=args.split)
    else:
        _, loader, _, _ = get_dataset_and_loader(args, split=args.split)
        trainer = pl.Trainer(accelerator=args.accelerator, devices=args.devices[:1], default_root_dir=args.ckpt_dir, max_epochs=1,
                             logger=False)
        out = trainer.test(model, dataloaders=loader, ckpt_path=os.path.join(args.ckpt_dir, args.load_ckpt))
