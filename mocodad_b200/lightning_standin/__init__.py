"""A stand-in for the slice of ``pytorch_lightning`` the reference's scoring scripts touch, for machines without Lightning
(neither the build container nor the B200 boxes have it).  ``mocodad_b200.mocodad`` falls back to it when
``import pytorch_lightning`` fails; ``python -m mocodad_b200.dropin`` puts this directory on ``sys.path`` in that case, so that
``eval_MoCoDAD.py``'s own ``import pytorch_lightning as pl`` resolves here.  With Lightning installed none of this is used."""
