"""``from models.mocodad_latent import MoCoDADlatent`` (eval_MoCoDAD.py:7) -> the B200 module."""
from mocodad_b200.mocodad import MoCoDADlatent  # noqa: F401
