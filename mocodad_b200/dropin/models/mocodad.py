"""``from models.mocodad import MoCoDAD`` (eval_MoCoDAD.py:6) -> the B200 module."""
from mocodad_b200.mocodad import MoCoDAD  # noqa: F401
