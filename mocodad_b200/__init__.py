"""mocodad_b200 -- B200-native (sm_100a) implementation of MoCoDAD's reverse-diffusion
anomaly-scoring path (reference: models/mocodad.py:129-184), behind a C ABI
(include/mocodad_b200.h) with a Python mirror of the reference's ``MoCoDAD`` module surface.

    from mocodad_b200 import MoCoDAD          # drop-in for models.mocodad.MoCoDAD (eval path)
    from mocodad_b200 import ScoringEngine    # thin torch-facing wrapper over the C ABI

The CUDA library is the only implementation; importing the compute classes without it fails.
"""
from .params import state_dict_spec  # noqa: F401


def __getattr__(name):  # lazy: keep `import mocodad_b200` cheap and torch-free until needed
    if name == "ScoringEngine":
        from .engine import ScoringEngine
        return ScoringEngine
    if name == "MoCoDAD":
        from .mocodad import MoCoDAD
        return MoCoDAD
    if name == "MoCoDADlatent":   # name kept for eval_MoCoDAD.py:24; raises NotImplementedError
        from .mocodad import MoCoDADlatent
        return MoCoDADlatent
    raise AttributeError(name)
