"""``from utils.model_utils import processing_data`` (predict_MoCoDAD.py:10; utils/model_utils.py:110-137)."""
from mocodad_b200.mocodad import _processing_data as processing_data  # noqa: F401
