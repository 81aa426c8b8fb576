"""``from utils.dataset import get_dataset_and_loader`` (eval_MoCoDAD.py:9; utils/dataset.py:286-330) -> device ingest."""
from mocodad_b200.loader import DeviceBatchLoader, TrajectoryWindowDataset, get_dataset_and_loader  # noqa: F401
