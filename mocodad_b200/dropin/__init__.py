"""Import overlay that makes the reference's scripts run UNCHANGED on the B200 path.

``eval_MoCoDAD.py`` / ``predict_MoCoDAD.py`` import ``models.mocodad.MoCoDAD``, ``models.mocodad_latent.MoCoDADlatent``,
``utils.argparser.init_args``, ``utils.dataset.get_dataset_and_loader`` and ``utils.model_utils.processing_data``
(eval_MoCoDAD.py:6-9, predict_MoCoDAD.py:7-10).  This directory holds packages of exactly those names, backed by
``mocodad_b200``.  Python puts a script's own directory first on ``sys.path``, so the overlay is activated by the launcher

    python -m mocodad_b200.dropin /path/to/MoCoDAD/eval_MoCoDAD.py -c config/Avenue/mocodad_test.yaml

which inserts this directory (and, when Lightning is not installed, ``mocodad_b200/lightning_standin``) in front of it and
then runs the script as ``__main__`` with the remaining arguments -- not a line of the script changes.  See INTEGRATION.md."""
import os

OVERLAY_DIR = os.path.dirname(os.path.abspath(__file__))
STANDIN_DIR = os.path.join(os.path.dirname(OVERLAY_DIR), "lightning_standin")


def overlay_paths():
    """``sys.path`` entries of the overlay: this directory, plus the Lightning stand-in when the real package is absent."""
    import importlib.util
    paths = [OVERLAY_DIR]
    try:
        have_pl = importlib.util.find_spec("pytorch_lightning") is not None
    except (ImportError, ValueError):
        have_pl = False
    if not have_pl:
        paths.append(STANDIN_DIR)
    return paths
