"""Synthetic checkpoints / windows / noise for the parity harness.

TEST INFRASTRUCTURE ONLY (see oracle/ref_port.py header).  The generators live in
``mocodad_b200/synthetic.py`` (bench.py and smoke() need the same inputs); this module re-exports
them so the harness has one import.  ``oracle/make_golden.py`` asserts that ``state_dict_spec``
equals the real reference module's ``state_dict`` (names, order, shapes).
"""
from mocodad_b200.synthetic import (state_dict_spec, synth_batch, synth_noise,  # noqa: F401
                                    synth_state_dict)
