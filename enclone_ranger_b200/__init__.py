"""enclone_ranger_b200 — B200-native clonotype join (join_exacts hot path of enclone_ranger)."""
from .params import default_params  # noqa: F401
from .inputs import JoinInput, generate, synth_config  # noqa: F401
