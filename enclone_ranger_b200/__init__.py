"""enclone_ranger_b200 — B200-native clonotype join (join_exacts hot path of enclone_ranger)."""
from .params import default_params  # noqa: F401
from .inputs import JoinInput, generate, synth_config  # noqa: F401
from .join import JoinContext, JoinResult, EquivRel, JoinError, join_exacts, equiv_replay  # noqa: F401,E402
