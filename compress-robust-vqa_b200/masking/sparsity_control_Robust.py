"""Drop-in for the reference's ``masking/sparsity_control_Robust.py`` (identical to sparsity_control
but for one comment line)."""
from .sparsity_control import MaskerScheduler, automated_gradual_sparsity, stepwise_sparsity  # noqa: F401
