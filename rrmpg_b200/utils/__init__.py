"""Host-side helpers mirroring ``rrmpg.utils`` as far as the hot path needs them."""
from .array_checks import check_for_negatives, validate_array_input
from .metrics import calc_kge, calc_mse, calc_nse, calc_rmse

__all__ = ["check_for_negatives", "validate_array_input", "calc_kge", "calc_mse", "calc_nse", "calc_rmse"]
