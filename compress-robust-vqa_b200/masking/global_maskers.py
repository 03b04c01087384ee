"""Drop-in for the reference's ``masking/global_maskers.py``: the LXMERT masker whose magnitude init uses ONE
cut over the union of all masked weights (``compute_global_threshold`` :531-541, ``_magnitude_global`` :219-231)
instead of one per module.  Same public names; bodies in ``masking._core`` -> libcrvqa.so."""
from ._core import (  # noqa: F401
    MaskedLinear0, MaskedLinear1, MaskedLinear2, MaskedLinear3, MaskedLinearX, MaskerBase,
    _Binarizer1, _Binarizer2, _Binarizer3, _bert_roberta_names, _distilbert_names, _get_nnz_from,
    _lxmert_names, _scheme_idx_to_fn, binarizer_fn1, binarizer_fn2, binarizer_fn3, chain_names_plain,
    finish_magnitude_init, global_kth_value, reshape_mask_for_sp,
)


def chain_module_names(which_ptl, layer_idices, abbres):
    return chain_names_plain(_lxmert_names, which_ptl, layer_idices, abbres)


class Masker(MaskerBase):
    def __init__(self, masker_scheduler, logger, mask_biases, structured_masking_info, threshold, init_scale,
                 which_ptl, controlled_init, global_prune=True):
        self._setup(masker_scheduler, logger, mask_biases, structured_masking_info, threshold, init_scale,
                    which_ptl, controlled_init)
        self.global_prune = global_prune
        self.global_threshold = None
