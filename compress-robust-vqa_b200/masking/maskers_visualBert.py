"""Drop-in for the reference's ``masking/maskers_visualBert.py``: the baseline masker reading the
VisualBERT name table (reference maskers_visualBert.py:24-34, 83-95)."""
from ._core import (  # noqa: F401
    MaskedLinear0, MaskedLinear1, MaskedLinear2, MaskedLinear3, MaskedLinearX, MaskerBase,
    _Binarizer1, _Binarizer2, _Binarizer3, _bert_roberta_names, _distilbert_names, _get_nnz_from,
    _lxmert_names, _scheme_idx_to_fn, _visualbert_names, binarizer_fn1, binarizer_fn2, binarizer_fn3,
    chain_names_plain, finish_magnitude_init, reshape_mask_for_sp,
)


def chain_module_names(which_ptl, layer_idices, abbres):
    return chain_names_plain(_visualbert_names, which_ptl, layer_idices, abbres)


class Masker(MaskerBase):
    def __init__(self, masker_scheduler, logger, mask_biases, structured_masking_info, threshold, init_scale,
                 which_ptl, controlled_init):
        self._setup(masker_scheduler, logger, mask_biases, structured_masking_info, threshold, init_scale,
                    which_ptl, controlled_init)
