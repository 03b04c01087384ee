"""Drop-in for the reference's ``masking/maskers_Robust.py``: per-modality initial sparsity.

``chain_module_names`` returns (names, name_in_modal, name_in_module, name_in_layer) and ``Masker``
takes a leading ``hpmodel`` whose ``zerorate_dict[modal]`` is the initial sparsity of every module of
that modality (reference maskers_Robust.py:70-95, 491-642)."""
from ._core import (  # noqa: F401
    MaskedLinear0, MaskedLinear1, MaskedLinear2, MaskedLinear3, MaskedLinearX, MaskerBase,
    _Binarizer1, _Binarizer2, _Binarizer3, _bert_roberta_names, _distilbert_names, _get_nnz_from,
    _lxmert_names, _scheme_idx_to_fn, binarizer_fn1, binarizer_fn2, binarizer_fn3, chain_names_modal,
    finish_magnitude_init, reshape_mask_for_sp,
)


def chain_module_names(which_ptl, layer_idices, abbres):
    return chain_names_modal(_lxmert_names, which_ptl, layer_idices, abbres)


class Masker(MaskerBase):
    per_modal = True

    def __init__(self, hpmodel, masker_scheduler, logger, mask_biases, structured_masking_info, threshold,
                 init_scale, which_ptl, controlled_init):
        self._setup(masker_scheduler, logger, mask_biases, structured_masking_info, threshold, init_scale,
                    which_ptl, controlled_init, hpmodel=hpmodel)
