"""mPLUG masked training (reference: mPLUG/, BASELINE config 5).

Either import as ``mPLUG.masking.maskers`` with the package root on ``sys.path``, or -- the reference's own
layout, which runs with ``mPLUG/`` as the working directory -- put this directory first on ``sys.path`` and import
``masking.maskers`` / ``vqa_mplug``.  Built: the masking package, the threshold refresh with the reference's bf16
thresholds, mask export / sparsity report, ``init_masker``, the scheduler-driven training loop and the engine that
stands where the DeepSpeed engine stands, the mPLUG-VQA network (``models/``: CLIP ViT, text / fusion / decoder stacks,
beam-search generation and closed-set ranking) and the evaluation helpers.
"""
