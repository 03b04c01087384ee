"""Drop-in ``masking`` package of Compress-Robust-VQA's stage 2 (reference: masking/*.py), with the
masked-module bodies routed to the sm_100a kernels of libcrvqa.so."""
