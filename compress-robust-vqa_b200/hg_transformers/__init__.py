"""Stage-2 mask-training surface of the reference's ``hg_transformers`` package.

Only the modules on the hot path exist here: the three mask trainers,
the LXMERT / VisualBERT models whose Linear/Embedding call sites get masked, the answer head
and the debias losses.  The reference's vendored model zoo is out of scope.
"""
__version__ = "2.10.0"
