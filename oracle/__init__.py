"""CPU oracle for the stage-2 mask-training hot path of Compress-Robust-VQA.

TEST INFRASTRUCTURE ONLY.  Nothing under compress-robust-vqa_b200/ imports this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may, and there only as the
checker or as the CPU arm being timed -- never as the product path.

What it is: a plain-PyTorch-on-CPU / numpy restatement of the reference's algorithm for the path
(masking/maskers.py MaskedLinear1 + _Binarizer1, Trainer.reset_threshold, the BCE / LPF / LMH losses,
the LXMERT forward around the masked Linear call sites, clip + the root AdamW), each function citing
the reference file:line it follows (paths relative to the reference root).

Pinning: the reference ships NO tests, golden vectors or fixtures; the
only third-party boundary is PyTorch itself.  The oracle is therefore pinned against outputs of the
UNMODIFIED reference modules imported in the build container: tests/golden/make_golden.py (committed)
runs them on seeded synthetic inputs and writes tests/golden/*.pt; tests/test_oracle_golden.py checks
every oracle function against those files, and SURVEY.md 8(c)'s known-answer values are asserted too.
mplug_masking.py (the mPLUG masker / threshold refresh) is pinned the same way by tests/golden/make_golden_mplug.py ->
mplug_skeleton.pt (tests/test_mplug_cpu.py).
"""
