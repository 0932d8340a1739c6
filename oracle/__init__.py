"""oracle/ — TEST INFRASTRUCTURE ONLY.

A CPU restatement (torch fp32 / numpy fp64, functional style) of the reference's
denoiser path: train_options/models_original.py + the three timm 0.9.16 classes
it imports (PatchEmbed, Attention, Mlp) + diffusion/.  Every function cites the
reference file:line it follows.

Nothing under fast_dit_b200/ imports this package.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
call it, and only as the checker or the CPU baseline — never as the product.

Parity pin: the reference has no tests or golden vectors of its own (SURVEY.md
§4).  This oracle is pinned instead against the *reference itself*, imported
unmodified in the build container (with oracle/timm_standin on sys.path because
timm is not installed) by tools/gen_golden.py, which wrote tests/golden/*.npz;
tests/test_oracle_golden.py re-checks the oracle against those fixtures on every
run, and tests/test_reference_live.py re-checks against the live reference
whenever /root/reference is present.
"""
