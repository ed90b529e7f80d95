"""CPU oracle for the LS / ELS / bbELS analytic score machines.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`convolutional_diffusion_b200/`) may import this package.  The only
permitted importers are `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs, and there only as
the checker / the reported CPU baseline, never as the thing that is
shipped or measured as the product.

Parity status: the reference ships NO golden vectors or known-answer tests
for this path, so the oracle is pinned against outputs of
the reference itself, executed in the build container by
`oracle/make_golden.py` (which imports `/root/reference/src/utils/idealscore.py`
behind a matplotlib stub) and committed as `tests/golden/*.npz`.
`tests/test_oracle_golden.py` checks every oracle entry point against
those fixtures.

Modules
-------
score_oracle   numpy float64 brute-force restatement of the unified
               masked-softmax form.
score_port     torch float32 CPU port that follows the reference's own
               operator structure (unfold + conv2d + streaming softmax);
               it is what `bench.py` times as the CPU baseline.
ref_loader     imports the real reference when /root/reference exists
               (build container only; never on the GPU box).
"""
