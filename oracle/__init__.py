"""CPU oracle for the py-cs advection hot path -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference algorithm (luanfs/py-cubed-sphere,
`/root/reference/src/*.py`), written independently of the product code in
`py-cubed-sphere_b200/`.  Every function cites the reference file:line it
follows.  Array layout is the reference's: `[i][j][panel]`, C order.

Who may import this package: `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` -- as the checker or
as the timed CPU baseline, never as part of the product path.  The product
(`pycs_b200`) never imports it and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so
the pin is the reference itself, executed in the build container under stub
modules (`tests/golden/_refshim.py`); `tests/golden/make_golden.py` wrote the
fixtures in `tests/golden/*.npz` from that run and `tests/test_oracle_golden.py`
checks this oracle against them (bit-exact for index maps, <=1e-13 relative
for fields; in practice the fields agree bit for bit).
"""
