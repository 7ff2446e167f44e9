"""CPU oracle for GraphEcho's data-parallel hot path.  TEST INFRASTRUCTURE — not product code.

Plain-PyTorch (fp32, CPU) restatements of the reference's algorithms for every row of
SURVEY.md §8(a), written functionally (tensors + parameter dicts keyed by the reference's
state_dict names).  Each function cites the reference file:line it follows
(paths relative to the reference tree, xmed-lab/GraphEcho @ 2b6c4755).

Parity pinning: the reference ships no tests, golden vectors or fixtures, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the build container
by `oracle/make_golden.py` (which imports /root/reference unmodified) and committed as small
fixtures under tests/golden/.  `tests/test_oracle_golden.py` checks every oracle function
against those fixtures on CPU.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package, and only as the checker / the reported CPU baseline.  Nothing under
graphecho_b200/ imports it.
"""
