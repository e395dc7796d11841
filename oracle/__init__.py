"""CPU restatement of the reference's algorithm for the hot path — TEST INFRASTRUCTURE ONLY.

Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it, and there only as the checker (or as the timed
CPU baseline), never as the thing shipped.  The product path (nans_clip_b200) has no CPU fallback.

Pinning: the reference (n571e/NanS-CLIP) ships no tests or golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced by
importing /root/reference in the build container (tests/golden/make_golden.py) and committed as
fixtures under tests/golden/.  tests/test_oracle_golden.py checks every function here against them.
"""
