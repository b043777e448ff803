"""CPU oracle for the embedding-extraction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``3d-speaker_b200/b200spk``).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and
only as the checker or the reported CPU baseline.

Each module restates one stage of the reference and cites the reference
file:line it follows.  The restatements are pinned against golden vectors
minted from the *imported* reference (``oracle/gen_golden.py`` ->
``tests/golden/*.npz``); the reference has no tests or golden vectors of its own
(SURVEY.md section 4), so that is the only pin available.
"""
