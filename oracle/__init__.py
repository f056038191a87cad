"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's algorithm.

Nothing under ``oracle/`` is part of the product. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline; the product path (``tf_flash_attention_b200``) never does and
fails loudly when its CUDA library is missing.

Pinning status: **pinned**.
 * attended-index pattern: checked bit-for-bit against the reference's own host
   code (``sync_methods.cc`` + the policies in ``flash_attention.h``) compiled
   unmodified from /root/reference by ``oracle/ref_build/Makefile`` into
   ``oracle/_ref/ref_pattern``; its outputs are committed as
   ``tests/golden/pattern_*.json`` by ``tests/golden/make_pattern_golden.py``.
 * known answers: the alignment tables in the reference docstring
   (flash_attention/flash_attention.py:30-69) and the README example shapes.
 * numerics: the reference's own CUDA kernel, built unmodified into
   ``oracle/_ref/libref_fa.so`` and run on a B200
   (``tests/golden/make_refkernel_golden.py``), agrees with this oracle
   within the reference tests' own tolerance.
"""
